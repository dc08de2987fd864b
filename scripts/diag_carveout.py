"""Why did the same decode kernel run at 8.0 (bench) and 10.0 Gbps (kernel-only loop) in one gpurun call?

Times a workload's decode kernel inside loops that differ only in what runs BETWEEN two decodes:
nothing / the library's count_errors kernel / a small torch elementwise kernel / a cudaMemsetAsync.
    python scripts/diag_carveout.py scl8 bp504
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import polarcode_and_ldpc_b200 as P  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
for name in sys.argv[1:] or ["scl8", "bp504"]:
    w = dict(bench.WORKLOADS[name])
    llr, ref, code = bench.make_inputs(w, torch, dev, 1)
    dec = bench.make_decoder(w, code)
    cnt = P.ErrorCounters(1, device=dev)
    small = torch.zeros(1024, device=dev)
    big = torch.zeros(64 << 20, dtype=torch.uint8, device=dev)

    def between(kind, bits):
        if kind == "count":
            cnt.add(0, bits, ref, None)
        elif kind == "torch_small":
            small.add_(1.0)
        elif kind == "torch_big":
            big.add_(1)
        elif kind == "memset":
            small.zero_()

    for kind in ("none", "count", "torch_small", "torch_big", "memset", "none", "count"):
        for _ in range(3):
            between(kind, dec.decode_batch(llr))
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(22)]
        torch.cuda.synchronize()
        for s in range(10):
            ev[2 * s].record()
            bits = dec.decode_batch(llr)
            ev[2 * s + 1].record()
            between(kind, bits)
        torch.cuda.synchronize()
        ms = sorted(ev[2 * s].elapsed_time(ev[2 * s + 1]) for s in range(10))
        g = llr.shape[0] * bench.info_bits(w) / 1e6
        print(f"{name:8s} between={kind:12s} decode kernel: median {g / ms[5]:6.3f} Gbps  best {g / ms[0]:6.3f}  worst {g / ms[-1]:6.3f}",
              flush=True)
