"""Kernel-only throughput of SC / SCL-8 over the code lengths of the reference's parameter sweeps
(benchmarks/test_code_parameters.py:33: N = 128 .. 4096 at rate 1/2)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import polarcode_and_ldpc_b200 as P  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
for L in (1, 8):
    for N in (128, 256, 512, 1024, 2048, 4096):
        F = (1 << 28) // N if L == 1 else (1 << 27) // N
        w = dict(kind="polar", N=N, K=N // 2, L=L, snr=2.0, frames=F)
        llr, ref, code = bench.make_inputs(w, torch, dev, 1)
        dec = bench.make_decoder(w, code)
        for _ in range(2):
            bits = dec.decode_batch(llr)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(3):
            bits = dec.decode_batch(llr)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        info = dec.launch_info()
        print(f"L={L} N={N:5d} F={F:7d} {F * (N // 2) / ms / 1e6:8.2f} Gbps  G={info['global_levels']} compiled={info['compiled_code_length']}", flush=True)
        del dec, llr
