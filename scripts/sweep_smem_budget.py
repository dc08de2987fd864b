"""Shared-memory budget per warp (how many tree levels stay on chip) for the small list sizes: throughput per
(N, L, PCL_POLAR_SMEM_PER_WARP) on one GPU.  Picks the defaults in pcl_api.cu (polar_create_impl)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import polarcode_and_ldpc_b200 as P  # noqa: E402

dev = torch.device("cuda", 0)
budgets = [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else "8192,16384,24576,32768".split(","))]
for N in (256, 512, 1024, 2048, 4096):
    for L in (1, 2, 4):
        if N == 256 and L == 1:
            continue
        F = max(16384, (1 << 28) // (N * L))
        w = dict(kind="polar", N=N, K=N // 2, L=L, snr=2.0, frames=F)
        llr, ref, code = bench.make_inputs(w, torch, dev, 1)
        res = []
        for b in budgets:
            os.environ["PCL_POLAR_SMEM_PER_WARP"] = str(b)
            dec = P.SCDecoder(N, N // 2, frozen_bits=code["frozen"]) if L == 1 else P.SCLDecoder(N, N // 2, L, code["frozen"])
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
            ok = (bits == ref).all(dim=1).float().mean().item()
            li = dec.launch_info()
            res.append(f"{b}: {F * (N // 2) / ms / 1e6:7.2f} (G={li['global_levels']}, grid {li['grid']}, ok {ok:.3f})")
            del dec
        print(f"N={N} L={L}  " + "  ".join(res), flush=True)
