"""Resident blocks per SM (PCL_POLAR_BPS) x shared-memory budget for the run-time-N / round-1-layout polar
kernels: the level scratch of all resident frames should stay inside the 126 MB L2.  One line per (N, L)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import polarcode_and_ldpc_b200 as P  # noqa: E402

dev = torch.device("cuda", 0)
cases = [(1024, 1), (1024, 2), (1024, 4), (2048, 1), (2048, 2), (2048, 8), (512, 1), (512, 4), (1024, 16), (4096, 1), (4096, 8)]
if len(sys.argv) > 1:
    cases = [tuple(int(x) for x in c.split("x")) for c in sys.argv[1].split(",")]
for N, L in cases:
    F = max(8192, (1 << 28) // (N * L))
    w = dict(kind="polar", N=N, K=N // 2, L=L, snr=2.0, frames=F)
    llr, ref, code = bench.make_inputs(w, torch, dev, 1)
    for budget in (8192, 16384):
        res = []
        for bps in (0, 1, 2, 3, 4, 5):
            os.environ["PCL_POLAR_SMEM_PER_WARP"] = str(budget)
            os.environ["PCL_POLAR_BPS"] = str(bps)
            os.environ["PCL_POLAR_L2FIT"] = "0"
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
            li = dec.launch_info()
            res.append(f"bps{bps}: {F * (N // 2) / ms / 1e6:6.2f} (grid {li['grid']}, G={li['global_levels']})")
            del dec
        print(f"N={N} L={L} budget {budget}  " + "  ".join(res), flush=True)
