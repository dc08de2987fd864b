"""Time decode_batch for list sizes x compiled / run-time code length variants (PCL_POLAR_NL) on one GPU."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import polarcode_and_ldpc_b200 as P  # noqa: E402

dev = torch.device("cuda", 0)
for N, K, L, F in ((256, 128, 1, 524288), (1024, 512, 1, 262144), (1024, 512, 2, 262144), (1024, 512, 4, 131072),
                   (1024, 512, 8, 131072), (1024, 512, 16, 65536), (1024, 512, 32, 32768)):
    w = dict(kind="polar", N=N, K=K, L=L, snr=2.0, frames=F)
    llr, ref, code = bench.make_inputs(w, torch, dev, 1)
    for S in (1, 0):
        os.environ["PCL_POLAR_NL"] = str(S)
        dec = P.SCDecoder(N, K, frozen_bits=code["frozen"]) if L == 1 else P.SCLDecoder(N, K, L, code["frozen"])
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
        print(f"N={N} L={L} NL={S} {F * K / ms / 1e6:8.3f} Gbps  {F / ms / 1e3:8.2f} Mframes/s  frame-ok {ok:.4f} {dec.launch_info()}", flush=True)
        del dec
