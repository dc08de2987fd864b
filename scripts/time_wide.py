"""Throughput of the block-per-frame kernel (list sizes above 32) on one GPU."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import polarcode_and_ldpc_b200 as P  # noqa: E402

dev = torch.device("cuda", 0)
for N, K, L, F in ((1024, 512, 64, 8192), (1024, 512, 128, 4096), (1024, 512, 256, 2048), (256, 128, 64, 16384)):
    w = dict(kind="polar", N=N, K=K, L=L, snr=2.0, frames=F)
    llr, ref, code = bench.make_inputs(w, torch, dev, 1)
    dec = P.SCLDecoder(N, K, L, code["frozen"])
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
    print(f"N={N} L={L} {F * K / ms / 1e6:8.4f} Gbps  {F / ms:8.1f} kframes/s  frame-ok {ok:.4f} {dec.launch_info()}", flush=True)
