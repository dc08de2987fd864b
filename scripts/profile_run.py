"""Small fixed run for ncu: a few launches of each headline kernel (no timing claims)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workloads", default="scl8,bp504")
ap.add_argument("--frames", type=int, default=37888)   # 148 SMs x 32 warps x 8 frames
ap.add_argument("--reps", type=int, default=2)
a = ap.parse_args()
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
for name in a.workloads.split(","):
    w = dict(bench.WORKLOADS[name])
    w["frames"] = a.frames
    llr, ref, code = bench.make_inputs(w, torch, dev, 1)
    dec = bench.make_decoder(w, code)
    for _ in range(a.reps):
        bits = dec.decode_batch(llr)
    torch.cuda.synchronize()
    ok = (bits[:, :ref.shape[1]] == ref).all(dim=1).float().mean().item()
    print(name, "frames", a.frames, "frame-ok", ok, dec.launch_info())
