"""Kernel-only timing of one bench workload (device-resident input), for experiment builds:
    PCL_LIB=_variants/libpcl_x.so python scripts/time_workload.py scl8 [frames]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "scl8"
w = dict(bench.WORKLOADS[name])
if len(sys.argv) > 2:
    w["frames"] = int(sys.argv[2])
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
llr, ref, code = bench.make_inputs(w, torch, dev, 1)
dec = bench.make_decoder(w, code)
for _ in range(3):
    bits = dec.decode_batch(llr)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
for _ in range(5):
    bits = dec.decode_batch(llr)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
ok = (bits[:, :ref.shape[1]] == ref).all(dim=1).float().mean().item()
kinfo = w.get("K", ref.shape[1])
print(f"{os.environ.get('PCL_LIB', 'default')} {name} {llr.shape[0] * bench.info_bits(w) / ms / 1e6:7.3f} Gbps frame-ok {ok:.5f} "
      f"{dec.launch_info()}", flush=True)
