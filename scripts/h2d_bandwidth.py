"""Bare host->device copy bandwidth with all ranks copying at once: the ceiling of the end-to-end
(host-buffer) decode path at N GPUs.  Run under torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        scripts/h2d_bandwidth.py [--out gpurun_out/h2d_N.json]

Rank 0 also records `nvidia-smi topo -m` (which GPUs share a PCIe switch / NUMA node)."""
import argparse
import json
import os
import subprocess

import torch
import torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument("--mib", type=int, default=512)
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--out", default="")
a = ap.parse_args()

rank = int(os.environ.get("RANK", 0))
world = int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
nbytes = a.mib << 20
host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
host.fill_(rank + 1)
dst = torch.empty(nbytes, dtype=torch.uint8, device=dev)
back = torch.empty(nbytes // 64, dtype=torch.uint8).pin_memory()       # the packed result is 1/64 of the LLR bytes
res = {}
for name, with_d2h in (("h2d", False), ("h2d_plus_packed_d2h", True)):
    for _ in range(2):
        dst.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        dst.copy_(host, non_blocking=True)
        if with_d2h:
            back.copy_(dst[: nbytes // 64], non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    allms = [torch.zeros_like(t) for _ in range(world)]
    if world > 1:
        dist.all_gather(allms, t)
    else:
        allms = [t]
    per_rank = [a.reps * nbytes / (float(x.item()) * 1e-3) / 1e9 for x in allms]
    res[name] = {"per_rank_gbs": [round(x, 2) for x in per_rank],
                 "aggregate_gbs": round(world * a.reps * nbytes / (max(float(x.item()) for x in allms) * 1e-3) / 1e9, 2)}
if rank == 0:
    try:
        topo = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=30).stdout
    except Exception as e:  # noqa: BLE001
        topo = f"unavailable: {e}"
    out = {"n_gpus": world, "mib_per_copy": a.mib, "reps": a.reps, **res,
           "scl8_gbps_ceiling_fp32": round(res["h2d"]["aggregate_gbs"] * 512 / 4096 , 2),
           "host_cpus": os.cpu_count(), "topo": topo.splitlines()}
    print(json.dumps(out))
    if a.out:
        os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
        json.dump(out, open(a.out, "w"), indent=1)
if world > 1:
    dist.destroy_process_group()
