"""Frame-sharded multi-GPU BER/FER sweep (BASELINE configs[4]).

The sweep of /root/reference/benchmarks/test_snr_curves.py (simulate_snr_curve :25,
test_multiple_rates :166: Polar SCL and LDPC BP, rates {0.50, 0.67, 0.75, 0.83}, SNR -2..5 dB)
with frames sharded across ranks: rank r decodes frames [r F/R, (r+1) F/R) of every point,
counts errors on the device, and ONE NCCL allreduce of the [points, 4] int64 counter tensor
ends the sweep.  K = int(N * rate) as at :200.  The reference's sequential `max_errors` stop
(:143) has no batch equivalent; the frame count per point is fixed.

    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 benchmarks/snr_sweep.py --frames 1000000
    python benchmarks/snr_sweep.py --frames 20000          # single GPU
Frames are generated on the device by the library's generator (csrc/framegen.cuh: message ->
encode -> BPSK + AWGN, same formulas as src/channel/awgn.py:47,75).  Frame f of a point is a
function of (seed, point, f) only, so the counters do not depend on the number of ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402  (frame generator shared with the headline benchmark)


def main():
    import torch
    import torch.distributed as dist
    import polarcode_and_ldpc_b200 as P
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=20000, help="frames per (code, rate, SNR) point, all ranks")
    ap.add_argument("--chunk", type=int, default=65536)
    ap.add_argument("--list-size", type=int, default=8)
    ap.add_argument("--snrs", default="-2,-1,0,1,2,3,4,5")
    ap.add_argument("--rates", default="0.50,0.67,0.75,0.83")
    ap.add_argument("--seed", type=int, default=20240101)
    ap.add_argument("--output", default="results/snr_sweep.json")
    a = ap.parse_args()
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    snrs = [float(s) for s in a.snrs.split(",")]
    rates = [float(r) for r in a.rates.split(",")]
    N = 1024
    points = [("polar", r, s) for r in rates for s in snrs] + [("ldpc", 0.5, s) for s in snrs]
    counters = P.ErrorCounters(len(points), device=dev)
    lo, hi = P.shard_range(a.frames, rank, world)
    t0 = time.time()
    for pi, (code, rate, snr) in enumerate(points):
        if code == "polar":
            w = dict(kind="polar", N=N, K=int(N * rate), L=a.list_size, snr=snr)
        else:
            w = dict(kind="ldpc", n=1008, k=504, mode="bp", iters=20, snr=snr)
        dec = None
        for c0 in range(lo, hi, a.chunk):
            w["frames"] = min(a.chunk, hi - c0)
            llr, ref, codeinfo = bench.make_inputs(w, torch, dev, seed=a.seed + 7919 * pi, frame0=c0)
            if dec is None:
                if code == "polar":
                    dec = P.SCLDecoder(N, w["K"], list_size=a.list_size, frozen_bits=codeinfo["frozen"])
                else:
                    dec = P.BPDecoder(codeinfo["H"], max_iter=20)
            counters.add(pi, dec.decode_batch(llr), ref)
    counters.allreduce()
    torch.cuda.synchronize()
    ber, fer = counters.rates()
    if rank == 0:
        out = {"frames_per_point": a.frames, "world_size": world, "seed": a.seed, "seconds": time.time() - t0,
               "points": [{"code": c, "rate": r, "snr_db": s, "ber": float(ber[i]), "fer": float(fer[i]),
                           "counters": counters.t[i].tolist()} for i, (c, r, s) in enumerate(points)]}
        os.makedirs(os.path.dirname(a.output) or ".", exist_ok=True)
        with open(a.output, "w") as fh:
            json.dump(out, fh, indent=1)
        for pt in out["points"]:
            print(f"{pt['code']:5s} r={pt['rate']:.2f} {pt['snr_db']:+.1f} dB  BER {pt['ber']:.3e}  FER {pt['fer']:.3e}")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
