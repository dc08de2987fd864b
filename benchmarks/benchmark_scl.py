"""SC vs SCL benchmark re-pointed at decode_batch (B200).

Mirrors /root/reference/benchmarks/benchmark_scl.py: ber_simulation-style loops (:13-67) and
frame_error_rate_simulation (:70-130) decode the SAME LLRs with SC and SCL L in list_sizes.
Frames are generated on the host in the reference's RNG order (np.random.seed(123 +
int(snr*10)) per SNR, then randint -> encode -> AWGNChannel.transmit per trial, :95-103), so
the inputs are bit-identical to the reference's; each decoder then takes the whole SNR point
in one decode_batch call, and errors are counted on the device (pcl_count_errors).

    python benchmarks/benchmark_scl.py [--N 128 --K 64 --trials 200] [--cpu-check 32]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from polarcode_and_ldpc_b200 import (AWGNChannel, ErrorCounters, PolarEncoder, SCDecoder,  # noqa: E402
                                     SCLDecoder)


def frame_error_rate_simulation(N, K, snr_range, list_sizes, n_trials=500, frozen_bits=None):
    import torch
    print(f"=== FER Simulation: N={N}, K={K}, trials={n_trials} ===\n")
    encoder = PolarEncoder(N, K, frozen_bits=frozen_bits)
    decoders = {"SC": SCDecoder(N, K, frozen_bits=encoder.frozen_bits)}
    for L in list_sizes:
        decoders[f"SCL{L}"] = SCLDecoder(N, K, list_size=L, frozen_bits=encoder.frozen_bits)
    names = list(decoders)
    counters = ErrorCounters(len(snr_range) * len(names), device="cuda")
    for si, snr_db in enumerate(snr_range):
        channel = AWGNChannel(snr_db)
        np.random.seed(123 + int(snr_db * 10))
        msgs, llrs = [], []
        for _ in range(n_trials):
            message = np.random.randint(0, 2, K)
            llrs.append(channel.transmit(encoder.encode(message), return_llr=True))
            msgs.append(message)
        llr_dev = torch.from_numpy(np.array(llrs)).cuda()
        ref_dev = torch.from_numpy(np.array(msgs, dtype=np.uint8)).cuda()
        for di, name in enumerate(names):
            counters.add(si * len(names) + di, decoders[name].decode_batch(llr_dev), ref_dev)
    counters.allreduce()
    ber, fer = counters.rates()
    out = {"snr_db": [float(s) for s in snr_range], "trials": n_trials, "N": N, "K": K}
    for di, name in enumerate(names):
        out[name] = {"fer": [float(fer[si * len(names) + di]) for si in range(len(snr_range))],
                     "ber": [float(ber[si * len(names) + di]) for si in range(len(snr_range))]}
        print(f"  {name:6s} FER: " + " ".join(f"{v:.4f}" for v in out[name]["fer"]))
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--N", type=int, default=128)
    ap.add_argument("--K", type=int, default=64)
    ap.add_argument("--trials", type=int, default=200)
    ap.add_argument("--output", default="results/benchmark_scl.json")
    a = ap.parse_args()
    res = frame_error_rate_simulation(a.N, a.K, np.arange(0.0, 3.5, 0.5), [1, 2, 4, 8], a.trials)
    os.makedirs(os.path.dirname(a.output) or ".", exist_ok=True)
    with open(a.output, "w") as fh:
        json.dump(res, fh, indent=2)
