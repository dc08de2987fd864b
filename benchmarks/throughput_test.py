"""Throughput benchmark re-pointed at decode_batch (B200).

Mirrors /root/reference/benchmarks/throughput_test.py: run_throughput_test (:23),
measure_polar_throughput (:185) and measure_ldpc_throughput (:269) keep their arguments
and result keys (Mbps, seconds), the RNG call order that builds the frames (randint ->
encode -> AWGNChannel.transmit per frame, channel seeded with 42) and the YAML-shaped config
dicts.  The per-frame Python decode loop (:230-235, :317-322) becomes ONE decode_batch call;
`decoding_time` is timed with CUDA events around it (LLRs already on the device) and
`end_to_end_time` includes host encoding, channel and the host<->device copies.

Deviation (documented): the reference's LDPCEncoder falls back to "direct solving" for the
in-repo H and emits invalid codewords, so its BP always runs max_iter
iterations; here codewords are valid (GF(2) null-space generator), so early stop triggers.
Pass early_stop=False to time the reference's worst case.

    python benchmarks/throughput_test.py [--iterations 4096] [--snr 3.0]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
from pathlib import Path
from typing import Dict

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from polarcode_and_ldpc_b200 import (AWGNChannel, BPDecoder, LDPCEncoder, PolarEncoder,  # noqa: E402
                                     SCDecoder, SCLDecoder)

DEFAULT_POLAR = {"encoding": {"N": 1024, "K": 512}, "decoding": {"list_size": 8}}
DEFAULT_LDPC = {"encoding": {"n": 504, "k": 252, "dv": 3, "dc": 6}, "decoding": {"max_iterations": 20}}


def _timed_decode(decoder, llr: np.ndarray, **kw):
    """(seconds on the device for one decode_batch, decoded bits on the host)."""
    import torch
    dev = decoder._to_device(llr)
    decoder.decode_batch(dev, **kw)                      # warm-up (the reference does 10 frames)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    out = decoder.decode_batch(dev, **kw)
    e1.record()
    torch.cuda.synchronize()
    out = out[0] if isinstance(out, tuple) else out
    return e0.elapsed_time(e1) * 1e-3, out.cpu().numpy()


def measure_polar_throughput(config: Dict, num_iterations: int, snr_db: float, list_size: int = 1) -> Dict:
    N, K = config["encoding"]["N"], config["encoding"]["K"]
    print(f"Polar: N={N}, K={K}, rate={K/N:.3f}, list_size={list_size}")
    encoder = PolarEncoder(N, K)
    frozen = encoder.frozen_bits
    decoder = SCDecoder(N, K, frozen_bits=frozen) if list_size == 1 else \
        SCLDecoder(N, K, list_size=list_size, frozen_bits=frozen)
    channel = AWGNChannel(snr_db=snr_db, seed=42)
    messages = np.array([np.random.randint(0, 2, K) for _ in range(num_iterations)])

    t0 = time.time()
    codewords = encoder.encode_batch(messages)
    encoding_time = time.time() - t0
    llr = np.array([channel.transmit(cw, return_llr=True) for cw in codewords])   # reference order
    decoding_time, decoded = _timed_decode(decoder, llr)

    t0 = time.time()
    cw2 = encoder.encode_batch(messages)
    llr2 = channel.transmit_batch(cw2)
    _ = decoder.decode_batch(llr2)
    end_to_end_time = time.time() - t0

    total_bits = num_iterations * K
    res = {
        "N": N, "K": K, "rate": K / N, "num_iterations": num_iterations, "list_size": list_size,
        "encoding_time": encoding_time, "decoding_time": decoding_time, "end_to_end_time": end_to_end_time,
        "encoding_throughput": total_bits / encoding_time / 1e6,
        "decoding_throughput": total_bits / decoding_time / 1e6,
        "end_to_end_throughput": total_bits / end_to_end_time / 1e6,
        "frame_error_rate": float((decoded != messages).any(axis=1).mean()),
    }
    print(f"  Decoding: {decoding_time:.4f}s for {num_iterations} frames -> {res['decoding_throughput']:.2f} Mbps")
    print(f"  End-to-End: {end_to_end_time:.3f}s -> {res['end_to_end_throughput']:.2f} Mbps")
    return res


def measure_ldpc_throughput(config: Dict, num_iterations: int, snr_db: float, early_stop: bool = True) -> Dict:
    enc = config["encoding"]
    n, k, dv, dc = enc["n"], enc["k"], enc.get("dv", 3), enc.get("dc", 6)
    max_iter = config["decoding"].get("max_iterations", 50)
    print(f"LDPC: n={n}, k={k}, rate={k/n:.3f}, dv={dv}, dc={dc}, max_iter={max_iter}")
    encoder = LDPCEncoder(n, k, dv=dv, dc=dc, seed=42)          # in-repo column-random H, like :285
    decoder = BPDecoder(encoder.H, max_iter=max_iter, early_stop=early_stop)
    channel = AWGNChannel(snr_db=snr_db, seed=42)
    messages = np.array([np.random.randint(0, 2, encoder.k) for _ in range(num_iterations)])

    t0 = time.time()
    codewords = encoder.encode_batch(messages)
    encoding_time = time.time() - t0
    llr = np.array([channel.transmit(cw, return_llr=True) for cw in codewords])
    decoding_time, decoded = _timed_decode(decoder, llr, return_iterations=True)

    t0 = time.time()
    _ = decoder.decode_batch(channel.transmit_batch(encoder.encode_batch(messages)))
    end_to_end_time = time.time() - t0

    total_bits = num_iterations * encoder.k
    res = {
        "n": n, "k": encoder.k, "rate": encoder.k / n, "num_iterations": num_iterations, "max_iter": max_iter,
        "encoding_time": encoding_time, "decoding_time": decoding_time, "end_to_end_time": end_to_end_time,
        "encoding_throughput": total_bits / encoding_time / 1e6,
        "decoding_throughput": total_bits / decoding_time / 1e6,
        "end_to_end_throughput": total_bits / end_to_end_time / 1e6,
        "frame_error_rate": float((decoded != codewords).any(axis=1).mean()),
    }
    print(f"  Decoding: {decoding_time:.4f}s for {num_iterations} frames -> {res['decoding_throughput']:.2f} Mbps")
    print(f"  End-to-End: {end_to_end_time:.3f}s -> {res['end_to_end_throughput']:.2f} Mbps")
    return res


def run_throughput_test(polar_config: Dict, ldpc_config: Dict, output_dir: Path, num_iterations: int = 100,
                        snr_db: float = 3.0) -> Dict:
    print(f"\n{'=' * 60}\nThroughput Test (B200, decode_batch)\n{'=' * 60}")
    results = {"num_iterations": num_iterations, "snr_db": snr_db}
    results["polar"] = measure_polar_throughput(polar_config, num_iterations, snr_db)
    L = polar_config.get("decoding", {}).get("list_size", 8)
    results["polar_scl"] = measure_polar_throughput(polar_config, num_iterations, snr_db, list_size=L)
    results["ldpc"] = measure_ldpc_throughput(ldpc_config, num_iterations, snr_db)
    results["ldpc_no_early_stop"] = measure_ldpc_throughput(ldpc_config, num_iterations, snr_db, early_stop=False)
    output_dir = Path(output_dir)
    (output_dir / "data").mkdir(parents=True, exist_ok=True)
    with open(output_dir / "data" / "throughput_results.json", "w") as fh:
        json.dump(results, fh, indent=2)
    return results


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--iterations", type=int, default=4096)
    ap.add_argument("--snr", type=float, default=3.0)
    ap.add_argument("--output", default="results")
    a = ap.parse_args()
    run_throughput_test(DEFAULT_POLAR, DEFAULT_LDPC, Path(a.output), a.iterations, a.snr)
