"""Code length / code rate sweeps re-pointed at decode_batch (B200).

Mirrors /root/reference/benchmarks/test_code_parameters.py: test_code_lengths (:21-146: polar
N in 128..4096, LDPC n in 126..4032 at fixed rate) and test_code_rates (:149-260: N fixed,
rates 1/4..7/8) with the same result dictionaries and the same output file
(results/code_params/code_params_results.json: {'length_tests', 'rate_tests'} per code
type; lists of encoding_time, decoding_time [s / frame], encoding_throughput,
decoding_throughput [Mbps of information bits], ber, fer).  Encoding time is the on-device
generator (message + encode + channel), decoding time the decode kernels (CUDA events).

    python benchmarks/test_code_parameters.py [--num-frames 20000]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import polarcode_and_ldpc_b200 as P  # noqa: E402

KEYS = ("encoding_time", "decoding_time", "encoding_throughput", "decoding_throughput", "ber", "fer")


def _measure(code_type, N, K, snr_db, num_frames):
    if code_type == "polar":
        code = P.make_polar_code(N, K, 2.0)
        decoder = P.SCDecoder(N, K, frozen_bits=code["frozen_bits"])
    else:
        code = P.make_ldpc_code(N, dv=3, dc=6, seed=42)
        decoder = P.BPDecoder(code["H"], max_iter=20)
    P.simulate_point(code, {"d": decoder}, snr_db, min(num_frames, 256), None, seed=1)      # warm-up
    r = P.simulate_point(code, {"d": decoder}, snr_db, num_frames, None, seed=2, first_chunk=num_frames)
    d, frames, bits = r["d"], r["d"]["frames_tested"], r["d"]["total_bits"]
    out = {"encoding_time": r["_gen_seconds"] / frames, "decoding_time": d["decode_seconds"] / frames,
           "encoding_throughput": bits / r["_gen_seconds"] / 1e6, "decoding_throughput": bits / d["decode_seconds"] / 1e6,
           "ber": d["ber"], "fer": d["fer"]}
    print(f"  Encoding: {out['encoding_time'] * 1000:.5f}ms/frame, {out['encoding_throughput']:.1f} Mbps")
    print(f"  Decoding: {out['decoding_time'] * 1000:.5f}ms/frame, {out['decoding_throughput']:.1f} Mbps")
    print(f"  BER: {out['ber']:.6f}, FER: {out['fer']:.4f}")
    return out, code["K"]


def test_code_lengths(code_type="polar", rates=[0.5], snr_db=3.0, num_frames=50):
    code_lengths = [128, 256, 512, 1024, 2048, 4096] if code_type == "polar" else [126, 252, 504, 1008, 2016, 4032]
    results = {"code_lengths": code_lengths, "rates": {}, "snr_db": snr_db, "num_frames": num_frames}
    for rate in rates:
        print(f"\n{'=' * 70}\nTesting {code_type.upper()} - Code Rate: {rate}\n{'=' * 70}")
        rr = {k: [] for k in KEYS}
        for N in code_lengths:
            K = int(N * rate)
            print(f"\nN={N}, K={K}, rate={rate:.3f}")
            try:
                out, _ = _measure(code_type, N, K, snr_db, num_frames)
            except Exception as e:                       # the reference records None and goes on (:134-142)
                print(f"  Error: {e}")
                out = {k: None for k in KEYS}
            for k in KEYS:
                rr[k].append(out[k])
        results["rates"][rate] = rr
    return results


def test_code_rates(code_type="polar", N=1024, snr_db=3.0, num_frames=50):
    rates = [1 / 4, 1 / 3, 2 / 5, 1 / 2, 3 / 5, 2 / 3, 3 / 4, 4 / 5, 5 / 6, 7 / 8]     # :160
    results = {"N": N, "rates": rates, "K_values": [], "snr_db": snr_db, "num_frames": num_frames}
    results.update({k: [] for k in KEYS})
    print(f"\n{'=' * 70}\nTesting {code_type.upper()} - Code Length: N={N}\n{'=' * 70}")
    for rate in rates:
        K = int(N * rate)
        results["K_values"].append(K)
        print(f"\nRate={rate:.3f} (N={N}, K={K})")
        try:
            out, k_actual = _measure(code_type, N, K, snr_db, num_frames)
            results["K_values"][-1] = k_actual
        except Exception as e:
            print(f"  Error: {e}")
            out = {k: None for k in KEYS}
        for k in KEYS:
            results[k].append(out[k])
    return results


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--num-frames", type=int, default=20000)
    ap.add_argument("--snr-db", type=float, default=3.0)
    ap.add_argument("--output-dir", default=str(Path(__file__).parent.parent / "results" / "code_params"))
    a = ap.parse_args()
    results = {"length_tests": {}, "rate_tests": {}}
    for code_type in ("polar", "ldpc"):
        results["length_tests"][code_type] = test_code_lengths(code_type, [0.5], a.snr_db, a.num_frames)
        results["rate_tests"][code_type] = test_code_rates(code_type, 1024 if code_type == "polar" else 1008,
                                                           a.snr_db, a.num_frames)
    out = Path(a.output_dir)
    out.mkdir(parents=True, exist_ok=True)
    with open(out / "code_params_results.json", "w") as f:
        json.dump(results, f, indent=2)
    print(f"\nSaved: {out / 'code_params_results.json'}")


if __name__ == "__main__":
    main()
