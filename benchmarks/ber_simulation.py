"""BER / FER simulation re-pointed at decode_batch (B200).

Mirrors /root/reference/benchmarks/ber_simulation.py: run_ber_simulation (:24-129),
simulate_polar (:132-205: PolarEncoder + SCDecoder, frozen set of design SNR from the
config) and simulate_ldpc (:208-293: LDPCEncoder + BPDecoder with the config's max
iterations) with the same config dictionaries and the same results file
(results/data/ber_simulation_results.json: {'snr_db', 'polar': {'self': {ber, fer}},
'ldpc': {'self': {ber, fer}}}).  The per-frame loop is simulate_point (device generator,
decode_batch, device counters).  The third-party arm and the plots are not reproduced.

    python benchmarks/ber_simulation.py [--num-frames 100000 --polar-N 1024 --polar-K 512]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
from pathlib import Path
from typing import Dict, Tuple

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import polarcode_and_ldpc_b200 as P  # noqa: E402


def _curve(code, decoder, snr_db_range, num_frames, max_errors, seed0):
    ber_list, fer_list = [], []
    for si, snr_db in enumerate(snr_db_range):
        t0 = time.time()
        r = P.simulate_point(code, {"d": decoder}, float(snr_db), num_frames, max_errors, seed=seed0 + si)["d"]
        ber_list.append(r["ber"])
        fer_list.append(r["fer"])
        print(f"  SNR={snr_db:4.1f}dB: BER={r['ber']:.6f}, FER={r['fer']:.4f}, "
              f"Frames={r['frames_tested']}, Time={time.time() - t0:.2f}s")
    return np.array(ber_list), np.array(fer_list)


def simulate_polar(snr_db_range: np.ndarray, num_frames: int, max_errors: int, config: Dict) -> Tuple[np.ndarray, np.ndarray]:
    N, K = config["encoding"]["N"], config["encoding"]["K"]
    print(f"Polar: N={N}, K={K}, rate={K / N:.3f}")
    code = P.make_polar_code(N, K, 2.0)                   # PolarLibWrapper(N, K, 2.0), :146
    return _curve(code, P.SCDecoder(N, K, frozen_bits=code["frozen_bits"]), snr_db_range, num_frames, max_errors, 100)


def simulate_ldpc(snr_db_range: np.ndarray, num_frames: int, max_errors: int, config: Dict) -> Tuple[np.ndarray, np.ndarray]:
    n, k = config["encoding"]["n"], config["encoding"]["k"]
    cons = config.get("construction", config["encoding"])    # :218-219 reads config['construction']
    dv, dc = cons.get("dv", 3), cons.get("dc", 6)
    max_iter = config.get("decoding", {}).get("max_iterations", 50)
    code = P.make_ldpc_code(n, dv=dv, dc=dc, seed=42)
    print(f"LDPC: n={n}, k={code['K']} (requested {k}), dv={dv}, dc={dc}, max_iter={max_iter}")
    return _curve(code, P.BPDecoder(code["H"], max_iter=max_iter), snr_db_range, num_frames, max_errors, 200)


def run_ber_simulation(snr_db_range: np.ndarray, num_frames: int, max_errors: int, polar_config: Dict,
                       ldpc_config: Dict, output_dir: Path, use_third_party: bool = False) -> Dict:
    print("=" * 60 + "\nBER/FER Simulation\n" + "=" * 60)
    results = {"snr_db": snr_db_range.tolist(), "polar": {}, "ldpc": {}}
    print(f"\n{'-' * 60}\nTesting Polar Code (Self-Implementation)\n{'-' * 60}")
    ber, fer = simulate_polar(snr_db_range, num_frames, max_errors, polar_config)
    results["polar"]["self"] = {"ber": ber.tolist(), "fer": fer.tolist()}
    print(f"\n{'-' * 60}\nTesting LDPC (Self-Implementation)\n{'-' * 60}")
    ber, fer = simulate_ldpc(snr_db_range, num_frames, max_errors, ldpc_config)
    results["ldpc"]["self"] = {"ber": ber.tolist(), "fer": fer.tolist()}
    path = Path(output_dir) / "data" / "ber_simulation_results.json"
    path.parent.mkdir(parents=True, exist_ok=True)
    with open(path, "w") as f:
        json.dump(results, f, indent=2)
    print(f"\nResults saved to: {path}")
    return results


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--num-frames", type=int, default=10000)
    ap.add_argument("--max-errors", type=int, default=100)
    ap.add_argument("--polar-N", type=int, default=128)
    ap.add_argument("--polar-K", type=int, default=64)
    ap.add_argument("--ldpc-n", type=int, default=120)
    ap.add_argument("--ldpc-iters", type=int, default=50)
    ap.add_argument("--snr", default="0,6,1")
    ap.add_argument("--output-dir", default=str(Path(__file__).parent.parent / "results"))
    a = ap.parse_args()
    s0, s1, ds = (float(x) for x in a.snr.split(","))
    polar_config = {"encoding": {"N": a.polar_N, "K": a.polar_K}, "construction": {"design_snr_db": 2.0}}
    ldpc_config = {"encoding": {"n": a.ldpc_n, "k": a.ldpc_n // 2, "dv": 3, "dc": 6}, "decoding": {"max_iterations": a.ldpc_iters}}
    run_ber_simulation(np.arange(s0, s1, ds), a.num_frames, a.max_errors, polar_config, ldpc_config, Path(a.output_dir))
