"""SC vs SCL comparison re-pointed at decode_batch (B200).

Mirrors simulate_sc_vs_scl of /root/reference/benchmarks/sc_vs_scl.py (:203-349): SC and SCL
with every list size decode the SAME frames; the results dictionary keeps its layout
({'N', 'K', 'rate', 'snr_db', 'sc': {ber, fer, time}, 'scl': {L: {ber, fer, time}}}, time in
ms per frame) and is saved as results/sc_vs_scl/results.json (:622).  The stop rule is the
reference's: a point ends when SC and every SCL have reached `max_errors` frame errors (:317).
The quick demo and the plots (:37-200, :352-719) are not reproduced.

    python benchmarks/sc_vs_scl.py [--N 1024 --K 512 --num-frames 100000 --list-sizes 1,2,4,8,16,32]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
from pathlib import Path
from typing import Dict, List

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import polarcode_and_ldpc_b200 as P  # noqa: E402


def simulate_sc_vs_scl(N: int, K: int, snr_range: np.ndarray, list_sizes: List[int] = [1, 2, 4, 8, 16],
                       num_frames: int = 100, max_errors: int = 100, seed: int = 42) -> Dict:
    print(f"\n{'=' * 70}\nSC vs SCL: N={N}, K={K}, rate={K / N:.3f}\n{'=' * 70}")
    print(f"SNR range:   {snr_range[0]:.1f} to {snr_range[-1]:.1f} dB\nList sizes:  {list_sizes}")
    print(f"Max frames per SNR point: {num_frames}\nMax errors:  {max_errors}")
    code = P.make_polar_code(N, K, 2.0)
    decoders = {"sc": P.SCDecoder(N, K, frozen_bits=code["frozen_bits"])}
    for L in list_sizes:
        decoders[f"scl{L}"] = P.SCLDecoder(N, K, list_size=L, frozen_bits=code["frozen_bits"])
    results = {"N": N, "K": K, "rate": K / N, "snr_db": snr_range.tolist(),
               "sc": {"ber": [], "fer": [], "time": []},
               "scl": {L: {"ber": [], "fer": [], "time": []} for L in list_sizes}}
    for si, snr_db in enumerate(snr_range):
        r = P.simulate_point(code, decoders, float(snr_db), num_frames, max_errors, seed=seed + si)
        frames = r["sc"]["frames_tested"]
        print(f"\nSNR = {snr_db} dB - {frames} frames | SC: BER={r['sc']['ber']:.2e}, FER={r['sc']['fer']:.4f}")
        for key, dst in [("sc", results["sc"])] + [(f"scl{L}", results["scl"][L]) for L in list_sizes]:
            dst["ber"].append(r[key]["ber"])
            dst["fer"].append(r[key]["fer"])
            dst["time"].append(r[key]["decode_seconds"] / frames * 1000)
        for L in list_sizes:
            print(f"         SCL(L={L:2d}): BER={r[f'scl{L}']['ber']:.2e}, FER={r[f'scl{L}']['fer']:.4f}, "
                  f"{results['scl'][L]['time'][-1] * 1e3:.3f} us/frame")
    return results


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--N", type=int, default=128)
    ap.add_argument("--K", type=int, default=64)
    ap.add_argument("--snr", default="0,4.5,0.5")
    ap.add_argument("--list-sizes", default="1,2,4,8,16")
    ap.add_argument("--num-frames", type=int, default=100000)
    ap.add_argument("--max-errors", type=int, default=1000)
    ap.add_argument("--output-dir", default=str(Path(__file__).parent.parent / "results" / "sc_vs_scl"))
    a = ap.parse_args()
    s0, s1, ds = (float(x) for x in a.snr.split(","))
    res = simulate_sc_vs_scl(a.N, a.K, np.arange(s0, s1, ds), [int(x) for x in a.list_sizes.split(",")],
                             a.num_frames, a.max_errors)
    out = Path(a.output_dir)
    out.mkdir(parents=True, exist_ok=True)
    with open(out / "results.json", "w") as f:
        json.dump(res, f, indent=2)
    print(f"\nSaved: {out / 'results.json'}")
