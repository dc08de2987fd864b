"""SNR performance curves re-pointed at decode_batch (B200).

Mirrors /root/reference/benchmarks/test_snr_curves.py: simulate_snr_curve (:25-163),
test_multiple_rates (:166-240), analyze_snr_requirements (:357-408) and main (:411-505) keep
their names, arguments, printed lines and the JSON files they write
(results/snr_curves/{polar_results,ldpc_results,snr_analysis}.json; schema
docs/SNR_CURVES_TEST_SUMMARY.md:247-270).  The per-frame host loop becomes
polarcode_and_ldpc_b200.simulate_point (device frame generation, one decode_batch per chunk,
device error counting, `max_errors` applied between chunks).  Differences, all deliberate:
the third-party-library arm is not run (`library` stays None: polarcodes / pyldpc are not
used), plots are not drawn (out of scope), and --decoder picks SC (the reference) or SCL.

    python benchmarks/test_snr_curves.py --num-frames 100000 [--decoder scl --list-size 8]
    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 benchmarks/test_snr_curves.py --num-frames 1000000
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
from pathlib import Path
from typing import Dict, List, Tuple

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import polarcode_and_ldpc_b200 as P  # noqa: E402

POLAR_DECODER = {"kind": "sc", "list_size": 8}


def simulate_snr_curve(code_type: str, N: int, K: int, snr_range: np.ndarray, num_frames: int = 100,
                       max_errors: int = 100, use_library: bool = False) -> Tuple[List[float], List[float], Dict]:
    assert not use_library, "the third-party library arm is not part of this port"
    print(f"\n{'=' * 70}")
    print(f"Testing {code_type.upper()}: N={N}, K={K}, rate={K / N:.3f}")
    print(f"Implementation: Self (B200 decode_batch)")
    print(f"{'=' * 70}")
    if code_type == "polar":
        code = P.make_polar_code(N, K, 2.0)
        if POLAR_DECODER["kind"] == "sc":
            decoder = P.SCDecoder(N, K, frozen_bits=code["frozen_bits"])
        else:
            decoder = P.SCLDecoder(N, K, list_size=POLAR_DECODER["list_size"], frozen_bits=code["frozen_bits"])
    else:
        code = P.make_ldpc_code(N, dv=3, dc=6, seed=42)
        K = code["K"]                                     # the code's true dimension (:68-69)
        decoder = P.BPDecoder(code["H"], max_iter=20)
    ber_list, fer_list = [], []
    stats = {"frames_tested": [], "total_bits": [], "error_bits": [], "frame_errors": [], "simulation_time": []}
    for si, snr_db in enumerate(snr_range):
        print(f"\nSNR = {snr_db:.1f} dB", end=" ")
        t0 = time.time()
        r = P.simulate_point(code, {"d": decoder}, float(snr_db), num_frames, max_errors, seed=1000 + si)["d"]
        elapsed = time.time() - t0
        ber_list.append(r["ber"])
        fer_list.append(r["fer"])
        for k in ("frames_tested", "total_bits", "error_bits", "frame_errors"):
            stats[k].append(r[k])
        stats["simulation_time"].append(elapsed)
        print(f"-> BER: {r['ber']:.6f}, FER: {r['fer']:.4f} ({r['frames_tested']} frames, {elapsed:.1f}s)")
    return ber_list, fer_list, stats


def test_multiple_rates(code_type: str, N_base: int, rates: List[float], snr_range: np.ndarray,
                        num_frames: int = 100, max_errors: int = 100, test_library: bool = False) -> Dict:
    results = {"code_type": code_type, "N": N_base, "rates": rates, "snr_range": snr_range.tolist(),
               "self": {}, "library": None}
    for rate in rates:
        K = int(N_base * rate)
        print(f"\n{'#' * 70}")
        print(f"Rate = {rate:.3f} (K={K}) - Self Implementation")
        print(f"{'#' * 70}")
        ber, fer, stats = simulate_snr_curve(code_type, N_base, K, snr_range, num_frames, max_errors)
        results["self"][rate] = {"K": K, "ber": ber, "fer": fer, "stats": stats}
    return results


def analyze_snr_requirements(results_polar: Dict, results_ldpc: Dict, target_ber: float = 1e-3) -> Dict:
    snr_range = np.array(results_polar["snr_range"])
    analysis = {"target_ber": target_ber, "polar": {}, "ldpc": {}, "snr_gap": {}}
    print(f"\n{'=' * 70}")
    print(f"SNR Requirements Analysis (Target BER = {target_ber:.0e})")
    print(f"{'=' * 70}")
    print(f"{'Rate':<8} {'Polar SNR':<12} {'LDPC SNR':<12} {'Gap (dB)':<12}")
    print(f"{'-' * 70}")
    for rate in results_polar["rates"]:
        found = {}
        for name, res in (("polar", results_polar), ("ldpc", results_ldpc)):
            idx = np.where(np.array(res["self"][rate]["ber"]) < target_ber)[0]
            found[name] = float(snr_range[idx[0]]) if len(idx) else None
            analysis[name][rate] = found[name]
        gap = found["polar"] - found["ldpc"] if None not in found.values() else None
        analysis["snr_gap"][rate] = gap
        fmt = lambda v: f"{v:.1f}" if v is not None else ">" + str(snr_range[-1])  # noqa: E731
        print(f"{rate:<8.2f} {fmt(found['polar']):<12} {fmt(found['ldpc']):<12} {('%.2f' % gap) if gap is not None else 'N/A':<12}")
    return analysis


def main():
    import torch
    import torch.distributed as dist
    ap = argparse.ArgumentParser()
    ap.add_argument("--num-frames", type=int, default=100)
    ap.add_argument("--max-errors", type=int, default=100)
    ap.add_argument("--rates", default="0.50,0.67,0.75,0.83")
    ap.add_argument("--snr", default="-2,6,1", help="start,stop,step of np.arange")
    ap.add_argument("--decoder", choices=["sc", "scl"], default="sc")
    ap.add_argument("--list-size", type=int, default=8)
    ap.add_argument("--output-dir", default=str(Path(__file__).parent.parent / "results" / "snr_curves"))
    a = ap.parse_args()
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
        if rank:
            sys.stdout = open(os.devnull, "w")
    POLAR_DECODER.update(kind=a.decoder, list_size=a.list_size)
    s0, s1, ds = (float(x) for x in a.snr.split(","))
    config = {"polar_N": 1024, "ldpc_N": 1008, "rates": [float(r) for r in a.rates.split(",")],
              "snr_range": np.arange(s0, s1, ds), "num_frames": a.num_frames, "max_errors": a.max_errors,
              "output_dir": Path(a.output_dir)}
    print(f"\n{'#' * 70}\nSNR Performance Curve Testing\n{'#' * 70}")
    print("Configuration:")
    print(f"  Polar N: {config['polar_N']}\n  LDPC N:  {config['ldpc_N']}\n  Rates:   {config['rates']}")
    print(f"  SNR:     {config['snr_range'][0]:.1f} to {config['snr_range'][-1]:.1f} dB")
    print(f"  Frames:  {config['num_frames']} (max), stop at {config['max_errors']} errors")
    results_polar = test_multiple_rates("polar", config["polar_N"], config["rates"], config["snr_range"],
                                        config["num_frames"], config["max_errors"])
    results_ldpc = test_multiple_rates("ldpc", config["ldpc_N"], config["rates"], config["snr_range"],
                                       config["num_frames"], config["max_errors"])
    if rank == 0:
        out = config["output_dir"]
        out.mkdir(parents=True, exist_ok=True)
        for name, res in (("polar_results.json", results_polar), ("ldpc_results.json", results_ldpc)):
            with open(out / name, "w") as f:
                json.dump(res, f, indent=2)
            print(f"\nSaved: {out / name}")
        analysis = {"ber_1e-3": analyze_snr_requirements(results_polar, results_ldpc, 1e-3),
                    "ber_1e-5": analyze_snr_requirements(results_polar, results_ldpc, 1e-5)}
        with open(out / "snr_analysis.json", "w") as f:
            json.dump(analysis, f, indent=2)
        print(f"\nSaved: {out / 'snr_analysis.json'}")
        print(f"\n{'=' * 70}\nTesting Complete!\n{'=' * 70}\nResults saved to: {out}")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
