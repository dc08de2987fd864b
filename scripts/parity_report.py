"""Bulk parity report on the GPU box: CUDA path (through the C ABI) vs the fp64 oracle on
seeded AWGN frames at BASELINE sizes.  Prints and saves mismatch counts and LLR deviations.

    python scripts/parity_report.py --out gpurun_out/parity_report.json [--frames 8192]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import polarcode_and_ldpc_b200 as P  # noqa: E402
from oracle import oracle  # noqa: E402


def rel_err(got, ref, floor):
    return float(np.max(np.abs(got - ref) / np.maximum(np.abs(ref), floor)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=8192)
    ap.add_argument("--out", default="gpurun_out/parity_report.json")
    a = ap.parse_args()
    nt = os.cpu_count() or 8
    rep = {"frames_per_point": a.frames, "oracle_threads": nt, "cases": []}

    N, K = 1024, 512
    frozen = P.bhattacharyya_frozen_set(N, K, 2.0)
    enc = P.PolarEncoder(N, K, frozen)
    for L in (8, 32):
        d64 = P.SCLDecoder(N, K, list_size=L, frozen_bits=frozen, dtype="float64")
        d32 = P.SCLDecoder(N, K, list_size=L, frozen_bits=frozen, dtype="float32")
        F = a.frames if L == 8 else a.frames // 8
        for snr in (-2.0, -1.0, 0.0, 2.0):
            rng = np.random.default_rng(int(snr * 10) + 1000 + L)
            msg = rng.integers(0, 2, size=(F, K))
            np.random.seed(int(snr * 10) + 2000 + L)
            llr = P.AWGNChannel(snr).transmit_batch(enc.encode_batch(msg))
            t0 = time.time()
            ref, rpm = oracle.polar_scl(N, L, frozen, llr, want_pm=True, nthreads=nt)
            t_or = time.time() - t0
            b64, p64 = d64.decode_batch(llr, return_path_metrics=True)
            b32, p32 = d32.decode_batch(llr, return_path_metrics=True)
            fin = np.isfinite(rpm)
            rep["cases"].append({
                "case": f"polar SCL L={L} N={N} K={K}", "snr_db": snr, "frames": F,
                "fer_reference": float((ref != msg).any(axis=1).mean()),
                "fp64_frames_differing": int((b64 != ref).any(axis=1).sum()),
                "fp32_frames_differing": int((b32 != ref).any(axis=1).sum()),
                "fp64_metric_max_rel_err": rel_err(p64[fin], rpm[fin], 1.0),
                "fp32_metric_max_rel_err": rel_err(p32[fin], rpm[fin], float(np.mean(np.abs(llr)))),
                "oracle_seconds": t_or})
            print(rep["cases"][-1], flush=True)

    for mode, n in (("bp", 504), ("ms", 2016)):
        H = P.gallager_parity_check(n, 3, 6, 42)
        encl = P.LDPCEncoder(n, n // 2, H=H)
        F = a.frames if mode == "bp" else a.frames // 4
        for snr, es in ((-1.0, True), (0.0, True), (1.0, True), (0.0, False)):
            rng = np.random.default_rng(int(snr * 10) + 77)
            cw = encl.encode_batch(rng.integers(0, 2, size=(F, encl.k)))
            np.random.seed(int(snr * 10) + 78)
            llr = P.AWGNChannel(snr).transmit_batch(cw)
            kw = dict(max_iter=20, early_stop=es)
            if mode == "ms":
                kw["normalization"] = 0.75
            rb, ri, rt = oracle.ldpc(H, llr, mode, want_total=True, nthreads=nt, **kw)
            cls = P.BPDecoder if mode == "bp" else P.MSDecoder
            b64, i64, t64 = cls(H, dtype="float64", **kw).decode_batch(llr, return_iterations=True, return_total_llr=True)
            b32, i32, t32 = cls(H, dtype="float32", **kw).decode_batch(llr, return_iterations=True, return_total_llr=True)
            same32 = (b32 == rb).all(axis=1) & (i32 == ri)
            floor = float(np.mean(np.abs(llr)))
            rep["cases"].append({
                "case": f"ldpc {mode} n={n}", "snr_db": snr, "early_stop": es, "frames": F,
                "fer_reference": float((rb != cw).any(axis=1).mean()), "mean_iterations": float(ri.mean()),
                "fp64_frames_differing": int(((b64 != rb).any(axis=1) | (i64 != ri)).sum()),
                "fp32_frames_differing": int((~same32).sum()),
                "fp64_total_llr_max_rel_err": rel_err(t64, rt, floor),
                "fp32_total_llr_max_rel_err_same_frames": rel_err(t32[same32], rt[same32], floor),
                "fp32_total_llr_rel_err_p9999": float(np.quantile(
                    np.abs(t32[same32] - rt[same32]) / np.maximum(np.abs(rt[same32]), floor), 0.9999))})
            print(rep["cases"][-1], flush=True)
    tot = sum(c["frames"] for c in rep["cases"])
    rep["total_frames"] = tot
    rep["fp64_frames_differing_total"] = sum(c["fp64_frames_differing"] for c in rep["cases"])
    rep["fp32_frames_differing_total"] = sum(c["fp32_frames_differing"] for c in rep["cases"])
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    with open(a.out, "w") as fh:
        json.dump(rep, fh, indent=1)
    print("TOTAL", tot, "fp64 diff", rep["fp64_frames_differing_total"], "fp32 diff", rep["fp32_frames_differing_total"])


if __name__ == "__main__":
    main()
