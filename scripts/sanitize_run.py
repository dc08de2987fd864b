"""Small end-to-end decodes for compute-sanitizer (memcheck / racecheck): every kernel family once."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import polarcode_and_ldpc_b200 as P  # noqa: E402

rng = np.random.default_rng(0)
for N, K, L, F in ((1024, 512, 8, 40), (256, 100, 4, 37), (64, 32, 32, 9), (8, 4, 2, 5), (2048, 1000, 16, 6)):
    fz = P.bhattacharyya_frozen_set(N, K, 2.0)
    llr = rng.normal(1.0, 3.0, size=(F, N))
    for dt in ("float32", "float64"):
        out = P.SCLDecoder(N, K, L, fz, dtype=dt).decode_batch(llr, return_path_metrics=True, return_leaf_llr=True)
        print("scl", N, L, dt, out[0].shape)
    print("sc", P.SCDecoder(N, K, fz).decode_batch(llr).shape)
    print("crc", P.SCLDecoder(N, K, L, fz, use_crc=True).decode_batch(llr).shape)
os.environ["PCL_POLAR_GENERIC"] = "1"
fz = P.bhattacharyya_frozen_set(256, 128, 2.0)
print("generic", P.SCLDecoder(256, 128, 8, fz).decode_batch(rng.normal(1, 3, size=(11, 256))).shape)
os.environ.pop("PCL_POLAR_GENERIC")
for n, F in ((504, 40), (96, 33), (2016, 9)):
    H = P.gallager_parity_check(n, 3, 6, 42)
    llr = rng.normal(1.0, 2.5, size=(F, n))
    for dt in ("float32", "float64"):
        print("bp", n, dt, P.BPDecoder(H, max_iter=8, dtype=dt).decode_batch(llr, return_iterations=True, return_total_llr=True)[1][:4])
        print("ms", n, dt, P.MSDecoder(H, max_iter=8, normalization=0.75, dtype=dt).decode_batch(llr).shape)
Hm = P.mackay_parity_check(120, 60, 3, 6, seed=42)
print("irregular", P.BPDecoder(Hm, max_iter=6).decode_batch(rng.normal(1, 2, size=(21, 120))).shape)
print("sanitize run ok")
