import sys, os, numpy as np
sys.path.insert(0, os.getcwd())
import polarcode_and_ldpc_b200 as P
from oracle import oracle
n=504
H = P.gallager_parity_check(n, 3, 6, 42)
enc = P.LDPCEncoder(n, n // 2, H=H)
rng = np.random.default_rng(3)
for snr, es in ((-1.0, True), (1.0, True), (0.0, False), (3.0, False), (-2.0, True)):
    F = 4096
    cw = enc.encode_batch(rng.integers(0, 2, size=(F, enc.k)))
    np.random.seed(int(snr * 10) + 77)
    llr = P.AWGNChannel(snr).transmit_batch(cw)
    kw = dict(max_iter=20, early_stop=es)
    rb, ri, rt = oracle.ldpc(H, llr, "bp", want_total=True, nthreads=16, **kw)
    b32, i32, t32 = P.BPDecoder(H, dtype="float32", **kw).decode_batch(llr, return_iterations=True, return_total_llr=True)
    same = (b32 == rb).all(axis=1) & (i32 == ri)
    floor = float(np.mean(np.abs(llr)))
    err = np.abs(t32 - rt) / np.maximum(np.abs(rt), floor)
    es_ = err[same]
    fm = err.max(axis=1)
    print(f"snr {snr} es {es}: bad frames {int((~same).sum())}/{F}  rel err (same frames): max {es_.max():.3e} p99.99 {np.quantile(es_, 0.9999):.3e} p99 {np.quantile(es_, 0.99):.3e} median {np.median(es_):.3e}; frames with err>1e-4: {int((fm[same] > 1e-4).sum())}; iters of those: {ri[same][fm[same] > 1e-4][:10]}", flush=True)
