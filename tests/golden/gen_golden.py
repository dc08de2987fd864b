"""Generate the committed golden vectors by RUNNING THE REFERENCE ITSELF.

Run in the build container only (needs /root/reference):

    python tests/golden/gen_golden.py

It imports the reference's own decoders / encoders / channel from
/root/reference/src (the `sys.path` style its benchmarks use, e.g.
benchmarks/throughput_test.py:15-19), decodes seeded inputs and stores inputs and
outputs as small .npz fixtures next to this file.  Nothing here is imported at
test time; tests only read the .npz files.

Fixtures
  polar_sc.npz   SCDecoder  (src/polar/decoder.py:12)   bits, leaf LLRs dec.L[:, n]
  polar_scl.npz  SCLDecoder (src/polar/decoder.py:176)  bits, path_metrics, leaf LLRs
  ldpc.npz       BPDecoder / MSDecoder (src/ldpc/decoder.py:11,208) bits, iterations, totals
  doc_kat.npz    docs/SCL_DECODER_README.md:115-128 flow (tests/test_scl_decoder.py:13-48)
  wide.npz       what the reference accepts beyond a warp's width: SCLDecoder with list_size > 32
                 (decoder.py:194-196) and a BP check of degree > 32 (`python gen_golden.py wide`)
"""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, "/root/reference/src")

from polar.decoder import SCDecoder, SCLDecoder          # noqa: E402  (reference)
from polar.encoder import PolarEncoder                    # noqa: E402
from ldpc.decoder import BPDecoder, MSDecoder             # noqa: E402
from ldpc.matrix import mackay_construction               # noqa: E402
from channel.awgn import AWGNChannel                      # noqa: E402


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


pconstr = _load(os.path.join(REPO, "polarcode_and_ldpc_b200/polar/construction.py"), "pconstr")
lconstr = _load(os.path.join(REPO, "polarcode_and_ldpc_b200/ldpc/construction.py"), "lconstr")


def polar_inputs(rng, N, K, frozen, F, snr_db, kind):
    """LLR batches: 'awgn' = encode+AWGN through the reference's encoder/channel,
    'int' = small integers (forces exact metric ties), 'raw' = arbitrary reals."""
    if kind == "awgn":
        enc = PolarEncoder(N, K, frozen_bits=frozen)
        ch = AWGNChannel(snr_db)
        out = []
        for _ in range(F):
            msg = np.random.randint(0, 2, K)
            out.append(ch.transmit(enc.encode(msg)))
        return np.array(out)
    if kind == "int":
        return rng.integers(-3, 4, size=(F, N)).astype(np.float64)
    return rng.normal(0.5, 3.0, size=(F, N))


def gen_polar():
    rng = np.random.default_rng(2024)
    sc, scl = {}, {}
    sc_cases = [  # N, K, frozen kind, F, snr, llr kind
        (8, 4, "default", 6, 2.0, "awgn"), (16, 8, "default", 6, 2.0, "int"),
        (64, 32, "random", 8, 1.0, "awgn"), (64, 20, "random", 6, 0.0, "raw"),
        (256, 128, "bhatt", 8, 3.0, "awgn"), (256, 128, "bhatt", 4, 0.0, "int"),
        (1024, 512, "bhatt", 3, 2.0, "awgn"), (2, 1, "default", 4, 0.0, "raw"),
        (4, 3, "random", 4, 0.0, "int"),
    ]
    np.random.seed(777)
    for ci, (N, K, fk, F, snr, kind) in enumerate(sc_cases):
        if fk == "default":
            frozen = SCDecoder(N, K).frozen_bits
        elif fk == "random":
            frozen = np.sort(rng.choice(N, N - K, replace=False))
        else:
            frozen = pconstr.bhattacharyya_frozen_set(N, K, 2.0)
        llr = polar_inputs(rng, N, K, frozen, F, snr, kind)
        dec = SCDecoder(N, K, frozen_bits=frozen)
        bits, leaf = [], []
        for f in range(F):
            bits.append(dec.decode(llr[f]))
            leaf.append(dec.L[:, dec.n].copy())
        sc[f"c{ci}_N"] = N
        sc[f"c{ci}_frozen"] = np.asarray(frozen, dtype=np.int64)
        sc[f"c{ci}_llr"] = llr
        sc[f"c{ci}_bits"] = np.array(bits, dtype=np.int64)
        sc[f"c{ci}_leaf"] = np.array(leaf)
    sc["ncases"] = len(sc_cases)

    scl_cases = [  # N, K, L, frozen kind, F, snr, kind
        (16, 8, 1, "default", 4, 2.0, "awgn"), (16, 8, 2, "default", 4, 2.0, "int"),
        (16, 8, 4, "random", 6, 1.0, "int"), (16, 8, 8, "random", 6, 1.0, "awgn"),
        (64, 32, 3, "random", 6, 1.0, "awgn"), (64, 32, 8, "random", 8, 0.0, "int"),
        (64, 40, 16, "random", 4, 1.0, "raw"), (64, 32, 32, "random", 4, 0.0, "awgn"),
        (128, 64, 8, "bhatt", 8, 0.0, "awgn"), (128, 64, 5, "bhatt", 4, 0.0, "int"),
        (256, 128, 8, "bhatt", 6, 0.0, "awgn"), (256, 171, 4, "bhatt", 4, 1.0, "awgn"),
        (1024, 512, 8, "bhatt", 4, -1.0, "awgn"), (1024, 849, 32, "bhatt", 1, 3.0, "awgn"),
        (8, 7, 4, "random", 6, 0.0, "int"), (32, 31, 8, "random", 4, 0.0, "int"),
        (64, 32, 8, "randfz_tail", 6, 0.0, "awgn"),
    ]
    for ci, (N, K, L, fk, F, snr, kind) in enumerate(scl_cases):
        if fk == "default":
            frozen = SCDecoder(N, K).frozen_bits
        elif fk == "random":
            frozen = np.sort(rng.choice(N, N - K, replace=False))
        elif fk == "randfz_tail":
            # frozen leaves after the last info leaf -> argmax may not be slot 0
            frozen = np.sort(rng.choice(N - 1, N - K - 1, replace=False))
            frozen = np.sort(np.concatenate([frozen, [N - 1]]))
        else:
            frozen = pconstr.bhattacharyya_frozen_set(N, K, 2.0)
        llr = polar_inputs(rng, N, K, frozen, F, snr, kind)
        dec = SCLDecoder(N, K, list_size=L, frozen_bits=frozen)
        bits, pm, leaf = [], [], []
        for f in range(F):
            bits.append(dec.decode(llr[f]))
            pm.append(dec.path_metrics.copy())
            best = int(np.argmax(dec.path_metrics))
            leaf.append(dec.L_paths[best, :, dec.n].copy())
        scl[f"c{ci}_N"] = N
        scl[f"c{ci}_L"] = L
        scl[f"c{ci}_frozen"] = np.asarray(frozen, dtype=np.int64)
        scl[f"c{ci}_llr"] = llr
        scl[f"c{ci}_bits"] = np.array(bits, dtype=np.int64)
        scl[f"c{ci}_pm"] = np.array(pm)
        scl[f"c{ci}_leaf"] = np.array(leaf)
        print("scl case", ci, N, K, L, "done", flush=True)
    scl["ncases"] = len(scl_cases)
    np.savez_compressed(os.path.join(HERE, "polar_sc.npz"), **sc)
    np.savez_compressed(os.path.join(HERE, "polar_scl.npz"), **scl)


class _BPTotals(BPDecoder):
    """Records total LLRs: _variable_node_update returns them (ldpc/decoder.py:122)."""

    def decode(self, llr, return_iterations=False):
        self._tot = {}
        return super().decode(llr, return_iterations)

    def _variable_node_update(self, llr_channel, messages_in):
        out, total = super()._variable_node_update(llr_channel, messages_in)
        self._cur = getattr(self, "_cur", [])
        self._cur.append(total)
        if len(self._cur) == self.n:
            self._last_totals = np.array(self._cur)
            self._cur = []
        return out, total


def gen_ldpc():
    rng = np.random.default_rng(99)
    out = {}
    H96 = lconstr.gallager_parity_check(96, 3, 6, 42)
    H504 = lconstr.gallager_parity_check(504, 3, 6, 42)
    Hirr = mackay_construction(120, 60, 3, 6, seed=42)
    Hirr504 = mackay_construction(504, 252, 3, 6, seed=42)
    H12 = lconstr.gallager_parity_check(24, 3, 4, 7)
    # a matrix with a degree-1 check, a degree-0 check and a degree-9 variable
    Hodd = np.zeros((12, 20), dtype=int)
    r2 = np.random.RandomState(5)
    for c in range(12):
        Hodd[c, r2.choice(20, r2.randint(2, 7), replace=False)] = 1
    Hodd[3, :] = 0
    Hodd[3, 4] = 1          # degree-1 check
    Hodd[7, :] = 0          # degree-0 check
    Hodd[[0, 1, 2, 4, 5, 6, 8, 9, 10], 11] = 1   # degree-9 variable (numpy pairwise sum path)
    cases = [  # name, H, mode, max_iter, norm, early_stop, F, snr (None = raw llr)
        ("g96_bp", H96, "bp", 20, 1.0, True, 10, 1.0), ("g96_bp_raw", H96, "bp", 12, 1.0, True, 6, None),
        ("g96_bp_nostop", H96, "bp", 8, 1.0, False, 4, 0.0),
        ("g96_ms", H96, "ms", 20, 1.0, True, 8, 1.0), ("g96_ms75", H96, "ms", 20, 0.75, True, 8, 0.0),
        ("g96_ms_nostop", H96, "ms", 6, 0.75, False, 4, 0.0),
        ("g504_bp", H504, "bp", 20, 1.0, True, 4, -1.0), ("g504_ms", H504, "ms", 20, 0.75, True, 3, 0.0),
        ("irr120_bp", Hirr, "bp", 10, 1.0, True, 6, 2.0), ("irr504_bp", Hirr504, "bp", 5, 1.0, True, 2, 3.0),
        ("g24_bp", H12, "bp", 30, 1.0, True, 8, 2.0), ("odd_bp", Hodd, "bp", 10, 1.0, True, 6, None),
        ("g24_ms_int", H12, "ms", 10, 1.0, True, 8, "int"),
    ]
    np.random.seed(4242)
    names = []
    for name, H, mode, it, norm, es, F, snr in cases:
        m, n = H.shape
        if snr is None:
            llr = rng.normal(0.8, 2.5, size=(F, n))
        elif snr == "int":
            llr = rng.integers(-3, 5, size=(F, n)).astype(np.float64)
        else:
            ch = AWGNChannel(snr)
            llr = np.array([ch.transmit(np.zeros(n, dtype=int)) for _ in range(F)])
        bits, iters, tots = [], [], []
        if mode == "bp":
            dec = _BPTotals(H, max_iter=it, early_stop=es)
            for f in range(F):
                b, k = dec.decode(llr[f], return_iterations=True)
                bits.append(b); iters.append(k); tots.append(dec._last_totals.copy())
            out[name + "_total"] = np.array(tots)
        else:
            dec = MSDecoder(H, max_iter=it, normalization=norm, early_stop=es)
            for f in range(F):
                bits.append(dec.decode(llr[f]))
        out[name + "_H"] = H.astype(np.uint8)
        out[name + "_cfg"] = np.array([0 if mode == "bp" else 1, it, int(es)], dtype=np.int64)
        out[name + "_norm"] = np.float64(norm)
        out[name + "_llr"] = llr
        out[name + "_bits"] = np.array(bits, dtype=np.int64)
        if iters:
            out[name + "_iters"] = np.array(iters, dtype=np.int64)
        names.append(name)
        print("ldpc case", name, "done", iters, flush=True)
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "ldpc.npz"), **out)


def gen_doc_kat():
    """tests/test_scl_decoder.py:13-48 with the global RNG as the doc example ran it."""
    N, K = 16, 8
    enc = PolarEncoder(N, K)
    np.random.seed(42)
    message = np.random.randint(0, 2, K)
    codeword = enc.encode(message)
    llr = AWGNChannel(2.0).transmit(codeword, return_llr=True)
    dec = {}
    for L in (1, 2, 4, 8):
        dec[L] = SCLDecoder(N, K, list_size=L, frozen_bits=enc.frozen_bits).decode(llr)
    np.savez_compressed(os.path.join(HERE, "doc_kat.npz"), message=message, codeword=codeword, llr=llr,
                        frozen=np.asarray(enc.frozen_bits, dtype=np.int64),
                        **{f"dec_L{L}": v for L, v in dec.items()})
    print("doc KAT:", message, codeword, llr[:3])


def gen_wide():
    rng = np.random.default_rng(31337)
    np.random.seed(9001)
    out = {}
    scl_cases = [  # N, K, L, frozen kind, F, snr, kind
        (64, 32, 64, "random", 3, 0.0, "awgn"), (128, 64, 40, "bhatt", 2, -1.0, "awgn"),
        (32, 16, 48, "random", 3, 0.0, "int"), (64, 40, 100, "random", 2, 1.0, "raw"),
    ]
    for ci, (N, K, L, fk, F, snr, kind) in enumerate(scl_cases):
        frozen = np.sort(rng.choice(N, N - K, replace=False)) if fk == "random" else \
            pconstr.bhattacharyya_frozen_set(N, K, 2.0)
        llr = polar_inputs(rng, N, K, frozen, F, snr, kind)
        dec = SCLDecoder(N, K, list_size=L, frozen_bits=frozen)
        bits, pm = [], []
        for f in range(F):
            bits.append(dec.decode(llr[f]))
            pm.append(dec.path_metrics.copy())
        out[f"scl{ci}_N"] = N
        out[f"scl{ci}_L"] = L
        out[f"scl{ci}_frozen"] = np.asarray(frozen, dtype=np.int64)
        out[f"scl{ci}_llr"] = llr
        out[f"scl{ci}_bits"] = np.array(bits, dtype=np.int64)
        out[f"scl{ci}_pm"] = np.array(pm)
        print("wide scl case", ci, N, K, L, "done", flush=True)
    out["nscl"] = len(scl_cases)
    # parity checks of degree 44 and 70 next to ordinary ones, a degree-1 check
    r = np.random.RandomState(11)
    m, n = 12, 90
    H = np.zeros((m, n), dtype=int)
    for c in range(m):
        H[c, r.choice(n, r.randint(3, 9), replace=False)] = 1
    H[2, r.choice(n, 40, replace=False)] = 1
    H[5, :] = 0
    H[5, r.choice(n, 70, replace=False)] = 1
    H[9, :] = 0
    H[9, 7] = 1
    llr = rng.normal(1.5, 2.0, size=(5, n))
    for name, es in (("dense_bp", True), ("dense_bp_nostop", False)):
        dec = _BPTotals(H, max_iter=6, early_stop=es)
        bits, iters, tots = [], [], []
        for f in range(llr.shape[0]):
            b, k = dec.decode(llr[f], return_iterations=True)
            bits.append(b); iters.append(k); tots.append(dec._last_totals.copy())
        out[name + "_H"] = H.astype(np.uint8)
        out[name + "_cfg"] = np.array([0, 6, int(es)], dtype=np.int64)
        out[name + "_llr"] = llr
        out[name + "_bits"] = np.array(bits, dtype=np.int64)
        out[name + "_iters"] = np.array(iters, dtype=np.int64)
        out[name + "_total"] = np.array(tots)
        print("wide ldpc case", name, "degrees", H.sum(1), "iters", iters, flush=True)
    np.savez_compressed(os.path.join(HERE, "wide.npz"), **out)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "wide":
        gen_wide()
        sys.exit(0)
    gen_doc_kat()
    gen_ldpc()
    gen_polar()
    gen_wide()
