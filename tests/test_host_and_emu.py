"""CPU-only tests: host logic, the C-ABI library's exported surface, and the kernel
sources' logic run through the g++/SIMT-emulator build against the golden vectors."""
import ctypes
import os
import re

import numpy as np
import pytest

import polarcode_and_ldpc_b200 as P
from polarcode_and_ldpc_b200 import _build, _native
from polarcode_and_ldpc_b200.polar import utils as putils
from oracle import oracle
from tests.emu import emu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_and_exports_every_declared_symbol():
    path = _build.build_native()
    lib = ctypes.CDLL(path)
    hdr = open(os.path.join(ROOT, "include", "pcl.h")).read()
    declared = set(re.findall(r"\b(pcl_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_native.EXPORTS), declared ^ set(_native.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    lib.pcl_version.restype = ctypes.c_int
    assert lib.pcl_version() >= 100


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(_native.PclError):
        P.SCLDecoder(64, 32)
    with pytest.raises(_native.PclError):
        P.BPDecoder(P.gallager_parity_check(24, 3, 4, 7))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "polarcode_and_ldpc_b200")
    for dp, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, fn)).read()
                assert not re.search(r"(import\s+oracle|from\s+oracle|libpcl_oracle|oracle\.)", src), fn


def test_host_helpers_match_reference_semantics(golden_dir):
    assert [putils.bit_reverse(i, 3) for i in range(8)] == [0, 4, 2, 6, 1, 5, 3, 7]
    assert np.array_equal(putils.bit_reverse_permutation(4), [putils.bit_reverse(i, 4) for i in range(16)])
    g = np.load(os.path.join(golden_dir, "doc_kat.npz"))
    enc = P.PolarEncoder(16, 8)                      # default frozen rule
    assert np.array_equal(enc.frozen_bits, g["frozen"])
    assert np.array_equal(enc.encode(g["message"]), g["codeword"])
    data = np.array([1, 0, 1, 1, 0, 0, 1, 0, 1, 1])
    for poly in ("CRC-8", "CRC-16", "CRC-24", "nonsense"):
        cw = putils.crc_encode(data, poly)
        assert putils.crc_check(cw, poly)
        cw[3] ^= 1
        assert not putils.crc_check(cw, poly)
    fz = P.bhattacharyya_frozen_set(1024, 512, 2.0)
    assert len(fz) == 512 and list(np.setdiff1d(np.arange(1024), fz)[:8]) == [15, 23, 27, 29, 30, 31, 39, 43]


def test_awgn_batch_equals_sequential_transmit():
    bits = np.random.default_rng(0).integers(0, 2, size=(5, 33))   # odd N: cached gauss carries over
    np.random.seed(3)
    ch = P.AWGNChannel(1.5)
    seq = np.array([ch.transmit(b) for b in bits])
    np.random.seed(3)
    assert np.array_equal(P.AWGNChannel(1.5).transmit_batch(bits), seq)
    assert abs(ch.noise_std - np.sqrt(1 / (2 * 10 ** 0.15))) < 1e-15


def test_ldpc_constructions():
    H = P.gallager_parity_check(504, 3, 6, 42)
    assert H.shape == (252, 504) and set(H.sum(0)) == {3} and set(H.sum(1)) == {6}
    G, info = P.generator_from_parity(H)
    assert G.shape == (254, 504) and not ((H @ G.T) % 2).any() and np.array_equal(G[:, info], np.eye(254, dtype=int))
    Hm = P.mackay_parity_check(504, 252, 3, 6, seed=42)
    assert list(np.bincount(Hm.sum(1))) == [1, 2, 7, 29, 35, 35, 47, 34, 23, 15, 13, 8, 2, 1]
    enc = P.LDPCEncoder(504, 252, H=H)
    cw = enc.encode_batch(np.random.default_rng(1).integers(0, 2, size=(7, enc.k)))
    assert not ((H @ cw.T) % 2).any()


def test_shard_range_partitions():
    for F in (0, 1, 7, 1000):
        for W in (1, 2, 3, 8):
            spans = [P.shard_range(F, r, W) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == F
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


# ---- kernel logic through the emulator (same .cuh sources as the CUDA build) ----
@pytest.mark.parametrize("generic", [0, 1])
@pytest.mark.parametrize("reverse", [False, True])
def test_emu_polar_golden(golden_dir, reverse, generic):
    """generic=1 forces the all-shared-memory kernel (polar_scl.cuh); 0 lets the library pick
    the register-resident-bottom kernel (polar_scl_fast.cuh) wherever it applies (N >= 16)."""
    env = {"PCL_POLAR_GENERIC": generic}
    g = np.load(os.path.join(golden_dir, "polar_scl.npz"))
    for ci in range(int(g["ncases"])):
        N, L, fz, llr = int(g[f"c{ci}_N"]), int(g[f"c{ci}_L"]), g[f"c{ci}_frozen"], g[f"c{ci}_llr"]
        if N > 256:
            continue
        bits, pm, (leaf, par) = emu.polar_decode(N, N - len(fz), L, fz, llr[:3], "f64", want_pm=True,
                                                 want_leaf=True, reverse=reverse, env=env)
        assert (emu.polar_decode.last_fast > 0) == (not (generic or N < 16))
        assert np.array_equal(bits, g[f"c{ci}_bits"][:3]), f"case {ci}"
        ref = g[f"c{ci}_pm"][:3]
        fin = np.isfinite(ref)
        np.testing.assert_allclose(pm[fin], ref[fin], rtol=1e-12, atol=1e-11)
    g = np.load(os.path.join(golden_dir, "polar_sc.npz"))
    for ci in range(int(g["ncases"])):
        N, fz, llr = int(g[f"c{ci}_N"]), g[f"c{ci}_frozen"], g[f"c{ci}_llr"]
        if N > 256:
            continue
        for dt in ("f64", "f32"):
            bits = emu.polar_decode(N, N - len(fz), 1, fz, llr[:3], dt, reverse=reverse, env=env)
            assert np.array_equal(bits, g[f"c{ci}_bits"][:3]), f"SC case {ci} {dt}"


def test_emu_polar_global_levels_and_crc():
    N, K, L = 128, 70, 8
    fz = P.bhattacharyya_frozen_set(N, K, 2.0)
    enc = P.PolarEncoder(N, K, fz, use_crc=True)
    rng = np.random.default_rng(2)
    np.random.seed(2)
    llr = P.AWGNChannel(1.0).transmit_batch(enc.encode_batch(rng.integers(0, 2, size=(6, enc.K_data))))
    ref = oracle.polar_scl(N, L, fz, llr)
    for generic in (0, 1):
        for G in (0, 1, 3, 6):
            env = {"PCL_POLAR_G": G, "PCL_POLAR_GENERIC": generic}
            assert np.array_equal(emu.polar_decode(N, K, L, fz, llr, "f64", env=env), ref)
            assert np.array_equal(emu.polar_decode(N, K, L, fz, llr, "f32", env=env), ref)
        ref_crc = oracle.polar_scl(N, L, fz, llr, use_crc=True)
        assert np.array_equal(emu.polar_decode(N, K, L, fz, llr, "f64", crc=(0x1D, 8),
                                               env={"PCL_POLAR_GENERIC": generic}), ref_crc)


def test_emu_polar_tensor_memory_variant():
    """The one-block-per-SM variant of the fast kernel (tensor-memory / shared-memory mid levels, level 3
    fused from the channel through the cp.async ring, ticket-scheduled warp groups) on the SIMT emulator,
    both lane orders: bits of the oracle for SCL-8 N = 1024 with more chunks than warp groups."""
    N, K, L = 1024, 512, 8
    fz = P.bhattacharyya_frozen_set(N, K, 2.0)
    enc = P.PolarEncoder(N, K, fz)
    rng = np.random.default_rng(3)
    np.random.seed(5)
    llr = P.AWGNChannel(-1.0).transmit_batch(enc.encode_batch(rng.integers(0, 2, size=(37, K))))
    ref = oracle.polar_scl(N, L, fz, llr)
    for reverse in (False, True):
        got = emu.polar_decode(N, K, L, fz, llr, "f32", reverse=reverse)
        assert emu.polar_decode.last_fast == 3
        assert np.array_equal(got, ref)
    got = emu.polar_decode(N, K, L, fz, llr[:9], "f32", env={"PCL_POLAR_TM": 0})
    assert emu.polar_decode.last_fast == 2 and np.array_equal(got, ref[:9])
    # list size 32 takes the same variant by default (one frame per warp)
    ref32 = oracle.polar_scl(N, 32, fz, llr[:5])
    got = emu.polar_decode(N, K, 32, fz, llr[:5], "f32")
    assert emu.polar_decode.last_fast == 3 and np.array_equal(got, ref32)
    got = emu.polar_decode(N, K, 32, fz, llr[:5], "f32", env={"PCL_POLAR_TM32": 0})
    assert emu.polar_decode.last_fast == 2 and np.array_equal(got, ref32)
    # ... and so do list size 16 and SCL-8 at N = 2048 (two levels in the global scratch instead of one)
    got = emu.polar_decode(N, K, 16, fz, llr[:6], "f32", reverse=True)
    assert emu.polar_decode.last_fast == 3 and np.array_equal(got, oracle.polar_scl(N, 16, fz, llr[:6]))
    N2, K2 = 2048, 1024
    fz2 = P.bhattacharyya_frozen_set(N2, K2, 2.0)
    llr2 = P.AWGNChannel(-0.5).transmit_batch(P.PolarEncoder(N2, K2, fz2).encode_batch(rng.integers(0, 2, size=(9, K2))))
    got = emu.polar_decode(N2, K2, 8, fz2, llr2, "f32")
    assert emu.polar_decode.last_fast == 3 and np.array_equal(got, oracle.polar_scl(N2, 8, fz2, llr2))


def test_emu_polar_large_code_falls_back_instead_of_failing():
    """N = 8192 with list size 1 (and with CRC selection) used to be refused: the fast kernel's bit
    arrays need 32-64 KB per warp whatever the list size.  The handle now retries with fewer warps per
    block and then with the generic kernel."""
    N, K = 8192, 4096
    fz = P.bhattacharyya_frozen_set(N, K, 2.0)
    rng = np.random.default_rng(4)
    np.random.seed(4)
    llr = P.AWGNChannel(2.0).transmit_batch(P.PolarEncoder(N, K, fz).encode_batch(rng.integers(0, 2, size=(2, K))))
    assert np.array_equal(emu.polar_decode(N, K, 1, fz, llr, "f32"), oracle.polar_sc(N, fz, llr))
    enc = P.PolarEncoder(N, K, fz, use_crc=True)
    llr = P.AWGNChannel(2.0).transmit_batch(enc.encode_batch(rng.integers(0, 2, size=(1, enc.K_data))))
    ref = oracle.polar_scl(N, 2, fz, llr, use_crc=True)
    assert np.array_equal(emu.polar_decode(N, K, 2, fz, llr, "f32", crc=(0x1D, 8)), ref)


def test_emu_sc1024_register_resident_kernel():
    """SC N = 512 / 1024 / 2048 as 2 / 4 / 8 length-256 codes in a row (polar_sc_big_kernel<M>): warp-cooperative
    rows of level-m LLRs, bit-reversed parked partial sums folded upwards per frame, the N = 256 decoder per lane --
    bits of the oracle in both lane orders, with Bhattacharyya and random frozen sets, batches that leave the last
    warp partly empty."""
    rng = np.random.default_rng(3)
    for N, K, snr, F in ((1024, 512, 0.0, 37), (1024, 100, -2.0, 33), (512, 256, 0.0, 35), (2048, 1024, 0.5, 33), (2048, 300, -2.0, 33),
                          (4096, 2048, 0.5, 33)):
        fz = P.bhattacharyya_frozen_set(N, K, 2.0) if 2 * K == N else np.sort(rng.choice(N, N - K, replace=False))
        np.random.seed(5)
        llr = P.AWGNChannel(snr).transmit_batch(P.PolarEncoder(N, K, fz).encode_batch(rng.integers(0, 2, size=(F, K))))
        ref = oracle.polar_sc(N, fz, llr)
        for reverse in (False, True):
            got = emu.polar_decode(N, K, 1, fz, llr, "f32", reverse=reverse)
            assert emu.polar_decode.last_fast == 4 and np.array_equal(got, ref)
        got = emu.polar_decode(N, K, 1, fz, llr[:5], "f32", env={"PCL_POLAR_SC1024": 0})
        assert emu.polar_decode.last_fast != 4 and np.array_equal(got, ref[:5])


def test_emu_wide_list_and_wide_checks(golden_dir):
    """Beyond a warp's width: list sizes 40 .. 100 through the block-per-frame
    kernel (polar_scl_wide.cuh), both lane orders, CRC selection, and BP checks of degree 44 / 70 through
    the any-degree rule (cn_bp_loop) -- against vectors the reference itself produced."""
    g = np.load(os.path.join(golden_dir, "wide.npz"))
    for ci in range(int(g["nscl"])):
        N, L, fz, llr = int(g[f"scl{ci}_N"]), int(g[f"scl{ci}_L"]), g[f"scl{ci}_frozen"], g[f"scl{ci}_llr"]
        for reverse in (False, True):
            bits, pm = emu.polar_decode(N, N - len(fz), L, fz, llr, "f64", want_pm=True, reverse=reverse)[:2]
            assert np.array_equal(bits, g[f"scl{ci}_bits"]), f"wide case {ci}"
            ref = g[f"scl{ci}_pm"]
            fin = np.isfinite(ref)
            assert np.array_equal(np.isfinite(pm), fin)
            np.testing.assert_allclose(pm[fin], ref[fin], rtol=1e-12, atol=1e-11)
    N, K, L = 128, 70, 64
    fz = P.bhattacharyya_frozen_set(N, K, 2.0)
    enc = P.PolarEncoder(N, K, fz, use_crc=True)
    rng = np.random.default_rng(12)
    np.random.seed(12)
    llr = P.AWGNChannel(0.0).transmit_batch(enc.encode_batch(rng.integers(0, 2, size=(3, enc.K_data))))
    for dt in ("f64", "f32"):
        assert np.array_equal(emu.polar_decode(N, K, L, fz, llr, dt), oracle.polar_scl(N, L, fz, llr))
        assert np.array_equal(emu.polar_decode(N, K, L, fz, llr, dt, crc=(0x1D, 8)),
                              oracle.polar_scl(N, L, fz, llr, use_crc=True))
    for name in ("dense_bp", "dense_bp_nostop"):
        _, it, es = (int(x) for x in g[name + "_cfg"])
        H = g[name + "_H"].astype(np.int64)
        bits, iters, tot = emu.ldpc_decode(H, g[name + "_llr"], "bp", it, 1.0, bool(es), "f64", want_total=True)
        assert np.array_equal(bits, g[name + "_bits"]) and np.array_equal(iters, g[name + "_iters"])
        np.testing.assert_allclose(tot, g[name + "_total"], rtol=1e-9, atol=1e-9)
        bits, iters, tot = emu.ldpc_decode(H, g[name + "_llr"], "bp", it, 1.0, bool(es), "f32", want_total=True)
        assert np.array_equal(bits, g[name + "_bits"]) and np.array_equal(iters, g[name + "_iters"])
        assert np.max(np.abs(tot - g[name + "_total"]) / np.maximum(np.abs(g[name + "_total"]), 1.0)) < 1e-4


def test_emu_ldpc_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "ldpc.npz"))
    for name in g["names"]:
        name = str(name)
        if "504" in name:
            continue
        mode, it, es = (int(x) for x in g[name + "_cfg"])
        for dt in ("f64", "f32"):
            bits, iters, tot = emu.ldpc_decode(g[name + "_H"].astype(np.int64), g[name + "_llr"][:3],
                                               "bp" if mode == 0 else "ms", it, float(g[name + "_norm"]),
                                               bool(es), dt, want_total=True)
            assert np.array_equal(bits, g[name + "_bits"][:3]), f"{name} {dt}"
            if name + "_iters" in g:
                assert np.array_equal(iters, g[name + "_iters"][:3])
            if name + "_total" in g:
                ref = g[name + "_total"][:3]
                tol = 1e-9 if dt == "f64" else 1e-4
                assert np.max(np.abs(tot - ref) / np.maximum(np.abs(ref), 1.0)) < tol


def test_emu_ldpc_block_cooperative_mode():
    """PCL_LDPC_COOP=1: the whole block decodes one frame (the mode large codes get); same bits
    and iteration counts as one warp per frame and as the oracle, BP and Min-Sum, early stop on/off."""
    rng = np.random.default_rng(0)
    H = P.gallager_parity_check(96, 3, 6, 42)
    llr = rng.normal(1.0, 2.2, size=(5, 96))
    try:
        for mode in ("bp", "ms"):
            for es in (True, False):
                rb, ri = oracle.ldpc(H, llr, mode, max_iter=8, normalization=0.75, early_stop=es)
                for coop in ("0", "1"):
                    os.environ["PCL_LDPC_COOP"] = coop
                    for dt in ("f64", "f32"):
                        b, it, _ = emu.ldpc_decode(H, llr, mode, 8, 0.75, es, dt, want_total=True)
                        assert np.array_equal(b, rb) and np.array_equal(it, ri), (mode, es, coop, dt)
        Hm = P.mackay_parity_check(120, 60, 3, 6, seed=42)
        llr = rng.normal(1, 2, size=(3, 120))
        rb, ri = oracle.ldpc(Hm, llr, "bp", max_iter=5)
        os.environ["PCL_LDPC_COOP"] = "1"
        b, it = emu.ldpc_decode(Hm, llr, "bp", 5)[:2]
        assert np.array_equal(b, rb) and np.array_equal(it, ri)
    finally:
        os.environ.pop("PCL_LDPC_COOP", None)
