"""Pins oracle/pcl_oracle.c against golden vectors produced by the reference itself
(tests/golden/gen_golden.py ran /root/reference's SCDecoder / SCLDecoder / BPDecoder /
MSDecoder).  Decoded bits and iteration counts must be identical; leaf / total LLRs and
path metrics agree to fp64 round-off (numpy's SIMD exp/log1p/tanh differ from libm in the
last ulp)."""
import os

import numpy as np
import pytest

from oracle import oracle


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def test_doc_kat(golden_dir):
    g = _load(golden_dir, "doc_kat.npz")
    assert list(g["message"]) == [0, 1, 0, 0, 0, 1, 0, 0]
    assert list(g["codeword"]) == [0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 0, 0, 0, 0]
    for L in (1, 2, 4, 8):
        out = oracle.polar_scl(16, L, g["frozen"], g["llr"][None, :])
        assert np.array_equal(out[0], g[f"dec_L{L}"])
        assert np.array_equal(out[0], g["message"])


def test_sc_golden(golden_dir):
    g = _load(golden_dir, "polar_sc.npz")
    for ci in range(int(g["ncases"])):
        N = int(g[f"c{ci}_N"])
        bits, leaf = oracle.polar_sc(N, g[f"c{ci}_frozen"], g[f"c{ci}_llr"], want_leaf=True)
        assert np.array_equal(bits, g[f"c{ci}_bits"]), f"SC case {ci}"
        # f is exact, g is one correctly-rounded add: leaf LLRs are bit-identical
        assert np.array_equal(leaf, g[f"c{ci}_leaf"]), f"SC leaf case {ci}"


def test_scl_golden(golden_dir):
    g = _load(golden_dir, "polar_scl.npz")
    for ci in range(int(g["ncases"])):
        N, L = int(g[f"c{ci}_N"]), int(g[f"c{ci}_L"])
        bits, pm, leaf = oracle.polar_scl(N, L, g[f"c{ci}_frozen"], g[f"c{ci}_llr"],
                                          want_pm=True, want_leaf=True)
        assert np.array_equal(bits, g[f"c{ci}_bits"]), f"SCL case {ci}"
        assert np.array_equal(leaf, g[f"c{ci}_leaf"]), f"SCL leaf case {ci}"
        ref = g[f"c{ci}_pm"]
        assert np.array_equal(np.isinf(pm), np.isinf(ref))
        fin = np.isfinite(ref)
        np.testing.assert_allclose(pm[fin], ref[fin], rtol=1e-12, atol=1e-12)


def test_scl_l1_equals_sc(golden_dir):
    g = _load(golden_dir, "polar_sc.npz")
    for ci in range(int(g["ncases"])):
        N = int(g[f"c{ci}_N"])
        out = oracle.polar_scl(N, 1, g[f"c{ci}_frozen"], g[f"c{ci}_llr"])
        assert np.array_equal(out, g[f"c{ci}_bits"])


def test_ldpc_golden(golden_dir):
    g = _load(golden_dir, "ldpc.npz")
    for name in g["names"]:
        name = str(name)
        mode, it, es = (int(x) for x in g[name + "_cfg"])
        H = g[name + "_H"].astype(np.int64)
        bits, iters, total = oracle.ldpc(H, g[name + "_llr"], mode="bp" if mode == 0 else "ms",
                                         max_iter=it, normalization=float(g[name + "_norm"]),
                                         early_stop=bool(es), want_total=True)
        assert np.array_equal(bits, g[name + "_bits"]), name
        if name + "_iters" in g:
            assert np.array_equal(iters, g[name + "_iters"]), name
        if name + "_total" in g:
            np.testing.assert_allclose(total, g[name + "_total"], rtol=1e-9, atol=1e-9, err_msg=name)


def test_ms_degree1_raises(golden_dir):
    g = _load(golden_dir, "ldpc.npz")
    H = g["odd_bp_H"].astype(np.int64)
    with pytest.raises(ValueError):
        oracle.ldpc(H, np.ones((1, H.shape[1])), mode="ms", max_iter=3)


def test_threads_same_result(golden_dir):
    g = _load(golden_dir, "polar_scl.npz")
    ci = 10
    N, L = int(g[f"c{ci}_N"]), int(g[f"c{ci}_L"])
    a = oracle.polar_scl(N, L, g[f"c{ci}_frozen"], g[f"c{ci}_llr"], nthreads=1)
    b = oracle.polar_scl(N, L, g[f"c{ci}_frozen"], g[f"c{ci}_llr"], nthreads=4)
    assert np.array_equal(a, b)


def test_wide_golden(golden_dir):
    """List sizes above 32 and a BP check of degree 44 / 70: what the reference accepts beyond the
    width of a warp (tests/golden/gen_golden.py::gen_wide ran the reference)."""
    g = _load(golden_dir, "wide.npz")
    for ci in range(int(g["nscl"])):
        N, L = int(g[f"scl{ci}_N"]), int(g[f"scl{ci}_L"])
        bits, pm = oracle.polar_scl(N, L, g[f"scl{ci}_frozen"], g[f"scl{ci}_llr"], want_pm=True)
        assert np.array_equal(bits, g[f"scl{ci}_bits"]), f"wide SCL case {ci}"
        ref = g[f"scl{ci}_pm"]
        assert np.array_equal(np.isinf(pm), np.isinf(ref))
        fin = np.isfinite(ref)
        np.testing.assert_allclose(pm[fin], ref[fin], rtol=1e-12, atol=1e-12)
    for name in ("dense_bp", "dense_bp_nostop"):
        _, it, es = (int(x) for x in g[name + "_cfg"])
        bits, iters, total = oracle.ldpc(g[name + "_H"].astype(np.int64), g[name + "_llr"], mode="bp", max_iter=it,
                                         early_stop=bool(es), want_total=True)
        assert np.array_equal(bits, g[name + "_bits"]), name
        assert np.array_equal(iters, g[name + "_iters"]), name
        np.testing.assert_allclose(total, g[name + "_total"], rtol=1e-9, atol=1e-9, err_msg=name)
