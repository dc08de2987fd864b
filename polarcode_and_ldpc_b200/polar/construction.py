"""Library-free frozen-set construction for the benchmark configurations.

The reference's benchmarks take the frozen set from the un-vendored `polarcodes`
package (/root/reference/src/lib_wrappers/polar_wrapper.py:45-49).  This is the
Bhattacharyya-bound construction restated: equality with
that library's output is unverified ("parity unpinned" at this boundary), but the
frozen set is an *input* of the decode path, so decoder parity is unaffected.
It equals the in-repo bhattacharyya_bounds (src/polar/construction.py:11-48) read
at bit-reversed indices.
"""
from __future__ import annotations

import numpy as np


def _logdiff(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """log(exp(a) - exp(b)) for a > b."""
    return a + np.log1p(-np.exp(b - a))


def bhattacharyya_frozen_set(N: int, K: int, design_snr_db: float = 2.0) -> np.ndarray:
    """Sorted indices (reference index space) of the N-K frozen positions."""
    n = int(np.log2(N))
    assert 1 << n == N and 0 < K < N
    z = np.zeros(N, dtype=np.float64)
    z[0] = -(K / N) * 10.0 ** (design_snr_db / 10.0)
    for lev in range(1, n + 1):
        half = 1 << (lev - 1)
        t = z[:half].copy()
        z[:half] = _logdiff(np.log(2.0) + t, 2.0 * t)
        z[half:2 * half] = 2.0 * t
    order = np.argsort(z, kind="mergesort")
    return np.sort(order[K:]).astype(np.int64)


def bhattacharyya_bounds(N: int, snr_db: float) -> np.ndarray:
    """Bhattacharyya upper bounds Z(W_N^(i)), i = 0 .. N-1, of the reference helper of the same
    name (/root/reference/src/polar/construction.py:11-48): Z_0 = exp(-snr), then per level
    Z[2i] = 2 Z_i - Z_i^2 (degraded), Z[2i+1] = Z_i^2 (upgraded).  Vectorised over a level."""
    n = int(np.log2(N))
    z = np.array([np.exp(-(10.0 ** (snr_db / 10.0)))])
    for _ in range(n):
        nxt = np.empty(2 * z.size)
        nxt[0::2] = 2.0 * z - z * z
        nxt[1::2] = z * z
        z = nxt
    return z


def construct_polar_code(N: int, K: int, method: str = "bhattacharyya", snr_db: float = 0.0):
    """(frozen_bits, info_bits), both sorted -- reference src/polar/construction.py:100-140.  The
    "bhattacharyya" and "default" (bit-reversal heuristic) methods are provided; the Gaussian
    approximation variant is an out-of-scope construction helper."""
    if method == "bhattacharyya":
        order = np.argsort(bhattacharyya_bounds(N, snr_db))
        info, frozen = order[:K], order[K:]
    elif method == "gaussian_approximation":
        raise NotImplementedError("gaussian_approximation construction is not part of the decode hot path")
    else:
        from .utils import bit_reverse_permutation
        order = np.argsort(bit_reverse_permutation(int(np.log2(N))))
        info, frozen = order[-K:], order[:-K]
    return np.sort(frozen), np.sort(info)
