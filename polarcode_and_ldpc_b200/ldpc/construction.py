"""Host-side parity-check constructions used as decoder *inputs*.

* gallager_parity_check: regular (dv, dc) Gallager ensemble, the construction
  the reference obtains from the un-vendored `pyldpc.make_ldpc`
  (/root/reference/src/lib_wrappers/ldpc_wrapper.py:52).  Restated from the
  published algorithm; pyldpc's later column permutation
  is not reproduced -> "parity unpinned" at this boundary (H is an input of the
  decode path, decoder parity is unaffected).
* mackay_parity_check: the in-repo column-random construction
  (/root/reference/src/ldpc/matrix.py:12-50), same legacy-RNG call order, hence
  the same H for the same seed (irregular rows, degree 0..13 at n=504).
* generator_from_parity: GF(2) null-space generator for benchmark codewords.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np


def gallager_parity_check(n: int, dv: int = 3, dc: int = 6, seed: Optional[int] = 42) -> np.ndarray:
    if n % dc:
        raise ValueError("dc must divide n")
    rng = np.random.RandomState(seed)
    rows_per_block = n // dc
    block = np.zeros((rows_per_block, n), dtype=int)
    for i in range(rows_per_block):
        block[i, i * dc:(i + 1) * dc] = 1
    blocks = [block]
    for _ in range(dv - 1):
        blocks.append(rng.permutation(block.T).T)
    return np.concatenate(blocks, axis=0)


def mackay_parity_check(n: int, k: int, dv: int = 3, dc: int = 6, seed: Optional[int] = None) -> np.ndarray:
    m = n - k
    if dv * n != dc * m:
        raise ValueError(f"Degree constraint not satisfied: dv*n={dv*n} != dc*m={dc*m}")
    if seed is not None:
        np.random.seed(seed)
    H = np.zeros((m, n), dtype=int)
    for col in range(n):
        H[np.random.choice(m, dv, replace=False), col] = 1
    return H


def generator_from_parity(H: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """(G[k, n], info_positions[k]) with H @ G.T == 0 (mod 2), G[:, info] == I.

    Plain Gauss-Jordan over GF(2) with free column choice, so it works for the
    rank-deficient Gallager H (k = n - rank) as well.
    """
    A = (np.asarray(H) & 1).astype(np.uint8)
    m, n = A.shape
    pivots = []
    r = 0
    for c in range(n - 1, -1, -1):       # prefer parity at the tail, like [I | P]
        if r == m:
            break
        rows = np.flatnonzero(A[r:, c]) + r
        if rows.size == 0:
            continue
        if rows[0] != r:
            A[[r, rows[0]]] = A[[rows[0], r]]
        elim = np.flatnonzero(A[:, c])
        elim = elim[elim != r]
        A[elim] ^= A[r]
        pivots.append(c)
        r += 1
    pivots = np.array(pivots, dtype=np.int64)
    free = np.setdiff1d(np.arange(n), pivots)
    k = free.size
    G = np.zeros((k, n), dtype=np.uint8)
    G[np.arange(k), free] = 1
    # pivot row i reads: x[pivots[i]] = sum_j A[i, free_j] x[free_j]
    G[:, pivots] = A[:len(pivots)][:, free].T
    return G.astype(np.int64), free


# ---- the reference's helper names (src/ldpc/matrix.py, src/ldpc/utils.py), inputs only -----
def mackay_construction(n: int, k: int, dv: int, dc: int, seed: Optional[int] = None) -> np.ndarray:
    """Reference name of mackay_parity_check (/root/reference/src/ldpc/matrix.py:12)."""
    return mackay_parity_check(n, k, dv, dc, seed)


def generate_ldpc_matrix(n: int, k: int, method: str = "mackay", dv: int = 3, dc: int = 6,
                         seed: Optional[int] = None) -> np.ndarray:
    """H[m, n], m = n - k (/root/reference/src/ldpc/matrix.py:53-91): "mackay" re-derives dc from
    dv n / m when the degrees do not balance, "random" draws iid bits from the legacy RNG."""
    m = n - k
    if method == "mackay":
        if dv * n != dc * m:
            dc = (dv * n) // m
        return mackay_parity_check(n, k, dv, dc, seed)
    if method == "random":
        if seed is not None:
            np.random.seed(seed)
        return np.random.randint(0, 2, (m, n))
    raise ValueError(f"Unknown method: {method}")


def create_tanner_graph(H: np.ndarray):
    """(var_neighbors, check_neighbors) adjacency lists, ascending (/root/reference/src/ldpc/utils.py:11-34)."""
    H = np.asarray(H)
    check_neighbors = [np.flatnonzero(row == 1).tolist() for row in H]
    var_neighbors = [np.flatnonzero(col == 1).tolist() for col in H.T]
    return var_neighbors, check_neighbors


def check_syndrome(H: np.ndarray, codeword: np.ndarray) -> bool:
    """H c^T == 0 over GF(2) (/root/reference/src/ldpc/utils.py:37-49)."""
    return bool(np.all((np.asarray(H) @ np.asarray(codeword)) % 2 == 0))
