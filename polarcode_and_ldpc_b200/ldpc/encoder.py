"""Batched LDPC encoder (host side, input generation only).

Reference: /root/reference/src/ldpc/encoder.py:76-93 (c = m . G mod 2).  The
reference's direct-solving fallback for a singular parity part emits invalid
codewords and is deliberately not reproduced; a GF(2)
null-space generator is used instead when no G is supplied.
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from .construction import generator_from_parity, mackay_parity_check


class LDPCEncoder:
    def __init__(self, n: int, k: int, H: Optional[np.ndarray] = None, G: Optional[np.ndarray] = None,
                 dv: int = 3, dc: int = 6, seed: Optional[int] = None):
        assert n > 0
        self.n = n
        self.H = mackay_parity_check(n, k, dv, dc, seed) if H is None else np.asarray(H)
        assert self.H.shape[1] == n, f"H matrix must have {n} columns"
        self.m = self.H.shape[0]
        if G is not None:
            G = np.asarray(G)
            # pyldpc hands back (n, k); the reference accepts either orientation and nothing else (encoder.py:57-63)
            if G.shape == (n, k) and n != k:
                self.G = G.T
            elif G.shape == (k, n):
                self.G = G
            else:
                raise ValueError(f"G shape {G.shape} doesn't match (n,k)={n, k} or (k,n)={k, n}")
            self.info_positions = np.arange(self.G.shape[0])
        else:
            self.G, self.info_positions = generator_from_parity(self.H)
        # k is the true dimension of the code: equal to the caller's k for a supplied G, and
        # n - rank(H) when G comes from H (the reference keeps the nominal k there, documented difference)
        self.k_nominal = k
        self.k = self.G.shape[0]

    def encode(self, message: np.ndarray) -> np.ndarray:
        return self.encode_batch(np.asarray(message)[None, :])[0]

    def encode_batch(self, messages: np.ndarray) -> np.ndarray:
        messages = np.asarray(messages)
        assert messages.ndim == 2 and messages.shape[1] == self.k, f"Message length must be {self.k}"
        return ((messages.astype(np.int64) @ self.G) % 2).astype(np.int64)

    def verify_codeword(self, codeword: np.ndarray) -> bool:
        return bool(np.all((self.H @ np.asarray(codeword)) % 2 == 0))

    def get_code_rate(self) -> float:
        return self.k / self.n

    def get_parity_check_matrix(self) -> np.ndarray:
        return self.H.copy()

    def __repr__(self) -> str:
        return f"LDPCEncoder(n={self.n}, k={self.k}, rate={self.get_code_rate():.3f})"
