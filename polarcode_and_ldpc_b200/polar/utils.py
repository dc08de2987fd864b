"""Host-side polar helpers whose *semantics* the decode path needs.

Reference: /root/reference/src/polar/utils.py -- bit_reverse (:11-26),
generate_frozen_bits default rule (:48-83), crc_encode / crc_check (:86-163).
Vectorised restatements; no decode work happens here.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np

# polynomial, length -- utils.py:100-104 (an unknown name falls back to CRC-8, :104-105)
CRC_POLYNOMIALS = {"CRC-8": (0x1D, 8), "CRC-16": (0x1021, 16), "CRC-24": (0x1864CFB, 24)}


def bit_reverse(value: int, num_bits: int) -> int:
    """n-bit reversal of one integer (utils.py:11-26)."""
    out = 0
    for _ in range(num_bits):
        out = (out << 1) | (value & 1)
        value >>= 1
    return out


def bit_reverse_permutation(num_bits: int) -> np.ndarray:
    """perm[i] = bit_reverse(i, num_bits) for all i < 2**num_bits."""
    idx = np.arange(1 << num_bits, dtype=np.int64)
    out = np.zeros_like(idx)
    for b in range(num_bits):
        out |= ((idx >> b) & 1) << (num_bits - 1 - b)
    return out


def generate_frozen_bits(N: int, K: int, channel_param: np.ndarray = None) -> Tuple[np.ndarray, np.ndarray]:
    """Default frozen/info split used when a decoder gets frozen_bits=None.

    utils.py:64-83: without channel parameters the K indices whose bit-reversed
    value is largest carry information; with parameters, the K smallest do.
    """
    if channel_param is None:
        n = int(np.log2(N))
        order = np.argsort(bit_reverse_permutation(n))
        info, frozen = order[-K:], order[:-K]
    else:
        order = np.argsort(channel_param)
        info, frozen = order[:K], order[K:]
    return np.sort(frozen), np.sort(info)


def crc_resolve(polynomial: str) -> Tuple[int, int]:
    """(poly, length) with the reference's silent CRC-8 fallback."""
    if polynomial not in CRC_POLYNOMIALS:
        polynomial = "CRC-8"
    return CRC_POLYNOMIALS[polynomial]


def _crc_register(bits: np.ndarray, poly: int, crc_len: int) -> int:
    top, mask, reg = 1 << (crc_len - 1), (1 << crc_len) - 1, 0
    for b in np.asarray(bits, dtype=np.int64):
        reg ^= int(b) << (crc_len - 1)
        reg = ((reg << 1) ^ poly) if (reg & top) else (reg << 1)
        reg &= mask
    return reg


def crc_encode(data: np.ndarray, polynomial: str = "CRC-8") -> np.ndarray:
    """data || crc, MSB first, zero initial register (utils.py:86-125)."""
    poly, crc_len = crc_resolve(polynomial)
    reg = _crc_register(data, poly, crc_len)
    tail = np.array([(reg >> i) & 1 for i in range(crc_len - 1, -1, -1)], dtype=int)
    return np.concatenate([np.asarray(data), tail])


def crc_check(data: np.ndarray, polynomial: str = "CRC-8") -> bool:
    """True when the register over data||crc ends at zero (utils.py:128-163)."""
    poly, crc_len = crc_resolve(polynomial)
    return _crc_register(data, poly, crc_len) == 0


def polar_transform(u: np.ndarray) -> np.ndarray:
    """x = u . F^{(x)n} over GF(2) along the last axis (utils.py:193-229), batched."""
    x = np.array(u, dtype=np.uint8, copy=True)
    N = x.shape[-1]
    lead = x.shape[:-1]
    stride = 1
    while stride < N:
        v = x.reshape(lead + (N // (2 * stride), 2, stride))
        v[..., 0, :] ^= v[..., 1, :]
        stride *= 2
    return x
