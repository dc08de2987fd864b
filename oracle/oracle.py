"""ctypes front-end of the CPU oracle (oracle/pcl_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of pcl_oracle.c.  The product package
(polarcode_and_ldpc_b200/) never imports this module; only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.

Mirrors the call shapes of the reference decoders
(/root/reference/src/polar/decoder.py, /root/reference/src/ldpc/decoder.py) but
frame-batched: every function takes llr[F, N] and returns row f == decode(llr[f]).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libpcl_oracle.so")
_lib = None

CRC_POLYS = {"CRC-8": (0x1D, 8), "CRC-16": (0x1021, 16), "CRC-24": (0x1864CFB, 24)}


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (recipe: oracle/Makefile)."""
    src = os.path.join(_HERE, "pcl_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        c_u8p = ctypes.POINTER(ctypes.c_uint8)
        c_f64p = ctypes.POINTER(ctypes.c_double)
        c_i32p = ctypes.POINTER(ctypes.c_int32)
        L.oracle_polar_sc.argtypes = [ctypes.c_int, c_u8p, c_f64p, ctypes.c_int64, c_u8p, c_f64p, ctypes.c_int]
        L.oracle_polar_sc.restype = ctypes.c_int
        L.oracle_polar_scl.argtypes = [ctypes.c_int, ctypes.c_int, c_u8p, c_f64p, ctypes.c_int64, c_u8p,
                                       c_f64p, c_f64p, ctypes.c_int, ctypes.c_uint32, ctypes.c_int,
                                       ctypes.c_int]
        L.oracle_polar_scl.restype = ctypes.c_int
        L.oracle_ldpc.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_i32p, c_i32p,
                                  c_i32p, c_i32p, ctypes.c_double, ctypes.c_int, ctypes.c_int, c_f64p,
                                  ctypes.c_int64, c_u8p, c_i32p, c_f64p, ctypes.c_int]
        L.oracle_ldpc.restype = ctypes.c_int
        L.oracle_max_threads.restype = ctypes.c_int
        _lib = L
    return _lib


def _p(a, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct)) if a is not None else None


def max_threads() -> int:
    return int(lib().oracle_max_threads())


def _frozen_mask(N: int, frozen_bits) -> np.ndarray:
    m = np.zeros(N, dtype=np.uint8)
    m[np.asarray(frozen_bits, dtype=np.int64)] = 1
    return m


def polar_sc(N: int, frozen_bits, llr: np.ndarray, want_leaf: bool = False, nthreads: int = 1):
    """SCDecoder.decode over a batch. Returns info bits [F, K] (and leaf LLRs [F, N])."""
    llr = np.ascontiguousarray(np.atleast_2d(llr), dtype=np.float64)
    F = llr.shape[0]
    assert llr.shape[1] == N
    fm = _frozen_mask(N, frozen_bits)
    u = np.zeros((F, N), dtype=np.uint8)
    leaf = np.zeros((F, N), dtype=np.float64) if want_leaf else None
    rc = lib().oracle_polar_sc(N, _p(fm, ctypes.c_uint8), _p(llr, ctypes.c_double), F,
                               _p(u, ctypes.c_uint8), _p(leaf, ctypes.c_double), nthreads)
    if rc:
        raise AssertionError(f"oracle_polar_sc rc={rc}")
    info = np.flatnonzero(fm == 0)
    out = u[:, info].astype(np.int64)
    return (out, leaf) if want_leaf else out


def polar_scl(N: int, L: int, frozen_bits, llr: np.ndarray, want_pm: bool = False,
              want_leaf: bool = False, use_crc: bool = False, crc_polynomial: str = "CRC-8",
              nthreads: int = 1):
    """SCLDecoder.decode over a batch. Returns info bits [F, K] (+ pm [F, L], leaf [F, N])."""
    llr = np.ascontiguousarray(np.atleast_2d(llr), dtype=np.float64)
    F = llr.shape[0]
    assert llr.shape[1] == N
    fm = _frozen_mask(N, frozen_bits)
    u = np.zeros((F, N), dtype=np.uint8)
    pm = np.zeros((F, L), dtype=np.float64)
    leaf = np.zeros((F, N), dtype=np.float64) if want_leaf else None
    poly, clen = CRC_POLYS.get(crc_polynomial, CRC_POLYS["CRC-8"])
    rc = lib().oracle_polar_scl(N, L, _p(fm, ctypes.c_uint8), _p(llr, ctypes.c_double), F,
                                _p(u, ctypes.c_uint8), _p(pm, ctypes.c_double),
                                _p(leaf, ctypes.c_double), int(use_crc), poly, clen, nthreads)
    if rc:
        raise AssertionError(f"oracle_polar_scl rc={rc}")
    info = np.flatnonzero(fm == 0)
    out = [u[:, info].astype(np.int64)]
    if want_pm:
        out.append(pm)
    if want_leaf:
        out.append(leaf)
    return out[0] if len(out) == 1 else tuple(out)


def tanner_tables(H: np.ndarray) -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
    """Edge tables in the reference's neighbour order (ldpc/decoder.py:43-47).

    Edges are enumerated check-major (row-major scan of H == 1); vperm lists the
    same edge ids variable-major, checks ascending.
    """
    H = np.asarray(H)
    rows, cols = np.nonzero(H == 1)
    m, n = H.shape
    cptr = np.zeros(m + 1, dtype=np.int32)
    np.cumsum(np.bincount(rows, minlength=m), out=cptr[1:])
    vperm = np.lexsort((rows, cols)).astype(np.int32)
    vptr = np.zeros(n + 1, dtype=np.int32)
    np.cumsum(np.bincount(cols, minlength=n), out=vptr[1:])
    return cptr, cols.astype(np.int32), vptr, vperm


def ldpc(H: np.ndarray, llr: np.ndarray, mode: str = "bp", max_iter: int = 50,
         normalization: float = 1.0, early_stop: bool = True, want_total: bool = False,
         nthreads: int = 1):
    """BPDecoder / MSDecoder .decode over a batch.

    Returns (bits[F, n] int64, iterations[F] int32[, total_llr[F, n]]).
    """
    H = np.asarray(H)
    m, n = H.shape
    llr = np.ascontiguousarray(np.atleast_2d(llr), dtype=np.float64)
    F = llr.shape[0]
    assert llr.shape[1] == n
    cptr, col, vptr, vperm = tanner_tables(H)
    E = int(col.shape[0])
    bits = np.zeros((F, n), dtype=np.uint8)
    iters = np.zeros(F, dtype=np.int32)
    total = np.zeros((F, n), dtype=np.float64) if want_total else None
    rc = lib().oracle_ldpc(0 if mode == "bp" else 1, m, n, E, _p(cptr, ctypes.c_int32),
                           _p(col, ctypes.c_int32), _p(vptr, ctypes.c_int32),
                           _p(vperm, ctypes.c_int32), float(normalization), int(max_iter),
                           int(bool(early_stop)), _p(llr, ctypes.c_double), F,
                           _p(bits, ctypes.c_uint8), _p(iters, ctypes.c_int32),
                           _p(total, ctypes.c_double), nthreads)
    if rc == 2:
        raise ValueError("zero-size array to reduction operation minimum which has no identity")
    if rc == 1:
        raise UnboundLocalError("max_iter < 1: the reference never binds 'decoded'")
    out = (bits.astype(np.int64), iters)
    return out + (total,) if want_total else out
