"""TEST-ONLY driver for the g++/SIMT-emulator build of the kernel sources.

Builds tests/emu/libpcl_emu.so from polarcode_and_ldpc_b200/csrc/*.cu[h] with
-DPCL_EMU (tests/emu/simt_emu.h) and calls the same C ABI with host numpy buffers
(the emulator's cudaMalloc is malloc).  Used by -m "not gpu" tests to check the
kernel logic against the oracle on the CPU-only build container.  The product
never loads this library.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "polarcode_and_ldpc_b200", "csrc")
LIB = os.path.join(HERE, "libpcl_emu.so")
_lib = None


def build(force=False):
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "simt_emu.h")]
    newest = max(os.path.getmtime(s) for s in srcs)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < newest:
        subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-DPCL_EMU", "-fPIC", "-shared", "-x", "c++",
                               os.path.join(CSRC, "pcl_api.cu"), "-I", HERE, "-I", CSRC, "-o", LIB])
    return LIB


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(LIB)
        _lib.pcl_last_error.restype = ctypes.c_char_p
        _lib.simt_set_reverse.argtypes = [ctypes.c_int]
    return _lib


def _vp(a):
    return ctypes.c_void_p(a.ctypes.data) if a is not None else None


def polar_decode(N, K, L, frozen_bits, llr, dtype="f64", want_pm=False, want_leaf=False,
                 crc=None, reverse=False, env=None):
    L_ = lib()
    for k, v in (env or {}).items():
        os.environ[k] = str(v)
    try:
        L_.simt_set_reverse(int(reverse))
        fm = np.zeros(N, dtype=np.uint8)
        fm[np.asarray(frozen_bits, dtype=np.int64)] = 1
        h = ctypes.c_void_p()
        crc_len, crc_poly = (0, 0) if crc is None else (crc[1], crc[0])
        rc = L_.pcl_polar_create(ctypes.byref(h), N, K, L, _vp(fm), crc_len, ctypes.c_uint32(crc_poly),
                                 1 if dtype == "f64" else 0)
        if rc:
            raise RuntimeError(f"create rc={rc}: {L_.pcl_last_error().decode()}")
        rt = np.float64 if dtype == "f64" else np.float32
        llr = np.ascontiguousarray(np.atleast_2d(llr), dtype=rt)
        F = llr.shape[0]
        LP = L_.pcl_polar_lp(h)
        fa = ctypes.c_int()
        L_.pcl_polar_launch_info(h, None, None, None, None, ctypes.byref(fa))
        polar_decode.last_fast = fa.value
        bits = np.full((F, K), 7, dtype=np.uint8)
        pm = np.zeros((F, L), dtype=np.float64) if want_pm else None
        leaf = np.zeros((F, N, LP), dtype=rt) if want_leaf else None
        par = np.zeros((F, N, LP), dtype=np.uint8) if want_leaf else None
        rc = L_.pcl_polar_decode_batch(h, _vp(llr), ctypes.c_int64(F), _vp(bits), _vp(pm), _vp(leaf), _vp(par), None)
        if rc:
            raise RuntimeError(f"decode rc={rc}: {L_.pcl_last_error().decode()}")
        L_.pcl_polar_destroy(h)
    finally:
        for k in (env or {}):
            os.environ.pop(k, None)
    out = [bits.astype(np.int64)]
    if want_pm:
        out.append(pm)
    if want_leaf:
        out.append((leaf, par))
    return out[0] if len(out) == 1 else tuple(out)


def ldpc_decode(H, llr, mode="bp", max_iter=50, normalization=1.0, early_stop=True, dtype="f64",
                want_total=False, reverse=False):
    L_ = lib()
    L_.simt_set_reverse(int(reverse))
    H8 = np.ascontiguousarray(np.asarray(H) == 1, dtype=np.uint8)
    m, n = H8.shape
    h = ctypes.c_void_p()
    rc = L_.pcl_ldpc_create(ctypes.byref(h), m, n, _vp(H8), 0 if mode == "bp" else 1,
                            ctypes.c_double(normalization), max_iter, int(early_stop), 1 if dtype == "f64" else 0)
    if rc:
        raise RuntimeError(f"create rc={rc}: {L_.pcl_last_error().decode()}")
    rt = np.float64 if dtype == "f64" else np.float32
    llr = np.ascontiguousarray(np.atleast_2d(llr), dtype=rt)
    F = llr.shape[0]
    bits = np.full((F, n), 7, dtype=np.uint8)
    iters = np.zeros(F, dtype=np.int32)
    total = np.zeros((F, n), dtype=rt) if want_total else None
    rc = L_.pcl_ldpc_decode_batch(h, _vp(llr), ctypes.c_int64(F), _vp(bits), _vp(iters), _vp(total), None)
    if rc:
        raise RuntimeError(f"decode rc={rc}: {L_.pcl_last_error().decode()}")
    L_.pcl_ldpc_destroy(h)
    out = (bits.astype(np.int64), iters)
    return out + (total,) if want_total else out


def gen_frames(kind, N, K, table, F, snr_db, seed=0, frame0=0, dtype="f32", channel=0):
    """csrc/framegen.cuh through the emulator: (llr[F, N], msg[F, K], cw[F, N]).
    table = frozen_bits (polar) or G[k, n] (ldpc)."""
    L_ = lib()
    L_.pcl_gen_frames_channel.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_uint64, ctypes.c_int,
                                          ctypes.c_double, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                          ctypes.c_void_p]
    L_.simt_set_reverse(0)
    h = ctypes.c_void_p()
    if kind == "polar":
        fm = np.zeros(N, dtype=np.uint8)
        fm[np.asarray(table, dtype=np.int64)] = 1
        rc = L_.pcl_gen_polar_create(ctypes.byref(h), N, K, _vp(fm))
    else:
        G8 = np.ascontiguousarray(np.asarray(table) % 2, dtype=np.uint8)
        rc = L_.pcl_gen_ldpc_create(ctypes.byref(h), N, K, _vp(G8))
    if rc:
        raise RuntimeError(f"gen create rc={rc}: {L_.pcl_last_error().decode()}")
    rt = np.float64 if dtype == "f64" else np.float32
    llr = np.zeros((F, N), dtype=rt)
    msg = np.full((F, K), 7, dtype=np.uint8)
    cw = np.full((F, N), 7, dtype=np.uint8)
    rc = L_.pcl_gen_frames_channel(h, F, frame0, seed, channel, float(snr_db), 1 if dtype == "f64" else 0, _vp(msg),
                                   _vp(cw), _vp(llr), None)
    if rc:
        raise RuntimeError(f"gen rc={rc}: {L_.pcl_last_error().decode()}")
    L_.pcl_gen_destroy.argtypes = [ctypes.c_void_p]
    L_.pcl_gen_destroy(h)
    return llr, msg, cw


def philox(counter, key):
    L_ = lib()
    c, k, out = np.asarray(counter, dtype=np.uint32), np.asarray(key, dtype=np.uint32), np.zeros(4, dtype=np.uint32)
    L_.pcl_philox4x32_10_host(_vp(c), _vp(k), _vp(out))
    return out
