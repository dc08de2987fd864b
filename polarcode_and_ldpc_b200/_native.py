"""ctypes binding of libpcl.so (include/pcl.h).  No CPU fallback: if the library
cannot be loaded, or no CUDA device is present, decoding raises."""
from __future__ import annotations

import ctypes
import os

from . import _build

PCL_OK, PCL_EINVAL, PCL_ECUDA, PCL_EUNSUPPORTED, PCL_EDEGREE1 = 0, 1, 2, 3, 4
PCL_F32, PCL_F64, PCL_F16 = 0, 1, 2
PCL_OUT_BYTES, PCL_OUT_PACKED, PCL_OUT_INT64 = 0, 1, 2
PCL_LDPC_BP, PCL_LDPC_MS = 0, 1

_lib = None


class PclError(RuntimeError):
    pass


def lib() -> ctypes.CDLL:
    """Load (building if stale and nvcc is available) the native library."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("PCL_LIB", _build.LIB)       # PCL_LIB: experiment builds of the same ABI
    if path == _build.LIB and _build.is_stale():
        try:
            _build.build_native()
        except Exception as exc:  # stale-but-present is still usable on a box without nvcc
            if not os.path.exists(path):
                raise PclError(f"libpcl.so is missing and could not be built: {exc}") from exc
    try:
        L = ctypes.CDLL(path)
    except OSError as exc:
        raise PclError(f"cannot load {path}: {exc}") from exc
    vp, i64, i32, u32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_uint32
    L.pcl_version.restype = i32
    L.pcl_last_error.restype = ctypes.c_char_p
    L.pcl_device_count.restype = i32
    L.pcl_polar_create.argtypes = [ctypes.POINTER(vp), i32, i32, i32, vp, i32, u32, i32]
    L.pcl_polar_destroy.argtypes = [vp]
    L.pcl_polar_destroy.restype = None
    L.pcl_polar_decode_batch.argtypes = [vp, vp, i64, vp, vp, vp, vp, vp]
    L.pcl_polar_decode_host.argtypes = [vp, vp, i64, vp, vp]
    L.pcl_polar_decode_host_ex.argtypes = [vp, vp, i32, i64, vp, i32, vp]
    L.pcl_ldpc_decode_host_ex.argtypes = [vp, vp, i32, i64, vp, i32, vp, vp]
    L.pcl_polar_lp.argtypes = [vp]
    L.pcl_polar_launch_info.argtypes = [vp] + [ctypes.POINTER(i32)] * 5
    L.pcl_ldpc_create.argtypes = [ctypes.POINTER(vp), i32, i32, vp, i32, ctypes.c_double, i32, i32, i32]
    L.pcl_ldpc_destroy.argtypes = [vp]
    L.pcl_ldpc_destroy.restype = None
    L.pcl_ldpc_decode_batch.argtypes = [vp, vp, i64, vp, vp, vp, vp]
    L.pcl_ldpc_decode_host.argtypes = [vp, vp, i64, vp, vp, vp]
    L.pcl_ldpc_num_edges.argtypes = [vp]
    L.pcl_ldpc_launch_info.argtypes = [vp] + [ctypes.POINTER(i32)] * 3
    L.pcl_ldpc_layout_info.argtypes = [vp] + [ctypes.POINTER(i32)] * 3
    L.pcl_count_errors.argtypes = [vp, vp, i64, i32, i32, vp, vp]
    L.pcl_gen_polar_create.argtypes = [ctypes.POINTER(vp), i32, i32, vp]
    L.pcl_gen_ldpc_create.argtypes = [ctypes.POINTER(vp), i32, i32, vp]
    L.pcl_gen_destroy.argtypes = [vp]
    L.pcl_gen_destroy.restype = None
    L.pcl_gen_frames.argtypes = [vp, i64, i64, ctypes.c_uint64, ctypes.c_double, i32, vp, vp, vp, vp]
    L.pcl_gen_frames_channel.argtypes = [vp, i64, i64, ctypes.c_uint64, i32, ctypes.c_double, i32, vp, vp, vp, vp]
    L.pcl_philox4x32_10_host.argtypes = [vp, vp, vp]
    L.pcl_philox4x32_10_host.restype = None
    _lib = L
    return L


EXPORTS = [
    "pcl_version", "pcl_last_error", "pcl_device_count",
    "pcl_polar_create", "pcl_polar_destroy", "pcl_polar_decode_batch", "pcl_polar_decode_host", "pcl_polar_decode_host_ex",
    "pcl_ldpc_decode_host_ex",
    "pcl_polar_lp", "pcl_polar_launch_info",
    "pcl_ldpc_create", "pcl_ldpc_destroy", "pcl_ldpc_decode_batch", "pcl_ldpc_decode_host",
    "pcl_ldpc_num_edges", "pcl_ldpc_launch_info", "pcl_ldpc_layout_info", "pcl_count_errors",
    "pcl_gen_polar_create", "pcl_gen_ldpc_create", "pcl_gen_destroy", "pcl_gen_frames", "pcl_gen_frames_channel", "pcl_philox4x32_10_host",
]


def check(rc: int) -> None:
    """Map status codes to the exceptions the reference raises for the same mistake."""
    if rc == PCL_OK:
        return
    msg = lib().pcl_last_error().decode(errors="replace")
    if rc == PCL_EINVAL:
        raise AssertionError(msg)
    if rc == PCL_EDEGREE1:
        raise ValueError(msg)
    if rc == PCL_EUNSUPPORTED:
        raise NotImplementedError(msg)
    raise PclError(msg)


def require_cuda():
    """torch + a CUDA device, or a loud failure (there is no CPU decode path)."""
    import torch
    if not torch.cuda.is_available() or lib().pcl_device_count() < 1:
        raise PclError("polarcode_and_ldpc_b200 needs a CUDA device (B200, sm_100a); no CPU fallback exists")
    return torch


def dtype_code(dtype) -> int:
    s = str(dtype).replace("torch.", "").replace("numpy.", "")
    if s in ("float32", "f32", "fp32", "<class 'float32'>"):
        return PCL_F32
    if s in ("float64", "f64", "fp64", "double", "<class 'float64'>"):
        return PCL_F64
    raise AssertionError(f"dtype must be float32 or float64, got {dtype!r}")


def default_dtype() -> str:
    return os.environ.get("PCL_DTYPE", "float32")


def host_llr(llr, handle_code: int):
    """numpy LLRs -> (contiguous array, PCL dtype code) for the host-buffer calls: float64 and float32
    go through as they are (the library narrows float64 for an fp32 handle), anything else is read
    as float64 like the reference does (np.asarray(llr, dtype=np.float64), decoder.py:47)."""
    import numpy as np
    a = np.asarray(llr)
    if a.dtype == np.float32 and handle_code == PCL_F32:
        return np.ascontiguousarray(a), PCL_F32
    if a.dtype == np.float16 and handle_code == PCL_F32:
        return np.ascontiguousarray(a), PCL_F16
    return np.ascontiguousarray(a, dtype=np.float64), PCL_F64
