"""Drop-in SCDecoder / SCLDecoder backed by the sm_100a kernels.

Mirrors /root/reference/src/polar/decoder.py: same constructors
(SCDecoder :16, SCLDecoder :191-193), same attributes (:20-36, :198-223), same
decode(llr) contract (:38-71, :225-262: any length-N array-like in, np.int64[K]
info bits in ascending index order out, AssertionError on a bad shape), plus the
batched entry point decode_batch(llr[F, N]) whose row f equals decode(llr[f]).

Differences, all additive:
  * dtype= ("float32" production / "float64" validation build) and device=.
  * use_crc=True performs CRC-aided selection on the device.  The reference
    stores the flag but never applies it (decoder.py:202-203, 259), so this rule
    has no reference behaviour to match ("parity unpinned"); with use_crc=False
    the result is the reference's.
  * list sizes up to 32 run one path per lane (a warp carries 32 / L frames); 33 .. 1024 run one
    block per frame with a thread per slot (polar_scl_wide.cuh); N is limited to 65536 (16-bit leaf positions; beyond 8192 as far as a
    frame's bit arrays fit one block's shared memory: 32768 for the small lists) and list_size
    to 1024 (NotImplementedError beyond; per-leaf dumps stop at list size 256).
There is no CPU path: constructing a decoder without a CUDA device raises.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import numpy as np

from .. import _native
from .utils import crc_resolve, generate_frozen_bits


class _PolarBase:
    def _setup(self, N: int, K: int, list_size: int, frozen_bits, use_crc: bool, crc_polynomial: str,
               dtype, device):
        assert N > 0 and (N & (N - 1)) == 0, "N must be a power of 2"
        assert 0 < K < N, "K must be in (0, N)"
        assert list_size >= 1
        self.N, self.K = N, K
        self.n = int(np.log2(N))
        if frozen_bits is None:
            self.frozen_bits, self.info_bits = generate_frozen_bits(N, K)
        else:
            self.frozen_bits = np.array(frozen_bits, dtype=int)
            self.info_bits = np.setdiff1d(np.arange(N), self.frozen_bits)
        self.frozen_set = set(self.frozen_bits)
        self._list_size = list_size
        self._torch = _native.require_cuda()
        torch = self._torch
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.dtype_name = str(dtype or _native.default_dtype()).replace("torch.", "")
        self._code = _native.dtype_code(self.dtype_name)
        self._tdtype = torch.float64 if self._code == _native.PCL_F64 else torch.float32
        mask = np.zeros(N, dtype=np.uint8)
        mask[self.frozen_bits] = 1
        crc_poly, crc_len = crc_resolve(crc_polynomial) if use_crc else (0, 0)
        self._h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _native.check(_native.lib().pcl_polar_create(
                ctypes.byref(self._h), N, len(self.info_bits), list_size,
                ctypes.c_void_p(mask.ctypes.data), crc_len, ctypes.c_uint32(crc_poly), self._code))
        self._LP = _native.lib().pcl_polar_lp(self._h)
        self._K_out = len(self.info_bits)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                _native.lib().pcl_polar_destroy(h)
            except Exception:
                pass
            self._h = None

    # -- plumbing --------------------------------------------------------------
    def _to_device(self, llr):
        torch = self._torch
        if isinstance(llr, torch.Tensor):
            t = llr
        else:
            t = torch.from_numpy(np.ascontiguousarray(np.asarray(llr, dtype=np.float64)))
        assert t.dim() == 2 and t.shape[1] == self.N, f"expected LLR shape (F, {self.N}), got {tuple(t.shape)}"
        return t.to(device=self.device, non_blocking=True).to(self._tdtype).contiguous()

    def _run(self, llr_dev, want_pm: bool, want_leaf: bool):
        torch = self._torch
        F = llr_dev.shape[0]
        bits = torch.empty((F, self._K_out), dtype=torch.uint8, device=self.device)
        pm = torch.empty((F, self._list_size), dtype=torch.float64, device=self.device) if want_pm else None
        leaf = par = None
        if want_leaf:
            leaf = torch.empty((F, self.N, self._LP), dtype=self._tdtype, device=self.device)
            par = torch.empty((F, self.N, self._LP), dtype=torch.uint8, device=self.device)
        ptr = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None  # noqa: E731
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream().cuda_stream
            _native.check(_native.lib().pcl_polar_decode_batch(
                self._h, ptr(llr_dev), F, ptr(bits), ptr(pm), ptr(leaf), ptr(par), ctypes.c_void_p(stream)))
        return bits, pm, leaf, par

    def _leaf_of_best(self, pm, leaf, par):
        """L_paths[best, :, n] (decoder.py:259-260): walk the survivor parents back."""
        torch = self._torch
        F = leaf.shape[0]
        cur = torch.argmax(pm, dim=1) if pm is not None else torch.zeros(F, dtype=torch.long, device=self.device)
        out = torch.empty((F, self.N), dtype=self._tdtype, device=self.device)
        ar = torch.arange(F, device=self.device)
        br = self._bitrev()
        for i in range(self.N - 1, -1, -1):
            cur = par[ar, i, cur].long()
            out[:, br[i]] = leaf[ar, i, cur]
        return out

    def _bitrev(self):
        from .utils import bit_reverse_permutation
        return bit_reverse_permutation(self.n)

    def launch_info(self) -> dict:
        g, b, s, lv, fa = (ctypes.c_int() for _ in range(5))
        _native.check(_native.lib().pcl_polar_launch_info(self._h, ctypes.byref(g), ctypes.byref(b),
                                                          ctypes.byref(s), ctypes.byref(lv), ctypes.byref(fa)))
        if fa.value == 4:
            M = self.N // 256
            wpb = {2: 6, 8: 5, 16: 4}.get(M, 3)
            return {"grid": g.value, "block": 32 * wpb, "smem_bytes": wpb * 32 * (260 + (8 * (M - 1) if M > 1 else 0)) * 4, "global_levels": 0,
                    "kernel": "polar_sc256_kernel" if M == 1 else "polar_sc_big_kernel",
                    "lanes_per_path": 1, "frames_per_warp": 32, "compiled_code_length": True, "tensor_memory": False}
        if fa.value == 5:
            return {"grid": g.value, "block": b.value, "smem_bytes": s.value, "global_levels": lv.value,
                    "kernel": "polar_scl_wide_kernel", "lanes_per_path": 1, "frames_per_block": 1,
                    "compiled_code_length": False, "tensor_memory": False}
        return {"grid": g.value, "block": b.value, "smem_bytes": s.value, "global_levels": lv.value,
                "kernel": "polar_scl_fast_kernel" if fa.value else "polar_scl_kernel",
                "lanes_per_path": 1 if fa.value else 32 // self._LP,
                "frames_per_warp": (32 // self._LP) if fa.value else 1,
                "compiled_code_length": fa.value >= 2, "tensor_memory": fa.value == 3}

    def _decode_numpy(self, llr) -> np.ndarray:
        """The reference call shape: host float array [F, N] in, np.int64 [F, K] out, through the
        library's chunked host pipeline (pcl_polar_decode_host_ex: conversion threads, pinned
        staging, H2D, decode, packed D2H and unpacking overlap)."""
        a, code = _native.host_llr(llr, self._code)
        assert a.ndim == 2 and a.shape[1] == self.N, f"expected LLR shape (F, {self.N}), got {a.shape}"
        out = np.empty((a.shape[0], self._K_out), dtype=np.int64)
        torch = self._torch
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream().cuda_stream
            _native.check(_native.lib().pcl_polar_decode_host_ex(
                self._h, ctypes.c_void_p(a.ctypes.data), code, a.shape[0], ctypes.c_void_p(out.ctypes.data),
                _native.PCL_OUT_INT64, ctypes.c_void_p(stream)))
        return out

    def decode_batch_host(self, llr_host, bits_host=None, packed: bool = False):
        """C-ABI host-buffer path: llr_host is a CPU tensor/array [F, N] in the compute dtype, or
        float16 as an opt-in TRANSPORT format for an fp32 decoder (half the PCIe bytes; the decoder then
        sees the fp16-rounded LLRs).  Pinned memory makes the copies asynchronous.  Returns uint8 [F, K]
        on host, or with packed=True the bit-packed rows int32 [F, ceil(K / 32)] (bit k of a row at word
        k // 32, bit k % 32; packed on the device, 1/8 of the D2H bytes).  H2D, decode and D2H are
        chunked and overlapped inside the library."""
        torch = self._torch
        if not isinstance(llr_host, torch.Tensor):
            llr_host = torch.from_numpy(np.ascontiguousarray(llr_host))
        f16 = llr_host.dtype == torch.float16 and self._code == _native.PCL_F32
        assert llr_host.dim() == 2 and llr_host.shape[1] == self.N and (llr_host.dtype == self._tdtype or f16)
        assert llr_host.device.type == "cpu" and llr_host.is_contiguous()
        F = llr_host.shape[0]
        if packed or f16:
            words = (self._K_out + 31) // 32
            shape, dt, fmt = ((F, words), torch.int32, _native.PCL_OUT_PACKED) if packed else \
                             ((F, self._K_out), torch.uint8, _native.PCL_OUT_BYTES)
            if bits_host is None:
                bits_host = torch.empty(shape, dtype=dt, pin_memory=True)
            assert isinstance(bits_host, torch.Tensor) and bits_host.device.type == "cpu" and bits_host.dtype == dt \
                and bits_host.is_contiguous() and tuple(bits_host.shape) == shape, \
                f"bits_host must be a contiguous CPU {dt} tensor of shape {shape}"
            with torch.cuda.device(self.device):
                stream = torch.cuda.current_stream().cuda_stream
                _native.check(_native.lib().pcl_polar_decode_host_ex(
                    self._h, ctypes.c_void_p(llr_host.data_ptr()), _native.PCL_F16 if f16 else self._code, F,
                    ctypes.c_void_p(bits_host.data_ptr()), fmt, ctypes.c_void_p(stream)))
            return bits_host
        if bits_host is None:
            bits_host = torch.empty((F, self._K_out), dtype=torch.uint8, pin_memory=True)
        # the library writes F * K bytes through this pointer: refuse anything that is not exactly that
        assert isinstance(bits_host, torch.Tensor) and bits_host.device.type == "cpu" and bits_host.dtype == torch.uint8 \
            and bits_host.is_contiguous() and tuple(bits_host.shape) == (F, self._K_out), \
            f"bits_host must be a contiguous CPU uint8 tensor of shape ({F}, {self._K_out})"
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream().cuda_stream
            _native.check(_native.lib().pcl_polar_decode_host(
                self._h, ctypes.c_void_p(llr_host.data_ptr()), F, ctypes.c_void_p(bits_host.data_ptr()),
                ctypes.c_void_p(stream)))
        return bits_host


class SCDecoder(_PolarBase):
    """Successive-cancellation decoder (reference: src/polar/decoder.py:12-173)."""

    def __init__(self, N: int, K: int, frozen_bits: Optional[np.ndarray] = None, dtype=None, device=None):
        self._setup(N, K, 1, frozen_bits, False, "CRC-8", dtype, device)
        # the reference's LLR / bit matrices (decoder.py:35-36).  Callers read the channel column and
        # the leaf column (debug_scripts/compare_step_by_step.py:40-47); they are filled LAZILY, on the
        # first read of .L / .B after a decode(), so decode() itself is one kernel launch.
        self._L = np.full((N, self.n + 1), np.nan, dtype=np.float64)
        self._B = np.full((N, self.n + 1), np.nan, dtype=np.float64)
        self._pending = None                    # (llr_input, decoded bits) of the last decode()

    def _materialise(self):
        if self._pending is None:
            return
        llr_input, out = self._pending
        self._pending = None
        _, _, leaf, par = self._run(self._to_device(llr_input[None, :]), False, True)
        self._L[:, 0] = llr_input
        self._L[:, self.n] = self._leaf_of_best(None, leaf, par)[0].double().cpu().numpy()
        u = np.zeros(self.N)
        u[self.info_bits] = out
        self._B[:, self.n] = u

    @property
    def L(self) -> np.ndarray:
        self._materialise()
        return self._L

    @property
    def B(self) -> np.ndarray:
        self._materialise()
        return self._B

    def decode(self, llr_input: np.ndarray) -> np.ndarray:
        llr_input = np.asarray(llr_input, dtype=np.float64)
        assert llr_input.shape == (self.N,), f"expected LLR shape ({self.N},), got {llr_input.shape}"
        bits, _, _, _ = self._run(self._to_device(llr_input[None, :]), False, False)
        out = bits[0].cpu().numpy().astype(np.int64)
        self._pending = (llr_input.copy(), out)
        return out

    def decode_batch(self, llr, return_leaf_llr: bool = False):
        """llr[F, N] (numpy / torch, host or device) -> info bits [F, K].

        numpy in -> np.int64 out (what F decode() calls would return); CUDA tensor in ->
        uint8 CUDA tensor out, no host sync.  return_leaf_llr adds L[:, n] per frame."""
        on_device = isinstance(llr, self._torch.Tensor) and llr.is_cuda
        if not return_leaf_llr and not isinstance(llr, self._torch.Tensor):
            return self._decode_numpy(llr)
        bits, _, leaf, par = self._run(self._to_device(llr), False, return_leaf_llr)
        res = bits if on_device else bits.cpu().numpy().astype(np.int64)
        if return_leaf_llr:
            lf = self._leaf_of_best(None, leaf, par)
            return res, (lf if on_device else lf.double().cpu().numpy())
        return res

    def __repr__(self) -> str:
        return f"SCDecoder(N={self.N}, K={self.K})"


class SCLDecoder(_PolarBase):
    """Successive-cancellation list decoder (reference: src/polar/decoder.py:176-444)."""

    def __init__(self, N: int, K: int, list_size: int = 8, frozen_bits: Optional[np.ndarray] = None,
                 use_crc: bool = False, crc_polynomial: str = "CRC-8", dtype=None, device=None):
        assert list_size >= 1
        self.use_crc = use_crc
        self.crc_polynomial = crc_polynomial
        self._setup(N, K, list_size, frozen_bits, use_crc, crc_polynomial, dtype, device)
        self.L = list_size                      # decoder.py:200 -- L is the list size here
        self.path_metrics = np.full(self.L, -np.inf)
        self.active_paths = np.zeros(self.L, dtype=bool)

    def decode(self, llr_input: np.ndarray) -> np.ndarray:
        llr_input = np.asarray(llr_input, dtype=np.float64)
        assert llr_input.shape == (self.N,), f"expected LLR shape ({self.N},), got {llr_input.shape}"
        bits, pm, _, _ = self._run(self._to_device(llr_input[None, :]), True, False)
        self.path_metrics = pm[0].cpu().numpy()
        self.active_paths = np.isfinite(self.path_metrics)
        return bits[0].cpu().numpy().astype(np.int64)

    def decode_batch(self, llr, return_path_metrics: bool = False, return_leaf_llr: bool = False):
        """llr[F, N] -> info bits [F, K] (+ path_metrics[F, L], + L_paths[best, :, n])."""
        on_device = isinstance(llr, self._torch.Tensor) and llr.is_cuda
        want_pm = return_path_metrics or return_leaf_llr
        if not want_pm and not isinstance(llr, self._torch.Tensor):
            return self._decode_numpy(llr)
        bits, pm, leaf, par = self._run(self._to_device(llr), want_pm, return_leaf_llr)
        out = [bits if on_device else bits.cpu().numpy().astype(np.int64)]
        if return_path_metrics:
            out.append(pm if on_device else pm.cpu().numpy())
        if return_leaf_llr:
            lf = self._leaf_of_best(pm, leaf, par)
            out.append(lf if on_device else lf.double().cpu().numpy())
        return out[0] if len(out) == 1 else tuple(out)

    def __repr__(self) -> str:
        return f"SCLDecoder(N={self.N}, K={self.K}, L={self.L}, use_crc={self.use_crc})"
