"""On-device frame generation for BER / FER sweeps.

FrameGenerator replaces the per-frame host loop of the reference's callers
(/root/reference/benchmarks/benchmark_scl.py:95-103, test_snr_curves.py:121-130):

    message = np.random.randint(0, 2, K)
    codeword = encoder.encode(message)            # PolarEncoder / LDPCEncoder
    llr = AWGNChannel(snr_db).transmit(codeword)

with one kernel launch per batch (csrc/framegen.cuh).  Random bits and noise come from
Philox4x32-10 keyed by `seed` and addressed by the global frame index, so frame
`frame0 + f` is the same whatever the batch size, chunking or rank sharding -- a sweep
sharded over 8 GPUs sees exactly the frames a single GPU would.  The generator does not
reproduce numpy's legacy stream; `AWGNChannel.transmit_batch` (host) is the path that does.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import numpy as np

from . import _native


class FrameGenerator:
    """generate(F, snr_db, ...) -> (llr[F, N], message[F, K] uint8, codeword[F, N] uint8) on the device."""

    def __init__(self, handle, N: int, K: int, kind: str, device=None):
        self._h, self.N, self.K, self.kind, self.device = handle, N, K, kind, device

    @staticmethod
    def _device(device):
        torch = _native.require_cuda()
        return torch, (torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device))

    @classmethod
    def polar(cls, N: int, K: int, frozen_bits, device=None) -> "FrameGenerator":
        """Frames of PolarEncoder(N, K, frozen_bits) (src/polar/encoder.py:20-95, no CRC).  The code
        tables live on `device` (default: the current CUDA device); generate() runs there."""
        torch, device = cls._device(device)
        mask = np.zeros(N, dtype=np.uint8)
        mask[np.asarray(frozen_bits, dtype=np.int64)] = 1
        h = ctypes.c_void_p()
        with torch.cuda.device(device):
            _native.check(_native.lib().pcl_gen_polar_create(ctypes.byref(h), N, K, ctypes.c_void_p(mask.ctypes.data)))
        return cls(h, N, K, "polar", device)

    @classmethod
    def ldpc(cls, G, device=None) -> "FrameGenerator":
        """Frames of LDPCEncoder(G=G): codeword = message @ G mod 2 (src/ldpc/encoder.py:88-90); G is [k, n]."""
        torch, device = cls._device(device)
        G = np.ascontiguousarray(np.asarray(G) % 2, dtype=np.uint8)
        k, n = G.shape
        h = ctypes.c_void_p()
        with torch.cuda.device(device):
            _native.check(_native.lib().pcl_gen_ldpc_create(ctypes.byref(h), n, k, ctypes.c_void_p(G.ctypes.data)))
        return cls(h, n, k, "ldpc", device)

    CHANNELS = {"awgn": 0, "rayleigh": 1, "bsc": 2}

    def generate(self, F: int, snr_db: float, seed: int = 0, frame0: int = 0, dtype="float32",
                 device=None, want_codeword: bool = True, want_message: bool = True, channel: str = "awgn"):
        """channel: "awgn" (AWGNChannel(snr_db)), "rayleigh" (RayleighFadingChannel(snr_db), perfect
        channel knowledge) or "bsc" (BSCChannel(crossover_prob=snr_db): the second argument is then
        the crossover probability and the LLRs are +-ln((1 - p) / p))."""
        torch = _native.require_cuda()
        device = self.device if device is None else torch.device(device)
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        assert device == self.device, f"generator tables live on {self.device}, generate() was asked for {device}"
        code = _native.dtype_code(dtype)
        tdt = torch.float64 if code == _native.PCL_F64 else torch.float32
        llr = torch.empty((F, self.N), dtype=tdt, device=device)
        msg = torch.empty((F, self.K), dtype=torch.uint8, device=device) if want_message else None
        cw = torch.empty((F, self.N), dtype=torch.uint8, device=device) if want_codeword else None
        with torch.cuda.device(device):
            stream = torch.cuda.current_stream().cuda_stream
            _native.check(_native.lib().pcl_gen_frames_channel(
                self._h, F, frame0, ctypes.c_uint64(seed & 0xFFFFFFFFFFFFFFFF), self.CHANNELS[channel], float(snr_db), code,
                ctypes.c_void_p(msg.data_ptr()) if msg is not None and F else None,
                ctypes.c_void_p(cw.data_ptr()) if cw is not None and F else None,
                ctypes.c_void_p(llr.data_ptr()) if F else None, ctypes.c_void_p(stream)))
        return llr, msg, cw

    def close(self) -> None:
        if getattr(self, "_h", None):
            _native.lib().pcl_gen_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def philox4x32_10(counter, key) -> np.ndarray:
    """One Philox4x32-10 block from the library's own source (known-answer tests)."""
    c = np.asarray(counter, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    _native.lib().pcl_philox4x32_10_host(ctypes.c_void_p(c.ctypes.data), ctypes.c_void_p(k.ctypes.data),
                                         ctypes.c_void_p(out.ctypes.data))
    return out
