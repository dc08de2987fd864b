"""Rayleigh fading channel, host mirror of /root/reference/src/channel/fading.py
(RayleighFadingChannel :8-66: |h| s + n with perfect channel knowledge, LLR = 2 y |h| / sigma^2)
with a frame-batched form; the device generator's "rayleigh" channel is the sweep path."""
from __future__ import annotations

from typing import Optional

import numpy as np


class RayleighFadingChannel:
    def __init__(self, snr_db: float, seed: Optional[int] = None):
        self.snr_db = snr_db
        self.snr_linear = 10 ** (snr_db / 10.0)
        self.noise_std = np.sqrt(1.0 / (2.0 * self.snr_linear))
        if seed is not None:
            np.random.seed(seed)

    def transmit(self, bits: np.ndarray, return_llr: bool = True) -> np.ndarray:
        bits = np.asarray(bits)
        symbols = 1.0 - 2.0 * bits.astype(float)
        h_real = np.random.normal(0, 1 / np.sqrt(2), symbols.shape)
        h_imag = np.random.normal(0, 1 / np.sqrt(2), symbols.shape)
        h_mag = np.abs(h_real + 1j * h_imag)
        received = h_mag * symbols + np.random.normal(0, self.noise_std, symbols.shape)
        if return_llr:
            return 2.0 * received * h_mag / (self.noise_std ** 2)
        return (received <= 0).astype(int)

    def transmit_batch(self, bits: np.ndarray, return_llr: bool = True) -> np.ndarray:
        """bits[F, N] -> LLR[F, N].  Draw order differs from F transmit() calls (three arrays of
        shape (F, N) instead of 3 F arrays of length N); the distribution is the same."""
        bits = np.asarray(bits)
        assert bits.ndim == 2
        return self.transmit(bits, return_llr)

    def __repr__(self) -> str:
        return f"RayleighFadingChannel(SNR={self.snr_db:.2f}dB)"
