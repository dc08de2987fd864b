"""BPSK + AWGN channel with a frame-batched LLR feed.

Reference: /root/reference/src/channel/awgn.py -- AWGNChannel (:11-140):
0 -> +1, 1 -> -1 (:47), sigma = sqrt(1 / (2 snr)) (:32), LLR = 2 y / sigma^2 (:75),
noise from the legacy global np.random stream (:88).

transmit_batch(bits[F, N]) draws the noise with ONE legacy np.random.normal call of
shape (F, N); the legacy generator fills row-major and carries its cached
Box-Muller value across calls, so the result equals F successive transmit() calls
(with no other np.random use in between) bit for bit -- inputs to the GPU decoders
therefore match the reference's exactly.  The fp64 host LLRs are what gets
uploaded; decoders cast to their compute type on the device side of the copy.
"""
from __future__ import annotations

from typing import Optional

import numpy as np


class AWGNChannel:
    def __init__(self, snr_db: float, seed: Optional[int] = None):
        self.update_snr(snr_db)
        if seed is not None:
            np.random.seed(seed)

    def update_snr(self, snr_db: float) -> None:
        self.snr_db = snr_db
        self.snr_linear = 10 ** (snr_db / 10.0)
        self.noise_std = np.sqrt(1.0 / (2.0 * self.snr_linear))

    def modulate_bpsk(self, bits: np.ndarray) -> np.ndarray:
        return 1.0 - 2.0 * np.asarray(bits).astype(float)

    def demodulate_bpsk_hard(self, symbols: np.ndarray) -> np.ndarray:
        return (symbols <= 0).astype(int)

    def symbols_to_llr(self, symbols: np.ndarray) -> np.ndarray:
        return 2.0 * symbols / (self.noise_std ** 2)

    def add_noise(self, symbols: np.ndarray) -> np.ndarray:
        return symbols + np.random.normal(0, self.noise_std, symbols.shape)

    def transmit(self, bits: np.ndarray, return_llr: bool = True) -> np.ndarray:
        received = self.add_noise(self.modulate_bpsk(bits))
        return self.symbols_to_llr(received) if return_llr else self.demodulate_bpsk_hard(received)

    def transmit_batch(self, bits: np.ndarray, return_llr: bool = True, pinned: bool = False):
        """bits[F, N] -> LLR[F, N] float64 (same stream as F transmit() calls).

        pinned=True returns a page-locked torch tensor (float64) ready for an
        asynchronous host-to-device copy; the values are identical.
        """
        bits = np.asarray(bits)
        assert bits.ndim == 2, "transmit_batch expects bits[F, N]"
        out = self.transmit(bits, return_llr=return_llr)
        if not pinned:
            return out
        import torch
        t = torch.empty(out.shape, dtype=torch.float64 if return_llr else torch.int64, pin_memory=True)
        t.numpy()[...] = out
        return t

    def get_capacity(self) -> float:
        return 1.0 - np.log2(1.0 + np.exp(-self.snr_linear))

    def __repr__(self) -> str:
        return f"AWGNChannel(SNR={self.snr_db:.2f}dB, noise_std={self.noise_std:.4f})"
