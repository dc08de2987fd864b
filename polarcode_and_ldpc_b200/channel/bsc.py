"""Binary symmetric channel, host mirror of /root/reference/src/channel/bsc.py (BSCChannel :8-52)
with a frame-batched form; the device generator's "bsc" channel is the sweep path."""
from __future__ import annotations

from typing import Optional

import numpy as np


class BSCChannel:
    def __init__(self, crossover_prob: float, seed: Optional[int] = None):
        assert 0 <= crossover_prob <= 1, "Crossover probability must be in [0, 1]"
        self.crossover_prob = crossover_prob
        if seed is not None:
            np.random.seed(seed)

    def transmit(self, bits: np.ndarray) -> np.ndarray:
        bits = np.asarray(bits)
        flip_mask = np.random.random(bits.shape) < self.crossover_prob     # same stream as :36 for 1-D input
        output = bits.copy()
        output[flip_mask] = 1 - output[flip_mask]
        return output.astype(int)

    def transmit_batch(self, bits: np.ndarray) -> np.ndarray:
        """bits[F, N] -> received bits[F, N]; equals F successive transmit() calls."""
        bits = np.asarray(bits)
        assert bits.ndim == 2
        return self.transmit(bits)

    def bits_to_llr(self, received: np.ndarray) -> np.ndarray:
        """LLR of the received bits for the soft-input decoders: +-ln((1 - p) / p)."""
        p = self.crossover_prob
        return (1.0 - 2.0 * np.asarray(received, dtype=float)) * np.log((1.0 - p) / p)

    def __repr__(self) -> str:
        return f"BSCChannel(crossover_prob={self.crossover_prob})"
