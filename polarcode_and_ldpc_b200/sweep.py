"""Frame-sharded error counting for multi-GPU BER/FER sweeps.

Frames are independent in every decoder, so a sweep shards by frame batch: rank r
of R decodes frames [r F/R, (r+1) F/R) with replicated code tables and NO
data-path collective.  The only exchange is one allreduce(sum) of the int64
[points, 4] counter tensor (bit errors, frame errors, frames, bits) -- the
np.sum(message != decoded) bookkeeping of the reference's callers
(/root/reference/benchmarks/test_snr_curves.py:133-138), done on the device by
pcl_count_errors writing straight into the tensor that is reduced.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

from . import _native


def shard_range(F: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of F frames for `rank` of `world` (sizes differ by <= 1)."""
    assert 0 <= rank < world
    base, rem = divmod(F, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def count_errors(bits, ref, ncmp: Optional[int] = None, out=None):
    """Add (bit errors, frame errors, frames, bits) of uint8 CUDA tensors bits/ref[F, W]
    to `out` (int64[4] CUDA tensor, created zeroed if None); compares the first ncmp columns."""
    torch = _native.require_cuda()
    assert bits.is_cuda and ref.is_cuda and bits.dtype == torch.uint8 and ref.dtype == torch.uint8
    assert bits.shape == ref.shape and bits.dim() == 2
    bits, ref = bits.contiguous(), ref.contiguous()
    F, W = bits.shape
    ncmp = W if ncmp is None else int(ncmp)
    if out is None:
        out = torch.zeros(4, dtype=torch.int64, device=bits.device)
    assert out.is_cuda and out.dtype == torch.int64 and out.numel() == 4 and out.is_contiguous()
    with torch.cuda.device(bits.device):
        stream = torch.cuda.current_stream().cuda_stream
        _native.check(_native.lib().pcl_count_errors(
            ctypes.c_void_p(bits.data_ptr()), ctypes.c_void_p(ref.data_ptr()), F, W, ncmp,
            ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(stream)))
    return out


class ErrorCounters:
    """[points, 4] int64 counters; one allreduce at the end of a sweep."""

    COLUMNS = ("bit_errors", "frame_errors", "frames", "bits")

    def __init__(self, points: int, device=None, tensor=None):
        import torch
        self.t = tensor if tensor is not None else torch.zeros((points, 4), dtype=torch.int64, device=device)

    def add(self, point: int, bits, ref, ncmp: Optional[int] = None):
        count_errors(bits, ref, ncmp, out=self.t[point])

    def allreduce(self):
        """Sum over ranks (NCCL on GPUs, gloo for CPU tests); a no-op without a process group."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.t, op=dist.ReduceOp.SUM)
        return self

    def rates(self):
        t = self.t.cpu().double()
        ber = t[:, 0] / t[:, 3].clamp(min=1)
        fer = t[:, 1] / t[:, 2].clamp(min=1)
        return ber.numpy(), fer.numpy()
