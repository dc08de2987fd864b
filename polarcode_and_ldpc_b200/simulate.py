"""Batched Monte-Carlo loop shared by the re-pointed benchmark callers.

Every BER / FER script of the reference runs the same per-frame loop on the host
(/root/reference/benchmarks/test_snr_curves.py:101-144, ber_simulation.py:160-205,
sc_vs_scl.py:281-325, test_code_parameters.py:87-111):

    message -> encoder.encode -> channel.transmit -> decoder.decode -> np.sum(message != decoded)
    ... stop when frame_errors >= max_errors

Here a chunk of frames is generated on the device (FrameGenerator), decoded with one
decode_batch call per decoder and counted on the device (pcl_count_errors); the
`max_errors` stop is applied between chunks (chunks start small and double), so a point
tests at most one chunk more than the reference would.  With torch.distributed initialised
the frames of a chunk are sharded over the ranks by frame index and the counters are
all-reduced once per chunk (the only collective).
"""
from __future__ import annotations

import time
from typing import Dict, Optional

import numpy as np

from . import _native
from .framegen import FrameGenerator
from .ldpc.construction import gallager_parity_check, generator_from_parity
from .polar.construction import bhattacharyya_frozen_set
from .sweep import count_errors, shard_range


def make_polar_code(N: int, K: int, design_snr_db: float = 2.0) -> dict:
    """Frozen set as the reference's callers obtain it (PolarLibWrapper(N, K, 2.0), restated
    library-free: polar/construction.py) + a frame generator for it."""
    frozen = bhattacharyya_frozen_set(N, K, design_snr_db)
    return {"type": "polar", "N": N, "K": K, "frozen_bits": frozen, "gen": FrameGenerator.polar(N, K, frozen)}


def make_ldpc_code(n: int, dv: int = 3, dc: int = 6, seed: int = 42) -> dict:
    """H, G as the reference's callers obtain them (LDPCLibWrapper(n, k, dv, dc, seed): Gallager
    H, restated library-free in ldpc/construction.py; K becomes the true dimension of the code,
    test_snr_curves.py:68-69).  G is systematic on `info_positions`."""
    H = gallager_parity_check(n, dv, dc, seed)
    G, info = generator_from_parity(H)
    return {"type": "ldpc", "N": n, "K": G.shape[0], "H": H, "G": G, "info_positions": info,
            "gen": FrameGenerator.ldpc(G)}


def simulate_point(code: dict, decoders: Dict[str, object], snr_db: float, num_frames: int,
                   max_errors: Optional[int] = None, seed: int = 0, first_chunk: int = 4096,
                   max_chunk: int = 65536, dtype="float32") -> Dict[str, dict]:
    """One SNR point for several decoders on the SAME frames.  Returns per decoder
    {frames_tested, total_bits, error_bits, frame_errors, decode_seconds} plus '_gen_seconds'."""
    torch = _native.require_cuda()
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    dev = torch.device("cuda", torch.cuda.current_device())
    gen: FrameGenerator = code["gen"]
    K = code["K"]
    info_idx = None
    if code["type"] == "ldpc":
        info_idx = torch.from_numpy(np.asarray(code["info_positions"], dtype=np.int64)).to(dev)
    counters = {name: torch.zeros(4, dtype=torch.int64, device=dev) for name in decoders}
    secs = {name: 0.0 for name in decoders}
    gen_s = 0.0
    done, chunk = 0, min(first_chunk, num_frames)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    while done < num_frames:
        c = min(chunk, num_frames - done)
        lo, hi = shard_range(c, rank, world)
        t0 = time.time()
        llr, msg, _ = gen.generate(hi - lo, snr_db, seed=seed, frame0=done + lo, dtype=dtype, device=dev,
                                   want_codeword=False)
        torch.cuda.synchronize()
        gen_s += time.time() - t0
        for name, dec in decoders.items():
            ev[0].record()
            bits = dec.decode_batch(llr)
            ev[1].record()
            if info_idx is not None:                       # callers compare decoded[:K] with the message
                bits = bits.index_select(1, info_idx).contiguous()
            count_errors(bits, msg, out=counters[name])
            torch.cuda.synchronize()
            secs[name] += ev[0].elapsed_time(ev[1]) * 1e-3
        done += c
        chunk = min(2 * chunk, max_chunk)
        if max_errors is not None:
            tot = torch.stack([counters[n] for n in decoders])
            if world > 1:
                tot = tot.clone()
                dist.all_reduce(tot, op=dist.ReduceOp.SUM)
            if bool((tot[:, 1] >= max_errors).all().item()):
                break
    out = {}
    for name in decoders:
        t = counters[name]
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        e_bits, e_frames, frames, bits_total = (int(x) for x in t.tolist())
        out[name] = {"frames_tested": frames, "total_bits": bits_total, "error_bits": e_bits,
                     "frame_errors": e_frames, "decode_seconds": secs[name],
                     "ber": e_bits / bits_total if bits_total else 0.0,
                     "fer": e_frames / frames if frames else 0.0}
    out["_gen_seconds"] = gen_s
    return out
