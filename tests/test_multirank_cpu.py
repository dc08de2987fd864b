"""World-size-2 gloo test (CPU) of the multi-GPU host logic: frame sharding by rank and the
single allreduce(sum) of the [points, 4] error-counter tensor."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, F, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from polarcode_and_ldpc_b200.sweep import ErrorCounters, shard_range
    rng = np.random.default_rng(0)                      # every rank sees the same "truth"
    ref = rng.integers(0, 2, size=(F, 24), dtype=np.uint8)
    dec = ref.copy()
    dec[rng.random(ref.shape) < 0.02] ^= 1
    lo, hi = shard_range(F, rank, world)
    c = ErrorCounters(2, tensor=torch.zeros((2, 4), dtype=torch.int64))
    d = dec[lo:hi] != ref[lo:hi]                        # CPU stand-in for pcl_count_errors
    c.t[1] += torch.tensor([int(d.sum()), int(d.any(axis=1).sum()), hi - lo, (hi - lo) * 24])
    c.allreduce()
    if rank == 0:
        full = dec != ref
        exp = [int(full.sum()), int(full.any(axis=1).sum()), F, F * 24]
        out.put((c.t[1].tolist(), exp, c.t[0].tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_and_allreduce():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port, F = _free_port(), 1001                        # odd: shards differ by one frame
    procs = [ctx.Process(target=_worker, args=(r, 2, port, F, out)) for r in range(2)]
    for p in procs:
        p.start()
    got, exp, zero = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got == exp and zero == [0, 0, 0, 0]


def test_reference_import_paths_resolve_to_the_package():
    sys.path.insert(0, ROOT)
    from src.polar.decoder import SCDecoder, SCLDecoder
    from src.ldpc.decoder import BPDecoder, MSDecoder
    from src.channel.awgn import AWGNChannel
    import polarcode_and_ldpc_b200 as P
    assert SCLDecoder is P.SCLDecoder and SCDecoder is P.SCDecoder
    assert BPDecoder is P.BPDecoder and MSDecoder is P.MSDecoder and AWGNChannel is P.AWGNChannel
