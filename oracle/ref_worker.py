"""The reference's own numpy decoders, timed on the host cores (bench.py cpu_baseline / --impl reference).

TEST / MEASUREMENT INFRASTRUCTURE ONLY.  `stage()` copies /root/reference/src/{polar,ldpc,channel}
(pure Python, numpy / scipy only) unmodified into baseline/_ref/refsrc/ -- git-ignored, NOT
gpurun-ignored, so it travels to the GPU box with the snapshot.  The worker
functions import the reference from there and run its decode(llr) frame by frame:
    SCDecoder.decode   /root/reference/src/polar/decoder.py:38
    SCLDecoder.decode  /root/reference/src/polar/decoder.py:225
    BPDecoder.decode   /root/reference/src/ldpc/decoder.py:124
    MSDecoder.decode   /root/reference/src/ldpc/decoder.py:289
Nothing in polarcode_and_ldpc_b200/ imports this module.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import shutil
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SRC = "/root/reference/src"
STAGED = os.path.join(ROOT, "baseline", "_ref", "refsrc")
PACKAGES = ("polar", "ldpc", "channel")


def stage() -> bool:
    """Copy the reference's three pure-Python packages next to the repo (idempotent).  Returns
    whether a staged copy exists afterwards."""
    if os.path.isdir(REF_SRC):
        for p in PACKAGES:
            src, dst = os.path.join(REF_SRC, p), os.path.join(STAGED, p)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copytree(src, dst, dirs_exist_ok=True, ignore=shutil.ignore_patterns("__pycache__"))
    return available()


def available() -> bool:
    return all(os.path.isfile(os.path.join(STAGED, p, "__init__.py")) for p in PACKAGES)


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_model() -> str:
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


# ---- worker side (spawned processes) ---------------------------------------------------
def _import_reference():
    if STAGED not in sys.path:
        sys.path.insert(0, STAGED)
    import polar.decoder as pd      # noqa: E402  (the reference's module, from baseline/_ref/refsrc)
    import ldpc.decoder as ld       # noqa: E402
    assert os.path.realpath(pd.__file__).startswith(os.path.realpath(STAGED)), pd.__file__
    return pd, ld


def _warm(_):
    _import_reference()
    return os.getpid()


def _decode_shard(job):
    """job = (spec, llr[f, N]); returns (bits[f, *], seconds spent decoding, iterations or None)."""
    spec, llr = job
    pd, ld = _import_reference()
    kind = spec["kind"]
    if kind == "sc":
        dec = pd.SCDecoder(spec["N"], spec["K"], frozen_bits=np.asarray(spec["frozen"]))
    elif kind == "scl":
        dec = pd.SCLDecoder(spec["N"], spec["K"], list_size=spec["L"], frozen_bits=np.asarray(spec["frozen"]))
    elif kind == "bp":
        dec = ld.BPDecoder(np.asarray(spec["H"]), max_iter=spec["iters"], early_stop=spec["early_stop"])
    else:
        dec = ld.MSDecoder(np.asarray(spec["H"]), max_iter=spec["iters"], normalization=spec["normalization"],
                           early_stop=spec["early_stop"])
    out = []
    t0 = time.perf_counter()
    for f in range(llr.shape[0]):
        out.append(np.asarray(dec.decode(llr[f])))
    dt = time.perf_counter() - t0
    return np.stack(out) if out else np.zeros((0, 0), dtype=np.int64), dt


class ReferencePool:
    """multiprocessing.Pool(nproc) over frame shards, one reference decoder object per shard."""

    def __init__(self, procs: int | None = None):
        if not available():
            raise RuntimeError("reference sources are not staged under baseline/_ref/refsrc (run __graft_entry__.build() "
                               "in the build container)")
        self.procs = procs or host_cores()
        self.pool = mp.get_context("spawn").Pool(self.procs)
        self.pool.map(_warm, range(self.procs * 2))

    def decode(self, spec: dict, llr: np.ndarray):
        """Shard llr[F, N] over the workers.  Returns (bits, wall seconds, summed per-worker decode seconds)."""
        F = llr.shape[0]
        shards = [s for s in np.array_split(np.arange(F), self.procs) if len(s)]
        jobs = [(spec, np.ascontiguousarray(llr[s[0]:s[-1] + 1], dtype=np.float64)) for s in shards]
        t0 = time.perf_counter()
        res = self.pool.map(_decode_shard, jobs, chunksize=1)
        wall = time.perf_counter() - t0
        bits = np.concatenate([r[0] for r in res], axis=0)
        return bits, wall, float(sum(r[1] for r in res))

    def close(self):
        self.pool.close()
        self.pool.join()
