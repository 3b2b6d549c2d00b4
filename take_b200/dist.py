"""Multi-GPU host logic: sample-range sharding and the one sum-reduce.

The reference's only parallelism is a thread pool over 16x16 image tiles (src/parallel.cpp:183-237,
src/render.cpp:52-82).  Samples of a pixel are i.i.d., so the natural unit across GPUs is a range of sample indices
of EVERY pixel: perfect balance, scene replicated, and -- because a sample's random stream is keyed by
(seed, pixel, sample index) -- an image that does not depend on the number of ranks (up to summation order).
The partial accumulation buffers (sum, sum of squares: H x W x 3 doubles) are combined by one all-reduce.

One process per GPU (torch.distributed, backend "nccl" on GPUs; "gloo" in the CPU tests, where the per-rank renderer
is injected).
"""
from __future__ import annotations

import numpy as np


def shard_spp(spp_begin: int, spp_end: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced split of [spp_begin, spp_end) into `world` ranges (the first `rem` ranks get one more)."""
    n = max(0, spp_end - spp_begin)
    base, rem = divmod(n, world)
    lo = spp_begin + rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def render_sharded(render_range, spp_begin: int, spp_end: int, *, reduce="all"):
    """Render [spp_begin, spp_end) across the ranks of the default torch.distributed group.

    render_range(lo, hi) -> (sum, sumsq): per-rank renderer returning torch tensors (device tensors under NCCL) or
    numpy arrays of shape [H, W, 3], float64, holding this rank's partial sums.  Returns (mean, var_of_mean, n) on
    every rank (reduce="all") or only on rank 0 (reduce="root"; None elsewhere).
    """
    import torch
    import torch.distributed as dist

    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    lo, hi = shard_spp(spp_begin, spp_end, rank, world)
    s, s2 = render_range(lo, hi)
    as_t = lambda a: a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))
    s, s2 = as_t(s), as_t(s2)
    if world > 1:
        if reduce == "all":
            dist.all_reduce(s)      # sum over ranks: the path's only exchange step
            dist.all_reduce(s2)
        else:
            dist.reduce(s, dst=0)
            dist.reduce(s2, dst=0)
            if rank != 0:
                return None
    n = spp_end - spp_begin
    return finalize(s, s2, n)


def finalize(s, s2, n: int):
    """Epilogue after the reduce: mean = sum / spp (src/render.cpp:78) and the variance of that mean."""
    mean = s / n
    if n > 1:
        var = (s2 / n - mean * mean).clamp_min(0) * (n / (n - 1)) / n if hasattr(s, "clamp_min") else \
            np.maximum(s2 / n - mean * mean, 0) * (n / (n - 1)) / n
    else:
        var = mean * 0
    return mean, var, n


def shared_host_build(flat, directory: str = None):
    """One host build per NODE instead of one per rank: local rank 0 builds the acceleration structures and saves them
    into a fresh private directory (mkdtemp, mode 0700, under /dev/shm when it exists), the path and a success flag are
    broadcast, and the node's other ranks load the file.  A failure on the building rank is raised on EVERY rank (no rank
    is left waiting at a barrier).  Returns an `api.HostBuild` to pass as `GpuScene(flat, device=local_rank, prebuilt=...)`.
    Without an initialised process group it just builds.  (All ranks of the group are assumed to share one node's
    file system -- the single-node launch bench.py and the driver use.)"""
    import os
    import shutil
    import tempfile
    import torch.distributed as dist
    from . import api
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return api.HostBuild(flat)
    rank = dist.get_rank()
    hb, tmpdir, msg = None, None, [None, None]       # msg = [path or None, error text or None]
    if rank == 0:
        try:
            base = directory or ("/dev/shm" if os.path.isdir("/dev/shm") else None)
            tmpdir = tempfile.mkdtemp(prefix="take_hostbuild_", dir=base)
            hb = api.HostBuild(flat)
            path = os.path.join(tmpdir, "build.bin")
            hb.save(path)
            msg = [path, None]
        except Exception as ex:                       # reported to everybody below
            msg = [None, f"{type(ex).__name__}: {ex}"]
    dist.broadcast_object_list(msg, src=0)
    err = msg[1]
    if err is None and rank != 0:
        try:
            hb = api.HostBuild(path=msg[0])
        except Exception as ex:
            err = f"rank {rank}: {type(ex).__name__}: {ex}"
    errs = [None] * dist.get_world_size()
    dist.all_gather_object(errs, err)                 # also the point after which the file is no longer needed
    if tmpdir is not None:
        shutil.rmtree(tmpdir, ignore_errors=True)
    errs = [e for e in errs if e]
    if errs:
        if hb is not None:
            hb.close()
        raise RuntimeError("shared_host_build failed: " + "; ".join(sorted(set(errs))))
    return hb
