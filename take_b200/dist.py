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


def configure_nccl_for_overlap(channels: int = 4) -> None:
    """Call BEFORE torch.distributed.init_process_group("nccl").  ShardedRenderer's per-job reduce (2 x W x H x 3 doubles, 100 MB
    at 1080p) runs on a side stream next to the following job's kernels: the transfer needs a millisecond of a 15 ms job, so
    what matters is how many SMs NCCL's kernels hold while they wait for the slowest rank, not the link bandwidth they
    reach.  Four channels instead of NCCL's default (measured on 8 x B200, config 2, per-step reduce): 15.84 -> 15.43 ms per
    step, weak-scaling efficiency 95.8 -> 98.2 %.  An NCCL_MAX_NCHANNELS set by the user is left alone."""
    import os
    os.environ.setdefault("NCCL_MAX_NCHANNELS", str(int(channels)))


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


class ShardedJob:
    """One submitted sample range: `wait()` blocks until this rank's part of it (render, reduce, read-back) is complete
    and returns (stats, sum, sumsq): the host arrays on rank 0 when a read-back was asked for, else the device tensors
    (reduced on rank 0, partial elsewhere).  Both are VIEWS of one of the renderer's two buffer sets: valid until the
    submit after next reuses that set (copy what must live longer)."""

    def __init__(self, stats, d_sum, d_sq, h_sum, h_sq, done):
        self.stats, self.d_sum, self.d_sq, self.h_sum, self.h_sq, self.done = stats, d_sum, d_sq, h_sum, h_sq, done
        self.collected = False

    def wait(self):
        self.done.synchronize()
        self.collected = True
        return self.stats, (self.h_sum if self.h_sum is not None else self.d_sum), (self.h_sq if self.h_sq is not None else self.d_sq)


class _HostEngine:
    """Stand-ins for the CUDA pieces when ShardedRenderer runs on CPU tensors (the gloo tests): streams are the host's
    program order, events are already complete."""

    class _Ctx:
        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

    class _Event:
        def record(self, *_):
            pass

        def synchronize(self):
            pass

        def elapsed_time(self, other):
            return 0.0

    def stream(self, _):
        return self._Ctx()

    def event(self):
        return self._Event()


class ShardedRenderer:
    """The one-process-per-GPU render path (torchrun / the driver's launch): what `parallel_for` over image tiles
    (src/parallel.cpp:183-237) becomes across GPUs.

    * scene replicated: the host-side trees are built ONCE per node (`shared_host_build`) and every rank creates its
      `GpuScene` from that build;
    * `submit(lo, hi)`: this rank renders its share (`shard_spp`) of the sample-index range [lo, hi) of every pixel into
      one of two device buffer sets, then -- on a side stream, so that it overlaps the NEXT submit's kernels -- the
      partial sums are combined on rank 0 with one NCCL sum-reduce per buffer (the path's only exchange step) and, when
      `to_host` is set, rank 0 copies the reduced image into pinned host memory (ONE device->host copy per job; the
      other ranks copy nothing);
    * two jobs may be in flight; results depend only on (seed, pixel, sample index), so the reduced image equals the
      single-GPU image up to summation order.
    Works without a process group (world = 1: no reduce).

    `render_fn(sum, sumsq_or_None, integrator, max_depth, lo, hi, seed, flags) -> stats dict` replaces the GPU scene in the
    CPU tests (gloo backend, CPU tensors): the sharding, double buffering and reduce are the code under test there."""

    def __init__(self, flat, device=None, sumsq: bool = True, render_fn=None, share_host_build: bool = False):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.flat, self.sumsq = flat, sumsq
        H, W = flat.height, flat.width
        self.render_fn = render_fn
        if render_fn is None:
            from . import api
            import os
            self.device = torch.device(f"cuda:{device}")
            if share_host_build:
                # host-built trees, built once per node and loaded by the other ranks (take_gpu_host_build_save / _load)
                hb = shared_host_build(flat)
                try:
                    self.gs = api.GpuScene(flat, device=device, prebuilt=hb)
                finally:
                    hb.close()
            else:
                # default: every rank builds its fast tree on ITS device (milliseconds) and the reference-order tree on a
                # background host thread pool sized to its share of the node's cores
                local_world = int(os.environ.get("LOCAL_WORLD_SIZE", self.world))
                if self.world > 1 and "TAKE_HOST_THREADS" not in os.environ:
                    os.environ["TAKE_HOST_THREADS"] = str(max(1, (os.cpu_count() or 1) // max(1, local_world)))
                self.gs = api.GpuScene(flat, device=device)
            self.ext = torch.cuda.ExternalStream(self.gs.stream, device=self.device)   # the library's render stream
            self.side = torch.cuda.Stream(device=self.device)                          # reduce + read-back
            self._stream = lambda st: torch.cuda.stream(st)
            self._event = lambda: torch.cuda.Event(enable_timing=True)
        else:
            self.device = torch.device("cpu")
            self.gs = None
            eng = _HostEngine()
            self.ext = self.side = None
            self._stream, self._event = eng.stream, eng.event
        self.d_sum = [torch.zeros((H, W, 3), dtype=torch.float64, device=self.device) for _ in range(2)]
        self.d_sq = [torch.zeros((H, W, 3), dtype=torch.float64, device=self.device) for _ in range(2)] if sumsq else [None, None]
        self.h_sum = self.h_sq = None
        self.jobs = [None, None]
        self.n = 0

    def _host_buffers(self):
        if self.h_sum is None:
            t = self.torch
            H, W = self.flat.height, self.flat.width
            pin = self.device.type == "cuda"
            self.h_sum = [t.empty((H, W, 3), dtype=t.float64, pin_memory=pin) for _ in range(2)]
            self.h_sq = [t.empty((H, W, 3), dtype=t.float64, pin_memory=pin) for _ in range(2)] if self.sumsq else [None, None]

    def submit(self, integrator, max_depth, spp_begin, spp_end, seed=0, to_host=False, flags=0, shard=True, local=False) -> ShardedJob:
        """shard=False: this rank renders the whole range itself; local=True: no reduce (a single-rank job inside a
        multi-rank process group -- e.g. rank 0 checking a reduced image against its own render).  Every rank of the group
        must make the same sequence of non-local submits (the reduce is a collective)."""
        k = self.n & 1
        self.n += 1
        if self.jobs[k] is not None and not self.jobs[k].collected:
            self.jobs[k].wait()                       # its buffers are about to be reused
        lo, hi = shard_spp(spp_begin, spp_end, self.rank, self.world) if shard else (spp_begin, spp_end)
        with self._stream(self.ext):                  # stream-ordered before the kernels that accumulate into them
            self.d_sum[k].zero_()
            if self.sumsq:
                self.d_sq[k].zero_()
        # (returns when this rank's kernels are done: the reduce below is enqueued behind complete data)
        if self.render_fn is None:
            st = self.gs.render_sums_device(self.d_sum[k].data_ptr(), self.d_sq[k].data_ptr() if self.sumsq else 0, integrator,
                                            max_depth, lo, hi, seed=seed, flags=flags)
        else:
            st = self.render_fn(self.d_sum[k], self.d_sq[k], integrator, max_depth, lo, hi, seed, flags)
        h_sum = h_sq = None
        with self._stream(self.side):
            if self.world > 1 and not local:
                self.dist.reduce(self.d_sum[k], dst=0)      # NCCL sum-reduce onto rank 0 (NVLink / NVSwitch)
                if self.sumsq:
                    self.dist.reduce(self.d_sq[k], dst=0)
            if to_host and self.rank == 0:
                self._host_buffers()
                h_sum = self.h_sum[k]
                h_sum.copy_(self.d_sum[k], non_blocking=True)
                if self.sumsq:
                    h_sq = self.h_sq[k]
                    h_sq.copy_(self.d_sq[k], non_blocking=True)
            done = self._event()
            done.record(self.side)
        job = ShardedJob(st, self.d_sum[k], self.d_sq[k], None if h_sum is None else h_sum.numpy(),
                         None if h_sq is None else h_sq.numpy(), done)
        self.jobs[k] = job
        return job

    def drain(self):
        for j in self.jobs:
            if j is not None and not j.collected:
                j.wait()

    def close(self):
        self.drain()
        if self.gs is not None:
            self.gs.close()
