"""World-size-2 gloo test of the multi-GPU host logic (take_b200/dist.py) on CPU.  The per-rank renderer is the CPU
oracle (tests may use it); what is under test is the sharding and the reduce, which are the same code under NCCL."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from take_b200 import dist as tdist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_spp_partitions_exactly():
    for n, world in [(256, 8), (7, 3), (5, 8), (0, 4), (1024, 1)]:
        ranges = [tdist.shard_spp(10, 10 + n, r, world) for r in range(world)]
        assert ranges[0][0] == 10 and ranges[-1][1] == 10 + n
        assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
        sizes = [b - a for a, b in ranges]
        assert max(sizes) - min(sizes) <= 1 and sum(sizes) == n


def _worker(rank, world, port, spp, out_path):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import bindings as ob
    from take_b200 import scenes
    flat = scenes.cornell_box(24, 24, spp, materials="mixed").flat()
    sc = ob.OracleLib().load(flat)

    def render_range(lo, hi):
        return sc.render("mis", 3, lo, hi, seed=77, threads=2)

    res = tdist.render_sharded(render_range, 0, spp, reduce="all" if rank % 2 == 0 or True else "root")
    mean, var, n = res
    if rank == 0:
        np.savez(out_path, mean=mean.numpy(), var=var.numpy(), n=n)
    dist.barrier()
    dist.destroy_process_group()


def test_nccl_overlap_setting_respects_the_user(monkeypatch):
    """configure_nccl_for_overlap() limits NCCL's channels for the overlapped per-job reduce -- unless the user set them."""
    monkeypatch.delenv("NCCL_MAX_NCHANNELS", raising=False)
    tdist.configure_nccl_for_overlap()
    assert os.environ["NCCL_MAX_NCHANNELS"] == "4"
    monkeypatch.setenv("NCCL_MAX_NCHANNELS", "12")
    tdist.configure_nccl_for_overlap(2)
    assert os.environ["NCCL_MAX_NCHANNELS"] == "12"


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_render_equals_single_process(tmp_path, oracle_lib, world):
    from take_b200 import scenes
    spp = 7   # deliberately not divisible by the world size
    out = str(tmp_path / "out.npz")
    port = 29500 + (os.getpid() % 1000) + world
    mp.spawn(_worker, args=(world, port, spp, out), nprocs=world, join=True)
    got = np.load(out)
    flat = scenes.cornell_box(24, 24, spp, materials="mixed").flat()
    sc = oracle_lib.load(flat)
    s, s2 = sc.render("mis", 3, 0, spp, seed=77)
    mean = s / spp
    # same samples, summed in a different order: equal to rounding
    assert np.abs(got["mean"] - mean).max() <= 1e-12 * np.abs(mean).max()
    var = np.maximum(s2 / spp - mean ** 2, 0) * spp / (spp - 1) / spp
    assert np.abs(got["var"] - var).max() <= 1e-9 * max(var.max(), 1e-30)
    assert int(got["n"]) == spp


def _build_worker(rank, world, port, out_dir, fail=False):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["LOCAL_RANK"] = str(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from take_b200 import scenes
    flat = scenes.multi_light(16, 16, 1, n_side=6).flat()
    if fail:                          # the building rank fails: every rank must raise, none may hang at a barrier
        if rank == 0:
            flat.prim_material = flat.prim_material.copy()
            flat.prim_material[0] = 12345     # out of range -> take_gpu_host_build refuses the scene
        try:
            tdist.shared_host_build(flat, directory=os.path.join(out_dir, "x"))
            raised = ""
        except RuntimeError as ex:
            raised = str(ex)
        open(os.path.join(out_dir, f"raised{rank}.txt"), "w").write(raised)
        dist.barrier()
        dist.destroy_process_group()
        return
    os.makedirs(os.path.join(out_dir, "x"), exist_ok=True)
    hb = tdist.shared_host_build(flat, directory=os.path.join(out_dir, "x"))
    a = hb.arrays()
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), **{k: a[k] for k in ("ref_nodes", "dfs_rank", "wide_nodes", "leaf_prims", "leaf_records")},
             ms_fast=a["ms_fast"])
    hb.close()
    dist.barrier()
    dist.destroy_process_group()


def test_shared_host_build_builds_once_per_node(tmp_path):
    """take_b200.dist.shared_host_build: local rank 0 builds and saves, the other ranks load -- every rank ends up with the
    same structures (the loaded ones carry rank 0's build times: they were not rebuilt) and the file is removed."""
    world = 2
    port = 29500 + (os.getpid() % 1000) + 17
    mp.spawn(_build_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    for k in ("ref_nodes", "dfs_rank", "wide_nodes", "leaf_prims", "leaf_records"):
        assert np.array_equal(r0[k], r1[k]), k
    assert float(r0["ms_fast"]) == float(r1["ms_fast"])
    assert os.listdir(tmp_path / "x") == []          # the private exchange directory is gone


def test_shared_host_build_failure_reaches_every_rank(tmp_path):
    world = 2
    port = 29500 + (os.getpid() % 1000) + 23
    os.makedirs(tmp_path / "x")
    mp.spawn(_build_worker, args=(world, port, str(tmp_path), True), nprocs=world, join=True)
    for r in range(world):
        assert "shared_host_build failed" in open(tmp_path / f"raised{r}.txt").read()
    assert os.listdir(tmp_path / "x") == []


def _sharded_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import bindings as ob
    from take_b200 import scenes
    flat = scenes.cornell_box(20, 16, 4, materials="mixed").flat()
    sc = ob.OracleLib().load(flat)
    calls = []

    def render_fn(d_sum, d_sq, integrator, max_depth, lo, hi, seed, flags):
        calls.append((lo, hi))
        s, s2, st = sc.render(integrator, max_depth, lo, hi, seed=seed, threads=2, stats=True)
        d_sum += torch.from_numpy(s)
        if d_sq is not None:
            d_sq += torch.from_numpy(s2)
        return {"extend_rays": int(st[0]), "shadow_rays": int(st[1]), "samples": (hi - lo) * flat.width * flat.height}

    sr = tdist.ShardedRenderer(flat, sumsq=True, render_fn=render_fn)
    # three pipelined jobs (two buffer sets: the third reuses the first), the last one not divisible by the world size
    ranges = [(0, 4), (4, 8), (8, 8 + world + 1)]
    jobs = [sr.submit("mis", 3, lo, hi, seed=5, to_host=True) for lo, hi in ranges[:2]]
    res = [tuple(np.array(a) if isinstance(a, np.ndarray) else a for a in jobs[0].wait())]   # views of a buffer set that job 2 reuses
    jobs.append(sr.submit("mis", 3, *ranges[2], seed=5, to_host=True))
    res += [jobs[1].wait(), jobs[2].wait()]
    # a local job inside the group: no collective, whole range on this rank
    st, loc, _ = sr.submit("mis", 3, 0, 2, seed=5, to_host=(rank == 0), shard=False, local=True).wait()
    if rank == 0:
        np.savez(os.path.join(out_dir, "sharded.npz"), s0=res[0][1], q0=res[0][2], s2=res[2][1], q2=res[2][2], loc=np.asarray(loc))
    assert calls[:3] == [tdist.shard_spp(lo, hi, rank, world) for lo, hi in ranges] and calls[3] == (0, 2)
    sr.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_renderer_jobs_equal_single_process(tmp_path, oracle_lib, world):
    """take_b200.dist.ShardedRenderer (the path bench.py --gpus N drives), gloo backend, CPU renderer injected: every job's
    reduced image on rank 0 equals the single-process render of the same sample range; buffer sets are reused safely."""
    from take_b200 import scenes
    port = 29500 + (os.getpid() % 1000) + 40 + world
    mp.spawn(_sharded_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "sharded.npz")
    sc = oracle_lib.load(scenes.cornell_box(20, 16, 4, materials="mixed").flat())
    for tag, (lo, hi) in (("0", (0, 4)), ("2", (8, 8 + world + 1))):
        s, s2 = sc.render("mis", 3, lo, hi, seed=5)
        assert np.abs(got["s" + tag] - s).max() <= 1e-12 * np.abs(s).max()
        assert np.abs(got["q" + tag] - s2).max() <= 1e-12 * np.abs(s2).max()
    s, _ = sc.render("mis", 3, 0, 2, seed=5)
    assert np.abs(got["loc"] - s).max() <= 1e-12 * np.abs(s).max()
