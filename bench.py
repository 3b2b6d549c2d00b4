#!/usr/bin/env python
"""bench.py -- throughput of the hot path (BVH ray-scene intersection + Monte-Carlo path integration) on B200.

    python bench.py --gpus N --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path on the host cores

Headline workload (BASELINE.json configs[1]): procedural 1 002 530-triangle displaced mesh, Blinn-Phong microfacet BRDF
(the reference's only microfacet model), one-sample MIS, 1920x1080, max_depth 5.  A "step" is one job of SPP_PER_STEP
samples of every pixel PER GPU: with N GPUs (one process each) step k covers the sample-index range
[k*N*S, (k+1)*N*S), sharded over the ranks (weak scaling, scene replicated), and EVERY step ends with its own NCCL
sum-reduce of the partial accumulation buffers onto rank 0 (take_b200.dist.ShardedRenderer: the reduce of step k runs on a
side stream and overlaps the kernels of step k+1).  `e2e` is the same loop with rank 0 copying every step's REDUCED image
into pinned host memory inside the timed region.

Beside the headline the line carries `scenes`: every BASELINE.json config rendered as ONE job whose sample range is
sharded over the N ranks (strong scaling: total work fixed; configs 4 and 5 are the ones BASELINE.json shards), timed
until rank 0 holds the reduced image on the host, with scene creation reported next to it; at N = 1 each scene also
carries the CPU renderer on a bounded sample and the roofline of its traversal / shade kernels.
Prints one JSON line (see DESIGN.md "Measurement" for every key).
"""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = "c2_heightfield_1M_tri_blinn_microfacet_one_sample_mis_1920x1080"
INTEGRATOR = "one_sample_mis"
MAX_DEPTH = 5
SEED = 20261018
# SURVEY.md 8(d) counting rules (algorithmic bytes / flops)
B_BOX, B_TRI, B_RAY_EXT, B_RAY_SH, B_VERTEX, B_SAMPLE = 32, 48, 48, 36, 352, 12
F_BOX, F_TRI = 27, 60


def scene_table():
    """BASELINE.json configs: (key, name, builder, integrator, spp of the config, spp of the bench job, note)."""
    from take_b200 import scenes, sceneio
    return [
        ("c1", "c1_cornell_512x512_mis_64spp", lambda: scenes.cornell_box(), "mis", 64, 64, None),
        ("c2", WORKLOAD, lambda: scenes.heightfield(), INTEGRATOR, 256, 256, None),
        ("c2_ggx", "c2_heightfield_1M_tri_GGX_one_sample_mis_1920x1080", lambda: scenes.heightfield(mtype=sceneio.MAT_GGX),
         INTEGRATOR, 256, 64, "GGX BRDF is an extension without a reference implementation: parity unpinned"),
        ("c3", "c3_ibl_textured_1024x1024_one_sample_mis", lambda: scenes.ibl_scene(), INTEGRATOR, 512, 512,
         "importance-sampled environment map is an extension without a reference implementation: parity unpinned"),
        ("c4", "c4_multi_light_400_1920x1080_mis", lambda: scenes.multi_light(), "mis", 1024, 1024, None),
        ("c5", "c5_instanced_10M_3840x2160_mis", lambda: scenes.instanced_spheres(), "mis", 4096, 512,
         "the job is ONE of the config's eight 512-spp ranges (4096 spp = 8 such jobs), sharded over the ranks"),
    ]


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), float(d.get("sm_max_mhz", 1965.0)), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1965.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region.  nvidia-smi needs a few hundred milliseconds to start and
    the timed region of the default run is ~140 ms, so the sampler is started EARLY (before the scene is built), every row is
    stamped on arrival, and only rows that arrived while the GPU was under the bench load (begin() .. stop()) are used; if none
    did, stop() keeps the same load running until two have (at most 2 s)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.t0 = [], None, None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def begin(self):
        self.t0 = time.perf_counter()

    def _under_load(self):
        return [r for t, r in list(self.rows) if self.t0 is not None and t >= self.t0 + 0.02]

    def stop(self, keep_busy=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        deadline = time.perf_counter() + 2.0
        while keep_busy is not None and len(self._under_load()) < 2 and time.perf_counter() < deadline:
            keep_busy()                       # same kernels as the timed region, untimed
        rows = self._under_load()
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); smax.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def base_config(flat, name=WORKLOAD, integrator=INTEGRATOR):
    """The keys both arms print under `config` (identical dicts, so that the driver can tell the configs are the same)."""
    return {"workload": name, "integrator": integrator, "max_depth": MAX_DEPTH, "resolution": [flat.width, flat.height],
            "triangles": int(flat.num_prims)}


# ------------------------------------------------------------------------------------------------------------
# CPU side: the reference's own code (oracle/_ref) when it was compiled, else our CPU restatement
# ------------------------------------------------------------------------------------------------------------
class CpuRenderer:
    def __init__(self, builder, flat, integrator=INTEGRATOR, allow_ref=True):
        from oracle import bindings as ob
        self.ob = ob
        self.integrator = integrator
        self.cores = os.cpu_count() or 1
        # the reference itself where it exists and the scene is one it can parse (no environment map, no GGX) and build in
        # reasonable time (its single-threaded parser + BVH build need minutes at 10 M triangles)
        can_ref = (allow_ref and ob.have_ref() and flat.env is None and flat.num_prims < 3_000_000 and
                   not (flat.materials["type"] == 12).any())
        self.port = None
        if can_ref:
            self.kind = "reference"
            self.tmp = tempfile.TemporaryDirectory()
            self.xml = builder.write(self.tmp.name)
            self.scene = ob.RefLib().load(self.xml)
            self.port = ob.OracleLib().load(flat)    # ray counts per sample (bit-identical paths)
        else:
            self.kind = "port"
            self.scene = self.port = ob.OracleLib().load(flat)
        self.W, self.H = flat.width, flat.height

    def run(self, sample_index, row_begin, row_step):
        """Render one sample of every row_step-th image row with all host threads; returns (seconds, samples)."""
        t0 = time.perf_counter()
        self.scene.render(self.integrator, MAX_DEPTH, sample_index, sample_index + 1, seed=SEED, threads=self.cores, sumsq=True,
                          row_begin=row_begin, row_step=row_step)
        dt = time.perf_counter() - t0
        rows = len(range(row_begin, self.H, row_step))
        return dt, rows * self.W

    def rays_per_sample(self, row_step=64):
        row_step = max(1, min(row_step, self.H // 8))
        _, _, st = self.port.render(self.integrator, MAX_DEPTH, 0, 1, seed=SEED, threads=self.cores, stats=True, row_begin=0,
                                    row_step=row_step)
        n = len(range(0, self.H, row_step)) * self.W
        return float(st[0] + st[1]) / n

    def pick_row_step(self, target_seconds):
        dt, n = self.run(0, 0, 64)
        rate = n / max(dt, 1e-6)
        want = rate * target_seconds
        return int(max(1, min(64, round(self.W * self.H / max(want, 1.0))))), rate


def cpu_baseline(builder, flat, integrator=INTEGRATOR, target_seconds=12.0):
    """Bounded sample: frames of 1 spp each (sample indices 1, 2, ...; every `step`-th row when a frame would take longer
    than the budget) until ~target_seconds of CPU work."""
    cpu = CpuRenderer(builder, flat, integrator)
    rps = cpu.rays_per_sample()
    dt0, n0 = cpu.run(0, 0, 16)                      # warm the caches / thread pool; also sizes the row subset
    est_frame = dt0 / max(n0, 1) * flat.width * flat.height
    step = int(max(1, min(16, round(est_frame / max(target_seconds / 3.0, 1e-3)))))
    t_total, n_total, k = 0.0, 0, 0
    while t_total < target_seconds and k < 64:
        k += 1
        dt, n = cpu.run(k, k % step, step)
        t_total += dt
        n_total += n
    what = "the full frame" if step == 1 else f"every {step}-th row of the frame"
    return {"value": n_total * rps / t_total / 1e6, "unit": "Mrays/s", "samples_per_s": n_total / t_total, "cores": cpu.cores,
            "kind": cpu.kind,
            "sample": f"{k} spp of {what} ({flat.width}x{flat.height}; {n_total} path samples, {t_total:.1f} s), {rps:.3f} rays/sample, "
                      f"integrator {integrator}, max_depth {MAX_DEPTH}"}, cpu


def stock_cli_baseline(cpu):
    """The stock reference executable (src/main.cpp -> render(), oracle/_ref/take_ref) on config 1 as specified: its own
    tile loop, its own random_device seeds, its own thread pool -- nothing of ours in the process.  Returns samples/s from
    the "Took X seconds" line render.cpp:83 prints."""
    exe = os.path.join(ROOT, "oracle", "_ref", "take_ref")
    if not os.path.exists(exe) or getattr(cpu, "xml", None) is None:
        return None
    with tempfile.TemporaryDirectory() as cwd:      # it writes ./image.exr
        out = subprocess.run([exe, cpu.xml, "-max_depth", str(MAX_DEPTH), "-t", str(cpu.cores)], cwd=cwd, capture_output=True,
                             text=True, timeout=600).stdout
    m = re.search(r"Finish building rendering\. Took ([0-9.eE+-]+) seconds", out)
    return float(m.group(1)) if m else None


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from take_b200 import scenes
    builder = scenes.heightfield()
    flat = builder.flat()
    cpu = CpuRenderer(builder, flat)
    rps = cpu.rays_per_sample()
    # bound every step to a few seconds so that warmup + steps end within minutes
    row_step, _ = cpu.pick_row_step(max(1.0, min(6.0, 150.0 / max(1, args.steps + args.warmup))))
    for i in range(args.warmup):
        cpu.run(i, i % row_step, row_step)
    t_total, n_total = 0.0, 0
    for i in range(args.steps):
        dt, n = cpu.run(args.warmup + i, i % row_step, row_step)
        t_total += dt
        n_total += n
    value = n_total * rps / t_total / 1e6
    sample = (f"each step: 1 spp of every {row_step}-th row of the 1920x1080 frame ({n_total // max(1, args.steps)} path samples), "
              f"{rps:.3f} rays/sample")
    emit({
        "impl": "reference", "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_total / max(1, args.steps), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "samples_per_s": n_total / t_total,
        "config": base_config(flat),
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": cpu.cores, "kind": cpu.kind, "sample": sample},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


# ------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------
def ncu_metrics():
    """Per-kernel ncu figures (DRAM bytes per launch, L2->SM GB/s, issue-active %, lanes per instruction) written by
    tools/ncu_summary.py from an `ncu --set full` capture.  They are only quoted when the capture was taken from the kernel
    sources this run is built from (hash of take_b200/csrc); otherwise null, with the file to regenerate named."""
    from take_b200 import api
    p = os.path.join(ROOT, "profiles", "ncu_metrics.json")
    if not os.path.exists(p):
        return None, "no capture (tools/gpu_round.sh + tools/ncu_summary.py write profiles/ncu_metrics.json)"
    d = json.load(open(p))
    if d.get("source_hash") != api.kernel_source_hash():
        return None, f"profiles/ncu_metrics.json was captured from other kernel sources ({d.get('source_hash')}): stale, not quoted"
    return d, f"profiles/ncu_metrics.json ({d.get('tag')}, sources {d.get('source_hash')})"


def run_ours(args):
    import torch
    import torch.distributed as dist
    from take_b200 import api
    from take_b200 import dist as tdist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"      # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
        tdist.configure_nccl_for_overlap()         # few NCCL channels: the per-step reduce overlaps kernels, it does not race them
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    peak, sm_max, peak_src = measured_peaks()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def allmax(x):
        if world == 1:
            return float(x)
        t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(xs):
        t = torch.tensor([float(x) for x in xs], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t)
        return [float(v) for v in t.tolist()]

    clocks = ClockSampler(local) if rank == 0 else None   # started early: see the class comment
    table = {k: (name, make, integ, spp_cfg, spp_job, note) for k, name, make, integ, spp_cfg, spp_job, note in scene_table()}
    builder = table["c2"][1]()
    flat = builder.flat()
    H, W = flat.height, flat.width
    S = args.spp_per_step
    barrier()
    t0 = time.perf_counter()
    sr = tdist.ShardedRenderer(flat, local, sumsq=True)   # one replica per rank; fast tree built on the device
    scene_create_ms = 1e3 * (time.perf_counter() - t0)
    gs = sr.gs
    create_phases = gs.create_timings()
    info = gs.info()                                      # joins the background build of the reference-order tree (diagnostics below)
    scene_ready_ms = 1e3 * (time.perf_counter() - t0)

    def step_range(step):
        return step * world * S, (step + 1) * world * S

    def run_steps(first, n, to_host, acc=None):
        """n pipelined steps; returns the last job (all earlier ones are collected)."""
        jobs = []
        for i in range(n):
            lo, hi = step_range(first + i)
            jobs.append(sr.submit(INTEGRATOR, MAX_DEPTH, lo, hi, seed=SEED, to_host=to_host))
            if len(jobs) >= 2:
                st, _, _ = jobs[-2].wait()
                if acc is not None:
                    for k in acc:
                        acc[k] += st[k]
        st, s, s2 = jobs[-1].wait()
        if acc is not None:
            for k in acc:
                acc[k] += st[k]
        return jobs[-1], s

    # ---- timed region: K steps, each with its own reduce, device-timed (max over ranks) ---------------------------
    run_steps(0, args.warmup, to_host=False)         # also warms the collective (communicator set-up, buffer registration)
    barrier()
    if clocks:
        clocks.begin()
    totals = {"extend_rays": 0, "shadow_rays": 0, "samples": 0, "kernel_launches": 0}
    e0 = torch.cuda.Event(enable_timing=True)
    e0.record(sr.ext)
    last, d_sum = run_steps(args.warmup, args.steps, to_host=False, acc=totals)
    ms = allmax(e0.elapsed_time(last.done))          # first kernel of step 0 -> end of the last step's reduce
    barrier()
    rays_all, samples_all, launches_all = allsum([totals["extend_rays"] + totals["shadow_rays"], totals["samples"], totals["kernel_launches"]])
    mean_check = float((d_sum / (S * world)).mean().item()) if rank == 0 else 0.0   # the last step's reduced image
    # (ranks other than 0 wait in the e2e barrier below while rank 0 tops up its clock samples, if it has to)
    clock_info = clocks.stop(keep_busy=lambda: (sr.submit(INTEGRATOR, MAX_DEPTH, 0, S, seed=SEED, shard=False, local=True).wait(),)) if clocks else None

    # ---- e2e: the same loop, every step's REDUCED image copied to pinned host memory on rank 0 inside the timed region ----
    e2e_steps = max(2, min(args.steps, 8))
    run_steps(0, 2, to_host=True)                     # pinned buffers allocated and warm
    barrier()
    e2e_tot = {"extend_rays": 0, "shadow_rays": 0}
    t0 = time.perf_counter()
    _, h_sum = run_steps(100, e2e_steps, to_host=True, acc=e2e_tot)
    e2e_s = allmax(time.perf_counter() - t0)
    (e2e_rays,) = allsum([e2e_tot["extend_rays"] + e2e_tot["shadow_rays"]])
    e2e = {"value": e2e_rays / e2e_s / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": 40,
           "d2h_bytes_per_step": int(2 * H * W * 3 * 8), "steps": e2e_steps, "ms_per_step": 1e3 * e2e_s / e2e_steps,
           "api": "take_b200.dist.ShardedRenderer.submit(to_host=True): take_gpu_render_device per rank, one NCCL sum-reduce per "
                  "buffer per step onto rank 0, ONE device->host copy of the reduced sum and sum-of-squares images on rank 0 "
                  "(pinned memory), pipelined with the next step; host wall clock, max over ranks",
           "note": "camera rays are generated on the device (replaces render.cpp:69-75), so the per-step host input is the "
                   "40-byte options struct; the scene is uploaded once by take_gpu_scene_create "
                   f"({scene_create_ms:.0f} ms; the fast tree is built on the device, the reference-order tree on a background "
                   f"host thread, done {scene_ready_ms:.0f} ms after the call started; renders do not wait for it)"}

    # ---- correctness inside the run (N > 1): the reduced image of a sharded job == the same range rendered on ONE GPU ----
    checks = {}
    if world > 1:
        lo, hi = 7, 7 + 3 * world + 1                 # not divisible by the world size on purpose
        _, red, red2 = sr.submit(INTEGRATOR, MAX_DEPTH, lo, hi, seed=SEED, to_host=True).wait()
        if rank == 0:
            red, red2 = red.copy(), red2.copy()
            _, one, one2 = sr.submit(INTEGRATOR, MAX_DEPTH, lo, hi, seed=SEED, to_host=True, shard=False, local=True).wait()
            rel = float(np.abs(red - one).max() / np.abs(one).max())
            rel2 = float(np.abs(red2 - one2).max() / np.abs(one2).max())
            checks["reduced_image_vs_single_gpu"] = {"spp_range": [lo, hi], "max_rel_diff_sum": rel, "max_rel_diff_sumsq": rel2,
                                                     "ok": bool(rel <= 1e-12 and rel2 <= 1e-12)}
        barrier()
        # the single-process product path (take_gpu_multi_create / _render / _destroy) on the same N devices, from rank 0
        if rank == 0:
            try:
                t0 = time.perf_counter()
                m = api.MultiGpuScene(flat, list(range(world)))
                t_create = time.perf_counter() - t0
                m.render_sums(INTEGRATOR, MAX_DEPTH, lo, hi, seed=SEED)          # warm
                t0 = time.perf_counter()
                ms_, ms2_, mst = m.render_sums(INTEGRATOR, MAX_DEPTH, lo, hi, seed=SEED)
                t_render = time.perf_counter() - t0
                m.close()
                rel = float(np.abs(ms_ - one).max() / np.abs(one).max())
                checks["multi_handle_vs_single_gpu"] = {"devices": world, "max_rel_diff_sum": rel, "ok": bool(rel <= 1e-12),
                                                        "create_ms": 1e3 * t_create, "render_ms": 1e3 * t_render,
                                                        "mrays_per_s": (mst["extend_rays"] + mst["shadow_rays"]) / t_render / 1e6}
            except Exception as ex:
                checks["multi_handle_vs_single_gpu"] = {"ok": False, "error": repr(ex)}
        barrier()

    # ---- the pure C-ABI host loop at N = 1, for continuity with round 1 (async render + wait, two pinned buffers) ----
    if world == 1:
        h_bufs = [(torch.empty((H, W, 3), dtype=torch.float64, pin_memory=True).numpy(),
                   torch.empty((H, W, 3), dtype=torch.float64, pin_memory=True).numpy()) for _ in range(2)]

        def host_steps(n):
            rays, tickets = 0, []
            for step in range(n + 1):
                if step < n:
                    tickets.append(gs.render_async(h_bufs[step & 1][0], h_bufs[step & 1][1], INTEGRATOR, MAX_DEPTH, step * S, (step + 1) * S, seed=SEED))
                if step >= 1:
                    st_ = gs.render_wait(tickets[step - 1])
                    rays += st_["extend_rays"] + st_["shadow_rays"]
            return rays

        host_steps(2)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        o_rays = host_steps(e2e_steps)
        torch.cuda.synchronize()
        e2e["c_abi_async_mrays_per_s"] = o_rays / (time.perf_counter() - t0) / 1e6
        t0 = time.perf_counter()
        _, _, st = gs.render_sums(INTEGRATOR, MAX_DEPTH, 0, S, seed=SEED, out=h_bufs[0])
        e2e["c_abi_blocking_mrays_per_s"] = (st["extend_rays"] + st["shadow_rays"]) / (time.perf_counter() - t0) / 1e6

    # ---- roofline of the dominant kernels (N = 1): events inside the library on the launching stream --------------
    roofline = None
    sm_mhz = (clock_info or {}).get("sm_mhz") or sm_max
    if rank == 0:
        roofline = kernel_rooflines(gs, sr, INTEGRATOR, S, peak, peak_src, sm_mhz, info["sm_count"], reps=max(2, min(args.steps, 4)))
    barrier()

    # ---- CPU baseline (reported, not the target) -----------------------------------------------------------------
    base = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            base, _ = cpu_baseline(builder, flat)
        except Exception as ex:  # the baseline must never take the GPU number down with it
            base = {"value": None, "unit": "Mrays/s", "cores": os.cpu_count(), "kind": "unavailable", "sample": repr(ex)}

    # ---- every BASELINE config as one sharded job (strong scaling), c2 first on the scene that is already resident ----
    scene_rows = []
    wanted = [k for k in args.scenes.split(",") if k]
    for key in wanted:
        if key not in table:
            continue
        try:
            scene_rows.append(scene_job(key, table[key], sr if key == "c2" else None, local, world, rank, barrier, allmax, allsum,
                                        peak, peak_src, sm_mhz, cpu=(world == 1 and not args.no_cpu_baseline)))
        except Exception as ex:
            scene_rows.append({"config": table[key][0], "error": repr(ex)})
            barrier()

    if rank == 0:
        line = {
            "metric": "Mrays/s", "value": rays_all / (ms * 1e-3) / 1e6, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "samples_per_s": samples_all / (ms * 1e-3),
            "config": base_config(flat),
            "run": {"spp_per_step_per_gpu": S, "parallelism": f"spp-range x{world}, scene replicated, one NCCL sum-reduce per step "
                                                              "(overlapped with the next step's kernels)",
                    "nccl_max_nchannels": os.environ.get("NCCL_MAX_NCHANNELS") if world > 1 else None,
                    "l2": "inputs larger than L2: each wave streams up to 33.5 M path records (300 B/slot, 10 GB) besides 130 MB of "
                          "tree + leaf records (L2 is 126 MB); no explicit flush",
                    "bvh": {"wide_nodes": int(info["fast_nodes"]), "depth": int(info["fast_tree_depth"]), "sah_cost": info["sah_cost"],
                            "build_ms_fast_tree_device": info["build_ms_fast_tree"],
                            "build_ms_reference_order_tree_host_background": info["build_ms_reference_tree"]},
                    "scene_create_phases_ms": {k: round(v, 2) for k, v in create_phases.items()}, "reference_tree_done_ms": scene_ready_ms},
            "e2e": e2e, "gpu_launches": int(launches_all), "clocks": clock_info, "roofline": roofline, "cpu_baseline": base,
            "rays_per_sample": rays_all / max(1.0, samples_all), "image_mean": mean_check, "scene_create_ms": scene_create_ms,
            "scenes": scene_rows, "checks": checks,
        }
        emit(line)
    sr.close()
    if world > 1:
        dist.destroy_process_group()


def kernel_rooflines(gs, sr, integrator, spp, peak, peak_src, sm_mhz, sm_count, reps=2):
    """Roofline entries of the traversal kernel (k_extend) and the shade kernel of `integrator` on the scene behind `gs`:
    algorithmic bytes per launch (SURVEY.md 8(d) counting rules x the counts of the instrumented kernels on the same inputs)
    over the kernel's launch duration, measured with CUDA events around every launch on the launching stream."""
    from take_b200 import api
    rs = {"ms_extend": 0.0, "ms_shade": 0.0, "ms_shadow": 0.0, "ms_generate": 0.0, "ms_sort": 0.0, "ms_other": 0.0, "ms_total": 0.0}
    for i in range(reps):
        s_ = sr.submit(integrator, MAX_DEPTH, i * spp, (i + 1) * spp, seed=SEED, shard=False, local=True, flags=api.RENDER_STAGE_TIMES).wait()[0]
        for k in rs:
            rs[k] += s_[k]
    cnt = sr.submit(integrator, MAX_DEPTH, 0, spp, seed=SEED, shard=False, local=True, flags=api.RENDER_COUNT_TESTS).wait()[0]
    passes = MAX_DEPTH + 2
    launches = max(1, cnt["waves"]) * passes
    ms_ext = rs["ms_extend"] / reps
    ms_shade = rs["ms_shade"] / reps
    ms_shadow = rs["ms_shadow"] / reps
    alg_bytes = B_BOX * cnt["box_tests"] + B_TRI * cnt["tri_tests"] + B_RAY_EXT * cnt["extend_rays"]
    alg_flops = F_BOX * cnt["box_tests"] + F_TRI * cnt["tri_tests"]
    achieved = alg_bytes / (ms_ext * 1e-3) / 1e9
    fp32_peak = 2 * 128 * sm_count * sm_mhz * 1e6 / 1e12
    ncu, ncu_src = ncu_metrics()
    kx = (ncu or {}).get("kernels", {}).get("k_extend") if integrator == INTEGRATOR else None
    r = {"bound": "hbm", "kernel": "k_extend", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
         "traffic": (kx or {}).get("dram_bytes_per_launch"), "peak_source": peak_src,
         "bytes_per_launch": alg_bytes / launches, "ms_per_launch": ms_ext / launches, "launches_per_step": launches,
         "extend_grays_per_s": cnt["extend_rays"] / (ms_ext * 1e-3) / 1e9,
         "box_tests_per_ray": cnt["box_tests"] / max(1, cnt["extend_rays"]),
         "tri_tests_per_ray": cnt["tri_tests"] / max(1, cnt["extend_rays"]),
         "test_rate_tflops": alg_flops / (ms_ext * 1e-3) / 1e12, "fp32_peak_tflops": fp32_peak,
         "test_rate_frac_fp32": alg_flops / (ms_ext * 1e-3) / 1e12 / fp32_peak,
         "stage_share": {k[3:]: rs[k] / max(rs["ms_total"], 1e-9) for k in rs if k != "ms_total"},
         "ncu": kx, "ncu_source": ncu_src,
         "note": "frac compares ALGORITHMIC bytes with the HBM copy peak; the tree is largely L1/L2-resident (see `traffic` and "
                 "`ncu`: measured DRAM bytes, L2->SM GB/s, issue-active %, lanes per instruction), so frac > 1 is possible and "
                 "does not mean the kernel is HBM-bound"}
    sh_bytes = B_VERTEX * cnt["shaded"]
    shade = {"bound": "hbm", "kernel": f"k_shade<{integrator}>", "achieved": sh_bytes / (ms_shade * 1e-3) / 1e9, "peak": peak,
             "unit": "GB/s", "frac": sh_bytes / (ms_shade * 1e-3) / 1e9 / peak, "bytes_per_launch": sh_bytes / launches,
             "ms_per_launch": ms_shade / launches, "vertices_per_s": cnt["shaded"] / (ms_shade * 1e-3),
             "rule": "352 algorithmic bytes per shaded vertex (SURVEY.md 8d)",
             "ncu": (ncu or {}).get("kernels", {}).get("k_shade") if integrator == INTEGRATOR else None}
    r["kernels"] = {"k_shade": shade}
    if ms_shadow > 0 and cnt["shadow_rays"] > 0:
        sb = B_BOX * cnt["shadow_box_tests"] + B_TRI * cnt["shadow_tri_tests"] + B_RAY_SH * cnt["shadow_rays"]
        r["kernels"]["k_shadow"] = {"bound": "hbm", "kernel": "k_shadow", "achieved": sb / (ms_shadow * 1e-3) / 1e9, "peak": peak,
                                    "unit": "GB/s", "frac": sb / (ms_shadow * 1e-3) / 1e9 / peak,
                                    "shadow_grays_per_s": cnt["shadow_rays"] / (ms_shadow * 1e-3) / 1e9,
                                    "box_tests_per_ray": cnt["shadow_box_tests"] / cnt["shadow_rays"],
                                    "tri_tests_per_ray": cnt["shadow_tri_tests"] / cnt["shadow_rays"]}
    return r


def scene_job(key, entry, sr, local, world, rank, barrier, allmax, allsum, peak, peak_src, sm_mhz, cpu):
    """One BASELINE config as ONE job: the sample range [0, spp_job) of every pixel sharded over the ranks, one NCCL reduce,
    the reduced image on rank 0's host.  Timed on the host from a barrier until every rank's part is complete (max over
    ranks); scene creation (host builds shared by the node + upload) is timed beside it."""
    from take_b200 import dist as tdist
    name, make, integ, spp_cfg, spp_job, note = entry
    own = sr is None
    builder = make()
    flat = builder.flat()
    create_ms = cold_ms = phases = prov = None
    npix = flat.width * flat.height
    if not own:
        # the headline scene is already resident (its creation was the first thing this process did with the library: the
        # top-level scene_create_ms carries module loading and the first growth of the memory pool); time a second replica
        from take_b200 import api
        barrier()
        t0 = time.perf_counter()
        g2 = api.GpuScene(flat, device=local)
        create_ms = 1e3 * allmax(time.perf_counter() - t0)
        phases = g2.create_timings()
        g2.close()
    if own:
        # cold: scene creation and the whole job back to back, as a host that renders one image does -- the fast tree is built
        # on the device, the reference-order tree on a background host thread that the render does not wait for
        barrier()
        t0 = time.perf_counter()
        sr = tdist.ShardedRenderer(flat, local, sumsq=False)
        create_ms = 1e3 * allmax(time.perf_counter() - t0)
        phases = sr.gs.create_timings()
        sr.submit(integ, MAX_DEPTH, 0, spp_job, seed=SEED, to_host=True).wait()
        cold_ms = 1e3 * allmax(time.perf_counter() - t0)
        prov = sr.gs.provisional_stats()
        sr.gs.info()        # joins the reference-order tree and uploads it, so that the warm job below does not contain that
    try:
        if not own:   # warm-up with the wave capacity of the real job (wave buffers and pinned memory are allocated on first use)
            per_rank = -(-spp_job // world)
            warm = min(per_rank, max(1, -(-(1 << 25) // npix)))
            sr.submit(integ, MAX_DEPTH, 0, warm * world, seed=SEED, to_host=True).wait()
        barrier()
        t0 = time.perf_counter()
        st, _, _ = sr.submit(integ, MAX_DEPTH, 0, spp_job, seed=SEED, to_host=True).wait()
        job_s = allmax(time.perf_counter() - t0)
        dev_ms = allmax(st["ms_total"])
        rays, samples, launches = allsum([st["extend_rays"] + st["shadow_rays"], st["samples"], st["kernel_launches"]])
        row = {"config": name, "key": key, "integrator": integ, "triangles": int(flat.num_prims), "resolution": [flat.width, flat.height],
               "spp_of_config": spp_cfg, "spp_job": spp_job, "scaling": "strong", "n_gpus": world,
               "job_ms": 1e3 * job_s, "job_device_ms_slowest_rank": dev_ms,
               "mrays_per_s": rays / job_s / 1e6, "samples_per_s": samples / job_s, "rays_per_sample": rays / max(samples, 1.0),
               "gpu_launches": int(launches), "d2h_bytes": int(flat.height * flat.width * 3 * 8 * (2 if sr.sumsq else 1)),
               "job": "ShardedRenderer.submit(0, spp_job, to_host=True): per-rank share + one NCCL reduce + one device->host copy on rank 0; "
                      "host wall clock from a barrier, max over ranks"}
        if create_ms is not None:
            row["scene_create_ms"] = create_ms               # take_gpu_scene_create: upload + fast tree built on the device
            row["scene_create_phases_ms"] = {k: round(v, 2) for k, v in phases.items()}
            row["e2e_job_ms"] = create_ms + 1e3 * job_s      # create + render + reduce + read-back (the Amdahl term)
            if cold_ms is not None:
                row["e2e_job_cold_ms"] = cold_ms             # the same measured cold in one go: + first-use allocations (wave buffers, pinned memory)
                row["provisional"] = prov                    # renders that ran ahead of the background reference-order tree / repeated
        if note:
            row["note"] = note
        if rank == 0 and world == 1:
            info = sr.gs.info()
            row["roofline"] = kernel_rooflines(sr.gs, sr, integ, min(spp_job, 16), peak, peak_src, sm_mhz, info["sm_count"], reps=2)
            row["bvh_build_ms"] = {"fast_tree_device": info["build_ms_fast_tree"], "reference_order_tree_host_background": info["build_ms_reference_tree"]}
            row["bvh"] = {"wide_nodes": int(info["fast_nodes"]), "depth": int(info["fast_tree_depth"]), "sah_cost": info["sah_cost"]}
        if cpu and rank == 0:
            try:
                row["cpu_baseline"], c = cpu_baseline(builder, flat, integ, target_seconds=5.0)
                if key == "c1":   # the stock executable on config 1 exactly as specified, as a second CPU figure
                    t = stock_cli_baseline(c)
                    if t:
                        row["cpu_stock_cli"] = {"samples_per_s": flat.width * flat.height * flat.spp / t, "seconds": t, "cores": c.cores,
                                                "what": "oracle/_ref/take_ref scene.xml -max_depth 5 (unmodified main.cpp + render.cpp), "
                                                        f"{flat.width}x{flat.height}, {flat.spp} spp, its own seeds"}
            except Exception as ex:
                row["cpu_baseline"] = {"value": None, "kind": "unavailable", "sample": repr(ex)}
        barrier()
        return row
    finally:
        if own:
            sr.close()


_REAL_STDOUT = None


def capture_stdout():
    """The contract is ONE JSON line on stdout.  Libraries underneath print there on their own (NCCL's version banner comes
    from C code, past sys.stdout), so file descriptor 1 is pointed at stderr for the whole run and the line is written to
    the saved descriptor at the end."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        os.write(1, data)
    else:
        os.write(_REAL_STDOUT, data)


def main():
    capture_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--spp-per-step", type=int, default=32)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--scenes", default="c1,c2,c2_ggx,c3,c4,c5", help="BASELINE configs rendered as one sharded job each ('' = none)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
