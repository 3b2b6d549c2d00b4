#!/usr/bin/env python
"""bench.py -- throughput of the hot path (BVH ray-scene intersection + Monte-Carlo path integration) on B200.

    python bench.py --gpus N --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path on the host cores

Workload (BASELINE.json configs[1]): procedural 1 002 530-triangle displaced mesh, Blinn-Phong microfacet BRDF
(the reference's only microfacet model), one-sample MIS, 1920x1080, max_depth 5.  A "step" renders SPP_PER_STEP
samples of every pixel (one pass of generate -> extend -> sort -> shade -> accumulate waves); with N GPUs every rank
renders its own sample-index range of every step (weak scaling, scene replicated), and the partial accumulation
buffers are combined by ONE NCCL all-reduce inside the timed region.
Prints one JSON line (see DESIGN.md "Measurement" for every key).
"""
import argparse
import json
import os
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


def build_scene():
    from take_b200 import scenes
    return scenes.heightfield()          # n=708 -> 1 002 528 + 2 triangles, 1920x1080


# ------------------------------------------------------------------------------------------------------------
# CPU side: the reference's own code (oracle/_ref) when it was compiled, else our CPU restatement
# ------------------------------------------------------------------------------------------------------------
class CpuRenderer:
    def __init__(self, builder, flat):
        from oracle import bindings as ob
        self.ob = ob
        self.cores = os.cpu_count() or 1
        self.port = ob.OracleLib().load(flat)        # ray counts per sample (bit-identical paths)
        if ob.have_ref():
            self.kind = "reference"
            self.tmp = tempfile.TemporaryDirectory()
            self.scene = ob.RefLib().load(builder.write(self.tmp.name))
        else:
            self.kind = "port"
            self.scene = self.port
        self.W, self.H = flat.width, flat.height

    def run(self, sample_index, row_begin, row_step):
        """Render one sample of every row_step-th image row with all host threads; returns (seconds, samples)."""
        t0 = time.perf_counter()
        self.scene.render(INTEGRATOR, MAX_DEPTH, sample_index, sample_index + 1, seed=SEED, threads=self.cores, sumsq=True,
                          row_begin=row_begin, row_step=row_step)
        dt = time.perf_counter() - t0
        rows = len(range(row_begin, self.H, row_step))
        return dt, rows * self.W

    def rays_per_sample(self, row_step=64):
        _, _, st = self.port.render(INTEGRATOR, MAX_DEPTH, 0, 1, seed=SEED, threads=self.cores, stats=True, row_begin=0,
                                    row_step=row_step)
        n = len(range(0, self.H, row_step)) * self.W
        return float(st[0] + st[1]) / n

    def pick_row_step(self, target_seconds):
        dt, n = self.run(0, 0, 64)
        rate = n / max(dt, 1e-6)
        want = rate * target_seconds
        return int(max(1, min(64, round(self.W * self.H / max(want, 1.0))))), rate


def cpu_baseline(builder, flat, target_seconds=12.0):
    """Bounded sample: whole 1920x1080 frames of 1 spp each (sample indices 1, 2, ...) until ~target_seconds of CPU work."""
    cpu = CpuRenderer(builder, flat)
    rps = cpu.rays_per_sample()
    cpu.run(0, 0, 16)                      # warm the caches / thread pool
    t_total, n_total, k = 0.0, 0, 0
    while t_total < target_seconds and k < 64:
        k += 1
        dt, n = cpu.run(k, 0, 1)
        t_total += dt
        n_total += n
    return {"value": n_total * rps / t_total / 1e6, "unit": "Mrays/s", "samples_per_s": n_total / t_total, "cores": cpu.cores,
            "kind": cpu.kind,
            "sample": f"{k} spp of the full 1920x1080 frame ({n_total} path samples, {t_total:.1f} s), {rps:.3f} rays/sample, "
                      f"integrator {INTEGRATOR}, max_depth {MAX_DEPTH}"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    builder = build_scene()
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
        "config": {"workload": WORKLOAD, "integrator": INTEGRATOR, "max_depth": MAX_DEPTH, "resolution": [flat.width, flat.height],
                   "triangles": flat.num_prims},
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": cpu.cores, "kind": cpu.kind, "sample": sample},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


# ------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from take_b200 import api

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"      # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")

    clocks = ClockSampler(local) if rank == 0 else None   # started early: see the class comment
    builder = build_scene()
    flat = builder.flat()
    t0 = time.perf_counter()
    gs = api.GpuScene(flat, device=local)
    scene_create_ms = 1e3 * (time.perf_counter() - t0)
    info = gs.info()
    H, W = flat.height, flat.width
    S = args.spp_per_step
    ext = torch.cuda.ExternalStream(gs.stream, device=dev)
    d_sum = torch.zeros((H, W, 3), dtype=torch.float64, device=dev)
    d_sq = torch.zeros_like(d_sum)
    torch.cuda.synchronize()

    def spp_range(step):
        lo = (step * world + rank) * S
        return lo, lo + S

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    totals = {"rays": 0, "samples": 0, "launches": 0, "extend": 0, "shadow": 0}

    def device_step(step, flags=0, acc=None):
        lo, hi = spp_range(step)
        st = gs.render_sums_device(d_sum.data_ptr(), d_sq.data_ptr(), INTEGRATOR, MAX_DEPTH, lo, hi, seed=SEED, flags=flags)
        if acc is not None:
            acc["rays"] += st["extend_rays"] + st["shadow_rays"]
            acc["samples"] += st["samples"]
            acc["launches"] += st["kernel_launches"]
            acc["extend"] += st["extend_rays"]
            acc["shadow"] += st["shadow_rays"]
        return st

    # ---- timed region: K steps + the one reduction, device-timed on the launching stream -----------------------
    for i in range(args.warmup):
        device_step(i)
    if world > 1:                    # warm the collective too (communicator set-up, buffer registration)
        dist.all_reduce(d_sum)
        dist.all_reduce(d_sq)
    d_sum.zero_(); d_sq.zero_()
    barrier()
    if clocks:
        clocks.begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ext)
    for i in range(args.steps):
        device_step(args.warmup + i, acc=totals)
    if world > 1:
        dist.all_reduce(d_sum)       # the path's one exchange step: sum of the partial accumulation buffers (NCCL)
        dist.all_reduce(d_sq)
        torch.cuda.synchronize()
    e1.record(ext)
    e1.synchronize()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    agg = torch.tensor([totals["rays"], totals["samples"], totals["launches"]], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(agg)
    ms = float(ms.item())
    rays_all, samples_all, launches_all = (float(v) for v in agg.tolist())
    mean_check = float((d_sum / (S * args.steps * world)).mean().item())
    # (ranks other than 0 wait in the e2e barrier below while rank 0 tops up its clock samples, if it has to)
    clock_info = clocks.stop(keep_busy=lambda: (device_step(0), torch.cuda.synchronize())) if clocks else None

    # ---- e2e: the C-ABI calls a host application makes (host buffers, D2H of every step's result inside the timed region) ----
    # take_gpu_render_async + take_gpu_render_wait with two pinned result buffers: the read-back of step k overlaps the
    # kernels of step k+1; every step's sums are complete on the host when its wait returns, inside the timed region.
    h_bufs = [(torch.empty((H, W, 3), dtype=torch.float64, pin_memory=True).numpy(),
               torch.empty((H, W, 3), dtype=torch.float64, pin_memory=True).numpy()) for _ in range(2)]
    o_rays, e2e_steps = 0, max(2, min(args.steps, 6))
    import ctypes as C

    def host_steps(n):
        rays, tickets = 0, []
        for step in range(n + 1):
            if step < n:
                lo, hi = spp_range(step)
                tickets.append(gs.render_async(h_bufs[step & 1][0], h_bufs[step & 1][1], INTEGRATOR, MAX_DEPTH, lo, hi, seed=SEED))
            if step >= 1:
                st_ = gs.render_wait(tickets[step - 1])
                rays += st_["extend_rays"] + st_["shadow_rays"]
        return rays

    host_steps(2)                      # both result slots allocated and warm
    barrier()                          # every rank runs its own host loop, all at the same time
    t0 = time.perf_counter()
    o_rays = host_steps(e2e_steps)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:                      # whole job: rays of all ranks / the slowest rank's time
        t_max = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        r_sum = torch.tensor([float(o_rays)], dtype=torch.float64, device=dev)
        dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
        dist.all_reduce(r_sum)
        e2e_s, o_rays = float(t_max.item()), float(r_sum.item())
    if rank != 0:
        gs.close()
        dist.destroy_process_group()
        return
    # the blocking call, for comparison (copy not overlapped)
    st = api.TakeStats()
    o = api.TakeRenderOpts(api.INTEGRATORS[INTEGRATOR], MAX_DEPTH, 0, S, SEED, 0, 0)
    gs.lib.take_gpu_render(gs.h, C.byref(o), h_bufs[0][0].ctypes.data, h_bufs[0][1].ctypes.data, C.byref(st))
    t1 = time.perf_counter()
    rc = gs.lib.take_gpu_render(gs.h, C.byref(o), h_bufs[0][0].ctypes.data, h_bufs[0][1].ctypes.data, C.byref(st))
    assert rc == 0, gs.lib.take_gpu_last_error()
    blocking_s = time.perf_counter() - t1
    e2e = {"value": o_rays / e2e_s / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": C.sizeof(api.TakeRenderOpts),
           "d2h_bytes_per_step": int(h_bufs[0][0].nbytes + h_bufs[0][1].nbytes), "steps": e2e_steps, "ms_per_step": 1e3 * e2e_s / e2e_steps,
           "api": "take_gpu_render_async + take_gpu_render_wait, two pinned host buffers per rank; whole job = rays of all ranks / slowest rank",
           "blocking_call_mrays_per_s": (st.extend_rays + st.shadow_rays) / blocking_s / 1e6,
           "note": "camera rays are generated on the device (replaces render.cpp:69-75), so the per-step host input is the "
                   "options struct; the scene is uploaded once by take_gpu_scene_create "
                   f"({scene_create_ms:.0f} ms incl. host BVH builds)"}

    # ---- roofline of the dominant kernel (k_extend): events inside the library on the launching stream -----------
    peak, sm_max, peak_src = measured_peaks()
    rs = {"ms_extend": 0.0, "ms_shade": 0.0, "ms_generate": 0.0, "ms_sort": 0.0, "ms_other": 0.0, "ms_total": 0.0, "launches": 0}
    r_steps = max(2, min(args.steps, 4))
    for i in range(r_steps):
        s_ = device_step(i, flags=api.RENDER_STAGE_TIMES)
        for k in ("ms_extend", "ms_shade", "ms_generate", "ms_sort", "ms_other", "ms_total"):
            rs[k] += s_[k]
    cnt = device_step(0, flags=api.RENDER_COUNT_TESTS)
    passes = MAX_DEPTH + 2
    ext_launches_per_step = max(1, cnt["waves"]) * passes
    alg_bytes = B_BOX * cnt["box_tests"] + B_TRI * cnt["tri_tests"] + B_RAY_EXT * cnt["extend_rays"]
    alg_flops = F_BOX * cnt["box_tests"] + F_TRI * cnt["tri_tests"]
    ms_ext_step = rs["ms_extend"] / r_steps
    achieved = alg_bytes / (ms_ext_step * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "extend_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get("dram_bytes_per_launch")
    sm_mhz = (clock_info or {}).get("sm_mhz") or sm_max
    fp32_peak = 2 * 128 * info["sm_count"] * sm_mhz * 1e6 / 1e12
    roofline = {"bound": "hbm", "kernel": "k_extend", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src,
                "bytes_per_launch": alg_bytes / ext_launches_per_step, "ms_per_launch": ms_ext_step / ext_launches_per_step,
                "launches_per_step": ext_launches_per_step,
                "extend_grays_per_s": cnt["extend_rays"] / (ms_ext_step * 1e-3) / 1e9,
                "box_tests_per_ray": cnt["box_tests"] / max(1, cnt["extend_rays"]),
                "tri_tests_per_ray": cnt["tri_tests"] / max(1, cnt["extend_rays"]),
                "test_rate_tflops": alg_flops / (ms_ext_step * 1e-3) / 1e12, "fp32_peak_tflops": fp32_peak,
                "test_rate_frac_fp32": alg_flops / (ms_ext_step * 1e-3) / 1e12 / fp32_peak,
                "stage_share": {k[3:]: rs[k] / max(rs["ms_total"], 1e-9) for k in rs if k.startswith("ms_") and k != "ms_total"}}

    # ---- CPU baseline (reported, not the target) -----------------------------------------------------------------
    base = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            base = cpu_baseline(builder, flat)
        except Exception as ex:  # the baseline must never take the GPU number down with it
            base = {"value": None, "unit": "Mrays/s", "cores": os.cpu_count(), "kind": "unavailable", "sample": repr(ex)}

    line = {
        "metric": "Mrays/s", "value": rays_all / (ms * 1e-3) / 1e6, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "samples_per_s": samples_all / (ms * 1e-3),
        "config": {"workload": WORKLOAD, "integrator": INTEGRATOR, "max_depth": MAX_DEPTH, "resolution": [W, H],
                   "triangles": flat.num_prims, "spp_per_step_per_gpu": S, "parallelism": f"spp-range x{world}, scene replicated",
                   "l2": "inputs larger than L2: each wave streams up to 33.5 M path records (300 B/slot, 10 GB) besides 130 MB of tree + leaf records (L2 is 126 MB); no explicit flush",
                   "bvh": {"nodes": int(info["fast_nodes"]), "depth": int(info["fast_tree_depth"]),
                           "build_ms": info["build_ms_fast_tree"] + info["build_ms_reference_tree"]}},
        "e2e": e2e, "gpu_launches": int(launches_all), "clocks": clock_info, "roofline": roofline, "cpu_baseline": base,
        "rays_per_sample": rays_all / max(1.0, samples_all), "image_mean": mean_check, "scene_create_ms": scene_create_ms,
    }
    emit(line)
    gs.close()
    if world > 1:
        dist.destroy_process_group()


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
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
