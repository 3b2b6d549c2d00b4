"""Throughput of every BASELINE.json config on the GPU next to the reference's CPU renderer (bounded samples).
Prints one JSON line per config; meant for DESIGN.md's table, not for the driver (that is bench.py)."""
import json, os, sys, tempfile, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from take_b200 import api, scenes
from oracle import bindings as ob

CONFIGS = [
    ("c1_cornell_512x512", lambda: scenes.cornell_box(), "mis", 64),
    ("c2_heightfield_1M_1920x1080", lambda: scenes.heightfield(), "one_sample_mis", 16),
    ("c3_ibl_textured_1024x1024", lambda: scenes.ibl_scene(), "one_sample_mis", 32),
    ("c4_multi_light_400_1920x1080", lambda: scenes.multi_light(), "mis", 16),
    ("c5_instanced_10M_3840x2160", lambda: scenes.instanced_spheres(), "mis", 4),
]
only = [a for a in sys.argv[1:] if not a.startswith("-")]
cores = os.cpu_count()
O = ob.OracleLib()
R = ob.RefLib() if ob.have_ref() else None
_warm = api.GpuScene(scenes.cornell_box(width=32, height=32).flat())   # CUDA context + module load happen here, not in the first timing
_warm.render_sums("mis", 2, 0, 1, seed=1, sumsq=False)
_warm.close()
for name, make, integ, spp in CONFIGS:
    if only and not any(o in name for o in only):
        continue
    b = make()
    flat = b.flat()
    t0 = time.perf_counter(); gs = api.GpuScene(flat); create_s = time.perf_counter() - t0
    gs.render_sums(integ, 5, 0, spp, seed=1, sumsq=False)
    best = None
    for rep in range(3):
        s, _, st = gs.render_sums(integ, 5, 1000 * rep, 1000 * rep + spp, seed=1, sumsq=False)
        if best is None or st["ms_total"] < best["ms_total"]:
            best = st
    rays = best["extend_rays"] + best["shadow_rays"]
    line = {"config": name, "prims": flat.num_prims, "integrator": integ, "spp_timed": spp, "gpu_ms": round(best["ms_total"], 2),
            "gpu_mrays_s": round(rays / best["ms_total"] / 1e3, 1), "gpu_msamples_s": round(best["samples"] / best["ms_total"] / 1e3, 1),
            "rays_per_sample": round(rays / best["samples"], 3), "scene_create_s": round(create_s, 2), **{k: round(v, 1) for k, v in gs.info().items()}}
    gs.close()
    if "--no-cpu" not in sys.argv:
        # CPU: the reference's own integrator where it exists (no environment map), else our CPU restatement
        use_ref = R is not None and flat.env is None and flat.num_prims < 3_000_000
        if use_ref:
            d = tempfile.mkdtemp()
            cpu, kind = R.load(b.write(d)), "reference"
        else:
            cpu, kind = O.load(flat), "port"
        port = cpu if kind == "port" else O.load(flat) if flat.num_prims < 3_000_000 else None
        step = max(1, int(round(flat.width * flat.height / 1000000)))           # ~1 M path samples per call
        rows = len(range(0, flat.height, step))
        cpu.render(integ, 5, 0, 1, seed=1, threads=cores, sumsq=False, row_begin=0, row_step=max(step, 8))   # warm-up
        dt, n, k = 0.0, 0, 0
        while dt < 5.0 and k < 64:                                                # bounded sample: about 5 s of CPU work
            t0 = time.perf_counter()
            cpu.render(integ, 5, k, k + 1, seed=1, threads=cores, sumsq=False, row_begin=k % step, row_step=step)
            dt += time.perf_counter() - t0
            n += len(range(k % step, flat.height, step)) * flat.width
            k += 1
        line.update(cpu_kind=kind, cpu_cores=cores, cpu_msamples_s=round(n / dt / 1e6, 4), cpu_sample=f"{n} samples in {dt:.1f} s",
                    cpu_mrays_s=round(n / dt / 1e6 * line["rays_per_sample"], 3),
                    speedup=round(line["gpu_msamples_s"] / (n / dt / 1e6), 0))
    print(json.dumps(line), flush=True)
