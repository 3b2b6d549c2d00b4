"""Diagnostic run on a GPU box: parity of the CUDA path against the CPU oracle, with verbose reporting.
(Development aid; the gating checks live in tests/ -m gpu.)"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from take_b200 import api, scenes
from oracle import bindings as ob

O = ob.OracleLib()
out = {}

def cmp_isect(name, gs, os_, rays, label):
    op, ot, ouv = os_.intersect(rays)
    res = {}
    for mode, exact in (("exact", True), ("fast", False)):
        t0 = time.time()
        gp, gt, guv = gs.intersect(rays, exact=exact)
        dt = time.time() - t0
        idbad = int((gp != op).sum())
        hit = op >= 0
        tbad = int((gt[hit & (gp == op)] != ot[hit & (gp == op)]).sum())
        uvbad = int((guv[hit & (gp == op)] != ouv[hit & (gp == op)]).any(axis=1).sum())
        res[mode] = dict(n=len(rays), id_mismatch=idbad, t_bit_mismatch=tbad, uv_bit_mismatch=uvbad, sec=round(dt, 3))
        if idbad:
            bad = np.nonzero(gp != op)[0][:5]
            res[mode]["examples"] = [dict(i=int(i), gpu=(int(gp[i]), float(gt[i])), cpu=(int(op[i]), float(ot[i]))) for i in bad]
    print(name, label, json.dumps(res))
    return res

def check_scene(name, builder, integrators=("mis", "raw", "one_sample_mis"), spp=2, nsamp=20000, render=True):
    flat = builder.flat()
    t0 = time.time(); os_ = O.load(flat); t_or = time.time() - t0
    t0 = time.time(); gs = api.GpuScene(flat); t_gpu = time.time() - t0
    info = gs.info()
    print(f"== {name}: prims={flat.num_prims} oracle_load={t_or:.2f}s gpu_create={t_gpu:.2f}s info={info}")
    H, W = flat.height, flat.width
    rng = np.random.default_rng(1)
    n = min(H * W, 200000)
    pix = rng.choice(H * W, n, replace=False)
    py, px = pix // W, pix % W
    rays = os_.primary_rays(px, py, seed=7)
    r = {"primary": cmp_isect(name, gs, os_, rays, "primary")}
    op, ot, _ = os_.intersect(rays)
    sec = ob.secondary_rays(rays, ot, op, seed=3)
    r["secondary"] = cmp_isect(name, gs, os_, sec, "secondary")
    # pixel-centre (degenerate ties possible) -- exact mode must still agree
    rays_c = os_.primary_rays(px, py, seed=7, jitter=False)
    r["centre"] = cmp_isect(name, gs, os_, rays_c, "pixel-centre")
    # occlusion on finite segments
    seg = sec.copy(); seg[:, 7] = rng.uniform(0.1, 3.0, len(seg)) * np.abs(flat.positions).max() * 0.2
    go = gs.occluded(seg); oo = os_.occluded(seg)
    r["occluded_mismatch"] = int((go != oo).sum())
    print(name, "occluded mismatches", r["occluded_mismatch"], "of", len(seg))
    # per-sample radiance
    m = min(nsamp, H * W)
    sp = rng.choice(H * W, m, replace=False)
    spy, spx = (sp // W).astype(np.int32), (sp % W).astype(np.int32)
    ss = rng.integers(0, 64, m).astype(np.int64)
    for integ in integrators:
        a = os_.radiance_samples(spx, spy, ss, integ, 5, seed=11)
        b = gs.radiance_samples(spx, spy, ss, integ, 5, seed=11)
        err = np.abs(a - b).max(axis=1) / (np.abs(a).max(axis=1) + 1e-300)
        exact = int((np.abs(a - b).max(axis=1) == 0).sum())
        big = int((err > 1e-9).sum())
        r[f"samples_{integ}"] = dict(n=m, bit_exact=exact, rel_gt_1e9=big, max_rel=float(err.max()), mean_cpu=float(a.mean()), mean_gpu=float(b.mean()))
        print(name, integ, "per-sample:", json.dumps(r[f"samples_{integ}"]))
        if big:
            for i in np.nonzero(err > 1e-9)[0][:5]:
                print("    ex", int(spx[i]), int(spy[i]), int(ss[i]), a[i], b[i])
    if render:
        for integ in integrators:
            cs, cs2, cst = os_.render(integ, 5, 0, spp, seed=5, stats=True)
            g, g2, st = gs.render_sums(integ, 5, 0, spp, seed=5)
            d = np.abs(cs - g)
            rel = d.max() / max(np.abs(cs).max(), 1e-300)
            nbad = int((d.max(axis=2) > 1e-9 * (np.abs(cs).max(axis=2) + 1e-12)).sum())
            r[f"render_{integ}"] = dict(max_abs=float(d.max()), pixels_off=nbad, sum_cpu=float(cs.sum()), sum_gpu=float(g.sum()),
                                        sq_max_abs=float(np.abs(cs2 - g2).max()), cpu_stats=[int(v) for v in cst],
                                        gpu_stats={k: st[k] for k in ("samples", "extend_rays", "shadow_rays", "shaded", "ms_total", "kernel_launches")})
            print(name, integ, "render:", json.dumps(r[f"render_{integ}"]))
    out[name] = r
    gs.close(); os_.close()

print("devices:", api.device_count(), api.load_library().take_gpu_version())
check_scene("cornell", scenes.cornell_box(128, 128, 4))
check_scene("cornell_mixed", scenes.cornell_box(128, 128, 4, materials="mixed"))
check_scene("multi_light", scenes.multi_light(160, 96, 4, n_side=6))
check_scene("heightfield_32k", scenes.heightfield(128, 160, 90, 4))
if "--big" in sys.argv:
    b = scenes.heightfield(708, 1920, 1080, 4)
    check_scene("heightfield_1m", b, integrators=("one_sample_mis",), spp=1, render=False)
    flat = b.flat()
    gs = api.GpuScene(flat)
    for integ in ("one_sample_mis", "mis"):
        for rep in range(3):
            os.environ["TAKE_STAGE_TIMES"] = "1" if rep == 2 else "0"
            s, s2, st = gs.render_sums(integ, 5, 0, 4, seed=1)
            rays = st["extend_rays"] + st["shadow_rays"]
            print(f"PERF heightfield_1m {integ} rep{rep}: {st['ms_total']:.1f} ms, {rays/st['ms_total']/1e3:.1f} Mrays/s, "
                  f"{st['samples']/st['ms_total']/1e3:.1f} Msamples/s", json.dumps(st))
    os.environ["TAKE_COUNT_TESTS"] = "1"
    s, s2, st = gs.render_sums("one_sample_mis", 5, 0, 1, seed=1)
    print("COUNTS", json.dumps(st))
    os.environ["TAKE_COUNT_TESTS"] = "0"
    gs.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/gpu_check.json", "w"), indent=1)
print("DONE")
