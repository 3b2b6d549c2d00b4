"""Probe: how much does ray ordering matter for secondary rays?  (development aid)"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from take_b200 import api, scenes
from oracle import bindings as ob
flat = scenes.heightfield().flat()
gs = api.GpuScene(flat)
sc = ob.OracleLib().load(flat)
H, W = flat.height, flat.width
py, px = np.mgrid[0:H, 0:W]
rays = sc.primary_rays(px.ravel(), py.ravel(), seed=1)
prim, t, uv = gs.intersect(rays)
hit = prim >= 0
o = rays[hit, 0:3] + rays[hit, 3:6] * t[hit, None]
rng = np.random.default_rng(0)
d = rng.normal(size=o.shape); d /= np.linalg.norm(d, axis=1, keepdims=True); d[:, 1] = np.abs(d[:, 1])
sec = np.empty((len(o), 8)); sec[:, 0:3] = o; sec[:, 3:6] = d; sec[:, 6] = 1e-7; sec[:, 7] = np.inf
print("secondary rays:", len(sec))
def key(sec, grid):
    lo, hi = flat.positions.min(0), flat.positions.max(0)
    q = np.clip(((sec[:, 0:3] - lo) / (hi - lo) * grid).astype(np.int64), 0, grid - 1)
    octant = (sec[:, 3] > 0) * 4 + (sec[:, 4] > 0) * 2 + (sec[:, 5] > 0) * 1
    m = np.zeros(len(sec), np.int64)
    for b in range(int(np.log2(grid))):
        for a in range(3):
            m |= ((q[:, a] >> b) & 1) << (3 * b + a)
    return octant, m
def run(order, label):
    r = torch.from_numpy(np.ascontiguousarray(sec[order])).cuda()
    hits = torch.empty((len(order), 4), dtype=torch.float64, device="cuda")
    best = 1e9
    for _ in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        rc = gs.lib.take_gpu_intersect_device(gs.h, r.data_ptr(), len(order), hits.data_ptr(), 0)
        torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
        assert rc == 0
    print(f"{label:40s} {best*1e3:8.3f} ms  {len(order)/best/1e9:6.3f} Grays/s")
n = len(sec)
run(np.arange(n), "pixel order (coherent origins)")
run(rng.permutation(n), "shuffled")
for grid in (16, 64):
    octant, m = key(sec, grid)
    run(np.lexsort((m, octant)), f"sorted octant, morton{grid}")
    run(np.lexsort((octant, m)), f"sorted morton{grid}, octant")
# primary for reference
r = np.arange(len(rays))
sec = rays
run(r, "primary rays, pixel order")
# ---- does grouping rays of similar LENGTH help?  (stable sorts: pixel order survives inside a class) ----
sec = np.empty((len(o), 8)); sec[:, 0:3] = o; sec[:, 3:6] = d; sec[:, 6] = 1e-7; sec[:, 7] = np.inf
p2, t2, _ = gs.intersect(sec)
tt = np.where(p2 >= 0, t2, 1e30)
for bins in (4, 16):
    cls = np.minimum((sec[:, 4] * bins).astype(np.int64), bins - 1)       # elevation of the direction (a predictor known at shade time)
    run(np.argsort(cls, kind="stable"), f"elevation classes x{bins}, pixel order inside")
q = np.quantile(tt[tt < 1e30], np.linspace(0, 1, 9)[1:-1]) if (tt < 1e30).any() else []
cls = np.searchsorted(q, tt)                                              # the hit distance itself (an oracle predictor: upper bound on the gain)
run(np.argsort(cls, kind="stable"), "hit-distance octiles (oracle), pixel order inside")
