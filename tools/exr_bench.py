"""Output step timing: the reference's imwrite (.exr, one host thread, vendored tinyexr + miniz) next to
take_gpu_exr_pack_device + take_gpu_exr_write_packed, on the same per-pixel sums.  Prints one JSON line per resolution."""
import json, os, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from take_b200 import api, scenes
from oracle import bindings as ob
from exr_reader import read_exr

R = ob.RefLib() if ob.have_ref() else None
for (W, H, spp) in ((1920, 1080, 256), (3840, 2160, 4096)):
    rng = np.random.default_rng(1)
    # a plausible image: smooth radiance field + path-tracing noise
    yy, xx = np.mgrid[0:H, 0:W]
    base = 0.4 + 0.3 * np.sin(xx / 97.0)[..., None] * np.cos(yy / 61.0)[..., None] + np.array([0.2, 0.1, 0.05])
    s = (base * (1 + 0.05 * rng.standard_normal((H, W, 3)))).clip(0) * spp
    gs = api.GpuScene(scenes.cornell_box(W, H, 1).flat())
    d = torch.from_numpy(s).cuda()
    torch.cuda.synchronize()
    td = tempfile.mkdtemp()
    ours = os.path.join(td, "ours.exr")
    best = {"pack_ms": 1e9, "write_ms": 1e9}
    for rep in range(3):
        t0 = time.perf_counter()
        packed = gs.exr_pack_device(d.data_ptr(), spp)
        t1 = time.perf_counter()
        api.write_exr_packed(ours, W, H, packed)
        t2 = time.perf_counter()
        best["pack_ms"] = min(best["pack_ms"], 1e3 * (t1 - t0)); best["write_ms"] = min(best["write_ms"], 1e3 * (t2 - t1))
    line = {"resolution": [W, H], "ours_pack_incl_d2h_ms": round(best["pack_ms"], 2), "ours_deflate_write_ms": round(best["write_ms"], 2),
            "ours_total_ms": round(best["pack_ms"] + best["write_ms"], 2), "host_threads": os.cpu_count(),
            "ours_file_bytes": os.path.getsize(ours), "d2h_bytes": int(packed.nbytes), "fp64_sums_bytes": int(s.nbytes)}
    if R is not None:
        ref = os.path.join(td, "ref.exr")
        mean = s * (1.0 / spp)
        t0 = time.perf_counter(); R.imwrite(ref, mean); t1 = time.perf_counter()
        line["reference_imwrite_ms"] = round(1e3 * (t1 - t0), 2)
        line["reference_file_bytes"] = os.path.getsize(ref)
        a, b = read_exr(ref), read_exr(ours)
        line["identical_pixels"] = bool(all(np.array_equal(a["planes"][c], b["planes"][c]) for c in "BGR"))
        line["speedup"] = round(line["reference_imwrite_ms"] / line["ours_total_ms"], 1)
    print(json.dumps(line))
    gs.close()
