"""Short profiling workload for `ncu`: config-2 scene, one bench step (32 spp x 1920x1080: two waves of 1 M pixels x 32 samples, 33.5 M slots each) of the
one-sample-MIS integrator (7 extend + 7 shade launches), then one wave of the multi-sample MIS integrator (adds the
shadow kernel).  No torch import."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from take_b200 import api, scenes
scene = "heightfield"
for a in sys.argv[1:]:
    if a.startswith("--scene="):
        scene = a.split("=", 1)[1]
flat = scenes.build(scene).flat()
integrators = [a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--integrator=")] or ["one_sample_mis", "mis"]
spp = int(([a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--spp=")] or ["32"])[0])
gs = api.GpuScene(flat)
for integ in integrators:
    s, s2, st = gs.render_sums(integ, 5, 0, spp, seed=1, flags=api.RENDER_COUNT_TESTS if "--count" in sys.argv else 0)
    print(integ, {k: st[k] for k in ("ms_total", "extend_rays", "shadow_rays", "box_tests", "tri_tests", "kernel_launches", "waves")})
gs.close()
