"""Short profiling workload: config-2 scene, one wave of 2 spp at 1920x1080 (one-sample MIS), then one wave of the
multi-sample MIS integrator (adds the shadow kernel).  No torch import, a handful of launches: meant for `ncu`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from take_b200 import api, scenes
flat = scenes.heightfield().flat()
gs = api.GpuScene(flat)
for integ in ("one_sample_mis", "mis"):
    s, s2, st = gs.render_sums(integ, 5, 0, 2, seed=1)
    print(integ, st["ms_total"], st["extend_rays"], st["shadow_rays"], st["kernel_launches"])
gs.close()
