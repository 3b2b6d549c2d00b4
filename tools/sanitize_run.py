"""Small workload touching every kernel, for `compute-sanitizer --tool memcheck` (one tool per gpurun call)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["TAKE_WAVE_SLOTS"] = "3000"          # several waves and pixel chunks even on tiny images
from take_b200 import api, scenes
todo = [scenes.cornell_box(40, 24, 2, materials="mixed").flat(), scenes.sphere_room(24, 24, 2).flat(),
        scenes.textured_room(24, 24, 2).flat(), scenes.heightfield(40, 48, 27, 2).flat(),
        scenes.ibl_scene(32, 24, 2, n_objects=4, env_size=(32, 16)).flat(), scenes.SceneBuilder(8, 8, (0, 0, 5), (0, 0, 0)).flat()]
rng = np.random.default_rng(0)
MODES = [{}, {"TAKE_REFILL": "0"}, {"TAKE_REFILL": "1", "TAKE_ORDERED_SORT": "0"}, {"TAKE_PACKET": "0", "TAKE_REFILL": "2"},
         {"TAKE_PROVISIONAL": "0"}, {"TAKE_DEVICE_BUILD": "0"}]
for flat in todo:
    for mode in MODES:
        for k, v in mode.items():
            os.environ[k] = v
        gs = api.GpuScene(flat)
        n = 777
        o = rng.uniform(-2, 2, (n, 3)) + np.array([0, 1, 0])
        d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
        rays = api.make_rays(o, d)
        gs.intersect(rays); gs.intersect(rays, exact=True); gs.occluded(rays)
        for integ in api.INTEGRATORS:
            if integ == "one_sample_mis_power" and (flat.env is not None or len(flat.lights) == 0):
                continue                      # refused by design (no power table for a sampled environment / no emitter)
            try:
                gs.render_sums(integ, 3, 0, 3, seed=1)
                gs.render_sums(integ, 3, 0, 2, seed=1, flags=api.RENDER_NO_SORT | api.RENDER_COUNT_TESTS | api.RENDER_STAGE_TIMES)
                if flat.env is None:
                    gs.render_sums(integ, 6, 0, 2, seed=1, flags=api.RENDER_RUSSIAN_ROULETTE)
                gs.radiance_samples(rng.integers(0, flat.width, 50), rng.integers(0, flat.height, 50), rng.integers(0, 9, 50), integ, 3, 2)
            except api.TakeGpuError as e:
                if "zero" in str(e) or "power" in str(e):
                    continue
                raise
        gs.close()
        for k in mode:
            del os.environ[k]
print("sanitize workload done")
