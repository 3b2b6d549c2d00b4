"""GPU vs oracle for the environment-map extension (see tests/test_env_extension.py for what pins the oracle here:
only the constant, non-sampled map is anchored in the reference; the rest is our own FP64 restatement)."""
import numpy as np
import pytest

from take_b200 import api, scenes, sceneio

from test_env_extension import furnace_scene

pytestmark = pytest.mark.gpu


def env_scenes():
    rng = np.random.default_rng(5)
    env = rng.uniform(0, 1, (16, 32, 3)) ** 3 * 5
    env[3, 7] = 300.0
    out = {}
    for sample in (False, True):
        out[f"furnace_{int(sample)}"] = furnace_scene(24, sample).flat()
        b = scenes.cornell_box(32, 32, 2, materials="mixed")
        b.meshes = [m for i, m in enumerate(b.meshes) if i != 1]      # open the ceiling to the sky
        b.environment(env, sample=sample)
        out[f"open_room_{int(sample)}"] = b.flat()
        out[f"ibl_{int(sample)}"] = scenes.ibl_scene(48, 40, 2, n_objects=9, env_size=(64, 32), sample_env=sample).flat()
    return out


@pytest.mark.parametrize("name", sorted(env_scenes()))
def test_env_scene_matches_oracle(gpu_lib, oracle_lib, name):
    flat = env_scenes()[name]
    gs, sc = api.GpuScene(flat), oracle_lib.load(flat)
    try:
        rng = np.random.default_rng(1)
        n = 2000
        px, py = rng.integers(0, flat.width, n), rng.integers(0, flat.height, n)
        s = rng.integers(0, 1 << 16, n)
        for integ in api.INTEGRATORS:
            if integ == "one_sample_mis_power" and flat.env_sample:
                continue      # the power pick has no entry for a sampled environment map (refused: test_ggx_extension)
            a = sc.radiance_samples(px, py, s, integ, 4, seed=3)
            b = gs.radiance_samples(px, py, s, integ, 4, seed=3)
            err = np.abs(a - b).max(axis=1) / (np.abs(a).max(axis=1) + 1e-30)
            assert (err > 1e-9).mean() <= 2e-3, (name, integ, float(err.max()))
            cs, cs2, cst = sc.render(integ, 4, 0, 3, seed=8, stats=True)
            g, g2, st = gs.render_sums(integ, 4, 0, 3, seed=8)
            bad = np.abs(g - cs).max(axis=2) > 1e-9 * (np.abs(cs).max(axis=2) + 1e-12)
            assert bad.mean() <= 3e-3, (name, integ)
            assert abs(g.sum() - cs.sum()) <= 1e-5 * abs(cs.sum()) + 1e-9
    finally:
        gs.close()
        sc.close()


def test_constant_environment_equals_background_on_gpu(gpu_lib):
    b = scenes.cornell_box(32, 32, 2, materials="mixed")
    bg = np.array([0.25, 0.5, 0.125])
    b.background = bg
    g0 = api.GpuScene(b.flat())
    b.environment(np.broadcast_to(bg, (4, 8, 3)).copy(), sample=False)
    g1 = api.GpuScene(b.flat())
    for integ in api.INTEGRATORS:
        a, a2, _ = g0.render_sums(integ, 5, 0, 3, seed=5)
        c, c2, _ = g1.render_sums(integ, 5, 0, 3, seed=5)
        assert np.array_equal(a, c) and np.array_equal(a2, c2)
    g0.close()
    g1.close()


def test_config3_full_size(gpu_lib, oracle_lib):
    """Config 3 at BASELINE size: 64 textured objects, 2048x1024 HDR environment, 1024x1024.  The oracle checks rows."""
    flat = scenes.ibl_scene().flat()
    assert (flat.width, flat.height) == (1024, 1024) and flat.env.shape == (1024, 2048, 3)
    gs, sc = api.GpuScene(flat), oracle_lib.load(flat)
    g, g2, st = gs.render_sums("one_sample_mis", 5, 0, 2, seed=6)
    cs, _ = sc.render("one_sample_mis", 5, 0, 2, seed=6, threads=16, row_begin=11, row_step=32)
    rows = np.arange(11, 1024, 32)
    err = np.abs(g[rows] - cs[rows]).max(axis=2) / (np.abs(cs[rows]).max(axis=2) + 1e-12)
    assert (err > 1e-9).mean() <= 2e-3
    assert np.isfinite(g).all()
    print("config 3:", flat.num_prims, "prims;", st["ms_total"], "ms for 2 spp;",
          (st["extend_rays"] + st["shadow_rays"]) / st["ms_total"] / 1e3, "Mrays/s")
    gs.close()
    sc.close()
