"""GGX BSDF (TAKE_MAT_GGX): an EXTENSION with no reference implementation (the reference's only microfacet model is
BlinnPhongMicrofacet).  Parity unpinned; what is checked is the internal consistency of our own FP64 restatement --
sampling, pdf and evaluation must describe the same distribution -- and, on the GPU, agreement with that restatement."""
import numpy as np
import pytest

from take_b200 import api, scenes, sceneio


def ggx_ball(res=24, alpha=0.3, ks=(1.0, 1.0, 1.0), sample_env=True):
    b = scenes.SceneBuilder(res, res, (0, 0, 4), (0, 0, 0), (0, 1, 0), 40.0, 4, (0, 0, 0))
    m = b.material(sceneio.MAT_GGX, ks, alpha=alpha)
    b.sphere((0, 0, 0), 1.0, m)
    b.environment(np.ones((8, 16, 3)), sample=sample_env)
    return b


@pytest.mark.parametrize("alpha", [0.08, 0.3, 0.8])
def test_bsdf_sampling_matches_pdf_and_eval(oracle_lib, alpha):
    """In a constant unit environment the radiance leaving a convex body is its directional albedo.  The `raw` integrator
    estimates it with BSDF sampling only (eval / pdf of sampled directions), `mis` additionally with environment
    sampling weighted by bsdf_pdf: the two agree only if sample, pdf and eval are mutually consistent.  With F = 1 the
    albedo can only lose energy to the masking term, never exceed 1."""
    n = 1500
    res = {}
    for integ, senv in (("raw", False), ("mis", True)):
        sc = oracle_lib.load(ggx_ball(alpha=alpha, sample_env=senv).flat())
        s, s2 = sc.render(integ, 0, 0, n, seed=7)      # max_depth 0: single scattering, no inter-reflection anyway
        mu = s / n
        centre = mu[9:15, 9:15, 0]
        var = np.maximum(s2 / n - mu ** 2, 0)[9:15, 9:15, 0] / n
        res[integ] = (centre.mean(), np.sqrt(var.sum()) / centre.size)
    (a, ea), (b, eb) = res["raw"], res["mis"]
    assert abs(a - b) <= 4 * np.hypot(ea, eb) + 5e-3, (alpha, res)
    # directional albedo of single-scatter GGX (F = 1) near normal incidence, by numerical quadrature of D G / (4 cos)
    expected = {0.08: 0.99, 0.3: 0.87, 0.8: 0.435}[alpha]
    assert abs(a - expected) < 0.03 and a <= 1.0 + 3 * ea, (alpha, res)


def test_heightfield_ggx_variant_builds(oracle_lib):
    b = scenes.heightfield(24, 40, 24, 2, mtype=sceneio.MAT_GGX)
    flat = b.flat()
    assert flat.materials[0]["type"] == sceneio.MAT_GGX and abs(flat.materials[0]["p"][0] - np.float32(0.14)) < 1e-9
    with pytest.raises(ValueError):
        b.write("/tmp/should_not_exist_ggx")
    sc = oracle_lib.load(flat)
    for integ in ("mis", "raw", "one_sample_mis"):
        s, _ = sc.render(integ, 5, 0, 2, seed=1)
        assert np.isfinite(s).all() and s.sum() > 0


@pytest.mark.gpu
def test_ggx_gpu_matches_restatement(gpu_lib, oracle_lib):
    for flat in (scenes.heightfield(32, 48, 27, 2, mtype=sceneio.MAT_GGX).flat(), ggx_ball(alpha=0.2).flat()):
        gs, sc = api.GpuScene(flat), oracle_lib.load(flat)
        rng = np.random.default_rng(3)
        n = 3000
        px, py = rng.integers(0, flat.width, n), rng.integers(0, flat.height, n)
        s = rng.integers(0, 1 << 16, n)
        for integ in api.INTEGRATORS:
            if integ == "one_sample_mis_power" and flat.env is not None and flat.env_sample:
                with pytest.raises(api.TakeGpuError):      # the power pick has no entry for a sampled environment map
                    gs.render_sums(integ, 4, 0, 1, seed=2)
                continue
            a = sc.radiance_samples(px, py, s, integ, 4, seed=5)
            b = gs.radiance_samples(px, py, s, integ, 4, seed=5)
            err = np.abs(a - b).max(axis=1) / (np.abs(a).max(axis=1) + 1e-30)
            assert (err > 1e-9).mean() <= 2e-3, (integ, float(err.max()))
            cs, _ = sc.render(integ, 4, 0, 2, seed=2)
            g, _, _ = gs.render_sums(integ, 4, 0, 2, seed=2)
            assert abs(g.sum() - cs.sum()) <= 1e-6 * abs(cs.sum()) + 1e-9
        gs.close()
        sc.close()
