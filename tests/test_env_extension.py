"""Environment-map lighting: an EXTENSION with no reference implementation (the reference's only environment is the
constant <background>, src/scene.h:23, src/parse/parse_scene.cpp:1006-1013).  "Parity unpinned" except for the constant
case, which IS pinned: a constant map that is not light-sampled must reproduce the <background> render bit for bit.
Everything else is checked for internal consistency of our own FP64 restatement: the sampling pdf integrates to one and
matches what the sampler draws, the light-sampled and the background-only estimators agree in the mean, and the white
furnace is exact."""
import numpy as np
import pytest

from take_b200 import scenes, sceneio


def furnace_scene(res=24, sample=True, albedo=0.6):
    b = scenes.SceneBuilder(res, res, (0, 0, 4), (0, 0, 0), (0, 1, 0), 40.0, 4, (0, 0, 0))
    m = b.material(sceneio.MAT_DIFFUSE, (albedo, albedo, albedo))
    b.sphere((0, 0, 0), 1.0, m)
    b.environment(np.ones((8, 16, 3)), sample=sample)
    return b


@pytest.mark.parametrize("integrator", ["mis", "raw", "one_sample_mis"])
def test_constant_environment_equals_background(oracle_lib, integrator):
    """Pinned sub-case: constant map, not light-sampled  ==  the reference's <background> (which the oracle reproduces
    bit for bit, tests/test_oracle_vs_reference.py)."""
    for make in (lambda: scenes.cornell_box(32, 32, 2, materials="mixed"), lambda: scenes.heightfield(24, 40, 24, 2)):
        b = make()
        bg = np.array([0.25, 0.5, 0.125])
        b.background = bg
        ref = oracle_lib.load(b.flat())
        b.environment(np.broadcast_to(bg, (4, 8, 3)).copy(), sample=False)
        env = oracle_lib.load(b.flat())
        a, a2 = ref.render(integrator, 5, 0, 3, seed=5)
        c, c2 = env.render(integrator, 5, 0, 3, seed=5)
        assert np.array_equal(a, c) and np.array_equal(a2, c2)


def test_sampling_pdf_is_consistent(oracle_lib):
    rng = np.random.default_rng(0)
    env = rng.uniform(0.0, 1.0, (12, 24, 3)) ** 4 * 20
    env[3, 5] = 500.0                      # a "sun"
    env[8:, :] = 0.0                       # black ground: never sampled
    b = furnace_scene()
    b.environment(env, sample=True)
    sc = oracle_lib.load(b.flat())
    # (1) the pdf integrates to one over the sphere (midpoint rule on a fine lat-long grid)
    n_t, n_p = 12 * 6, 24 * 6              # 6 x 6 midpoints per texel, aligned with the texel grid
    theta = (np.arange(n_t) + 0.5) / n_t * np.pi
    phi = (np.arange(n_p) + 0.5) / n_p * 2 * np.pi - np.pi
    total = 0.0
    for t in theta:
        d = np.stack([np.sin(t) * np.cos(phi), np.full(n_p, np.cos(t)), -np.sin(t) * np.sin(phi)], axis=1)
        total += sum(sc.env_eval(x)[1] for x in d) * np.sin(t)
    total *= (np.pi / n_t) * (2 * np.pi / n_p)
    assert abs(total - 1.0) < 1e-9
    # (2) sampled directions carry the pdf that env_eval reports for them, hit only non-black texels,
    #     and E[Le / pdf] equals the integral of Le over the sphere
    acc = np.zeros(3)
    n = 20000
    for u1, u2 in rng.uniform(0, 1, (n, 2)):
        d, pdf = sc.env_sample(u1, u2)
        le, pdf2 = sc.env_eval(d)
        assert pdf > 0 and abs(np.linalg.norm(d) - 1) < 1e-12
        assert abs(pdf - pdf2) <= 1e-9 * pdf
        assert le.sum() > 0
        acc += le / pdf
    tj = (np.arange(12) + 0.5) / 12 * np.pi
    exact = (env * (np.cos(tj - np.pi / 24) - np.cos(tj + np.pi / 24))[:, None, None]).sum(axis=(0, 1)) * (2 * np.pi / 24)
    assert np.allclose(acc / n, exact, rtol=0.05)


@pytest.mark.parametrize("integrator", ["mis", "raw", "one_sample_mis"])
def test_white_furnace(oracle_lib, integrator):
    """A convex diffuse body of albedo rho inside a constant unit environment shows exactly rho (no inter-reflection),
    whatever the sampling strategy; the background shows 1."""
    for sample in (False, True):
        sc = oracle_lib.load(furnace_scene(sample=sample).flat())
        n = 256
        s, _ = sc.render(integrator, 5, 0, n, seed=3)
        mean = s / n
        centre = mean[10:14, 10:14].mean()
        assert abs(centre - 0.6) < 0.02, (integrator, sample, centre)
        assert np.allclose(mean[0, 0], 1.0)


def test_light_sampled_and_background_only_estimators_agree(oracle_lib):
    """Sampling the environment changes the variance, not the expectation."""
    rng = np.random.default_rng(2)
    env = np.full((16, 32, 3), 0.3)
    env[2, 9] = 400.0
    means = {}
    for sample in (False, True):
        b = scenes.cornell_box(20, 20, 4)
        b.meshes = [m for m in b.meshes if m.get("radiance") is None][:5] + b.meshes[6:]   # open room, no area light
        b.meshes = [m for i, m in enumerate(b.meshes) if i != 1]                             # remove the ceiling
        b.environment(env, sample=sample)
        sc = oracle_lib.load(b.flat())
        n = 4000 if not sample else 600
        s, s2 = sc.render("mis", 3, 0, n, seed=11 + sample)
        mu = s / n
        var = np.maximum(s2 / n - mu ** 2, 0) / n
        means[sample] = (mu.mean(), np.sqrt(var.sum()) / mu.size)
    (m0, e0), (m1, e1) = means[False], means[True]
    assert abs(m0 - m1) <= 4 * np.hypot(e0, e1) + 0.01 * m1, (means,)
    assert e1 * np.sqrt(600) < e0 * np.sqrt(4000)          # importance sampling does reduce the per-sample variance


def test_xml_writer_refuses_environment():
    with pytest.raises(ValueError):
        furnace_scene().write("/tmp/should_not_exist_env")


def test_ibl_scene_builds(oracle_lib):
    b = scenes.ibl_scene(48, 48, 2, n_objects=9, env_size=(64, 32))
    flat = b.flat()
    assert flat.env is not None and flat.env.shape == (32, 64, 3) and flat.env_sample
    sc = oracle_lib.load(flat)
    s, _ = sc.render("one_sample_mis", 5, 0, 2, seed=1)
    assert np.isfinite(s).all() and s.sum() > 0
