"""Russian roulette (EXTENSION: the reference's README lists it as a goal, its integrators have none; include/take_gpu.h
TAKE_RENDER_RUSSIAN_ROULETTE).  Parity unpinned by nature: the CPU restatement (oracle/take_oracle.cpp: rr_survives) is our
own; what is checked is that it is unbiased, that it saves rays, that it is OFF unless asked for, and -- on the GPU -- that
the device reproduces the restatement sample by sample."""
import numpy as np
import pytest

from take_b200 import api, scenes


def test_roulette_is_unbiased_and_saves_rays(oracle_lib):
    flat = scenes.cornell_box(24, 24, 4, materials="mixed").flat()
    sc = oracle_lib.load(flat)
    n, depth = 192, 30
    for integ in ("mis", "one_sample_mis"):
        sc.set_russian_roulette(0)
        a, a2, st_a = sc.render(integ, depth, 0, n, seed=11, stats=True)
        b, _, _ = sc.render(integ, depth, 0, n, seed=11, stats=True)
        assert np.array_equal(a, b)                                   # off: deterministic, the reference's integrator
        sc.set_russian_roulette(3)
        r, r2, st_r = sc.render(integ, depth, 0, n, seed=12, stats=True)
        assert st_r[0] < 0.97 * st_a[0]                               # fewer extend rays at depth 30 (most paths end earlier anyway)
        # same expectation: image sums agree within 4 sigma of the two estimates' standard errors, per channel
        mu_a, mu_r = a / n, r / n
        var = (np.maximum(a2 / n - mu_a ** 2, 0) + np.maximum(r2 / n - mu_r ** 2, 0)) / (n - 1)
        for c in range(3):
            assert abs(mu_a[..., c].sum() - mu_r[..., c].sum()) <= 4 * np.sqrt(var[..., c].sum()), (integ, c)
    sc.set_russian_roulette(0)
    sc.close()


@pytest.mark.gpu
@pytest.mark.parametrize("integrator", ["mis", "raw", "one_sample_mis", "one_sample_mis_power"])
def test_gpu_roulette_matches_restatement(gpu_lib, oracle_lib, integrator):
    for flat in (scenes.cornell_box(32, 32, 4, materials="mixed").flat(), scenes.multi_light(40, 24, 4, n_side=4).flat()):
        gs, sc = api.GpuScene(flat), oracle_lib.load(flat)
        try:
            for start in (3, 1):
                sc.set_russian_roulette(start)
                gs.rr_start = start
                cs, cs2, cst = sc.render(integrator, 12, 0, 4, seed=7, stats=True)
                s, s2, st = gs.render_sums(integrator, 12, 0, 4, seed=7, flags=api.RENDER_RUSSIAN_ROULETTE)
                bad = np.abs(s - cs).max(axis=2) > 1e-9 * (np.abs(cs).max(axis=2) + 1e-12)
                assert bad.mean() <= 2e-3, (integrator, start)
                if not bad.any():
                    assert st["extend_rays"] == cst[0] and st["shadow_rays"] == cst[1] and st["shaded"] == cst[2]
            # without the flag nothing changes, whatever rr_start says
            sc.set_russian_roulette(0)
            cs, _ = sc.render(integrator, 12, 0, 2, seed=7)
            s, _, _ = gs.render_sums(integrator, 12, 0, 2, seed=7)
            assert (np.abs(s - cs).max(axis=2) > 1e-9 * (np.abs(cs).max(axis=2) + 1e-12)).mean() <= 2e-3
        finally:
            gs.close()
            sc.close()
