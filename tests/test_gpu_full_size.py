"""BASELINE.json-size checks on the GPU (config 2: 1 002 530 primitives, 1920x1080).  The oracle only sees bounded
subsets here; the rest is covered by size-independent properties."""
import numpy as np
import pytest

from oracle import bindings as ob
from take_b200 import api, scenes

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def big(gpu_lib):
    flat = scenes.heightfield().flat()
    gs = api.GpuScene(flat)
    yield flat, gs
    gs.close()


def test_fast_equals_exact_on_two_million_primaries(big, oracle_lib):
    flat, gs = big
    assert flat.num_prims == 1002530
    sc = oracle_lib.load(flat)
    H, W = flat.height, flat.width
    py, px = np.mgrid[0:H, 0:W]
    rays = sc.primary_rays(px.ravel(), py.ravel(), seed=123)
    pe, te, uve = gs.intersect(rays, exact=True)
    pf, tf, uvf = gs.intersect(rays, exact=False)
    assert np.array_equal(pe, pf) and np.array_equal(te, tf) and np.array_equal(uve, uvf)
    assert 0.2 < (pe >= 0).mean() < 0.9
    # oracle on a bounded subset (every 97th ray) and on secondary rays from those hits
    sub = rays[::97]
    op, ot, _ = sc.intersect(sub)
    assert np.array_equal(op, pf[::97]) and np.array_equal(ot[op >= 0], tf[::97][op >= 0])
    sec = ob.secondary_rays(sub, ot, op, seed=8)
    op2, ot2, _ = sc.intersect(sec)
    p2, t2, _ = gs.intersect(sec)
    assert np.array_equal(op2, p2) and np.array_equal(ot2[op2 >= 0], t2[op2 >= 0])
    sec[:, 7] = 40.0
    assert np.array_equal(gs.occluded(sec), sc.occluded(sec))
    sc.close()


def test_full_frame_properties(big, oracle_lib):
    flat, gs = big
    a, a2, st = gs.render_sums("one_sample_mis", 5, 0, 2, seed=3)
    lo, _, _ = gs.render_sums("one_sample_mis", 5, 0, 1, seed=3)
    hi, _, _ = gs.render_sums("one_sample_mis", 5, 1, 2, seed=3)
    assert np.abs(lo + hi - a).max() <= 1e-12 * np.abs(a).max()
    assert st["samples"] == 2 * flat.width * flat.height
    assert st["miss_after_light_sample"] == 0
    assert np.isfinite(a).all() and (a >= 0).all()
    # oracle on every 64th row of the first sample
    sc = oracle_lib.load(flat)
    cs, _ = sc.render("one_sample_mis", 5, 0, 1, seed=3, row_begin=5, row_step=64)
    rows = np.arange(5, flat.height, 64)
    err = np.abs(lo[rows] - cs[rows]).max(axis=2) / (np.abs(cs[rows]).max(axis=2) + 1e-12)
    assert (err > 1e-9).mean() <= 1e-3
    sc.close()
