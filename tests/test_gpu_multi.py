"""Single-process multi-GPU render (take_gpu_render_multi): needs at least two CUDA devices, skipped otherwise.
(The one-process-per-GPU path is bench.py under torchrun; its host logic is covered on CPU by tests/test_dist_gloo.py.)"""
import numpy as np
import pytest

from take_b200 import api, scenes

pytestmark = pytest.mark.gpu


def test_render_multi_equals_single_gpu(gpu_lib):
    n = api.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    flat = scenes.cornell_box(64, 64, 4, materials="mixed").flat()
    gs = api.GpuScene(flat, device=0)
    a, a2, sa = gs.render_sums("mis", 5, 0, 7, seed=21)
    gs.close()
    for devices in ([0, 1], list(range(min(n, 4)))):
        b, b2, sb = api.render_multi(flat, devices, "mis", 5, 0, 7, seed=21)
        # same samples (streams are keyed by pixel and sample index), summed in a different order
        assert np.abs(a - b).max() <= 1e-12 * np.abs(a).max()
        assert np.abs(a2 - b2).max() <= 1e-12 * np.abs(a2).max()
        assert sb["samples"] == sa["samples"] and sb["extend_rays"] == sa["extend_rays"]


def test_persistent_multi_handle(gpu_lib):
    """take_gpu_multi_create / _render / _destroy: one handle, several renders (different integrators and sample ranges,
    with and without the sum of squares) -- each equal to the single-GPU render up to summation order; sample ranges
    rendered through the handle are additive."""
    n = api.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    flat = scenes.cornell_box(64, 64, 4, materials="mixed").flat()
    gs = api.GpuScene(flat, device=0)
    m = api.MultiGpuScene(flat, list(range(min(n, 8))))
    try:
        for integ, lo, hi in (("mis", 0, 9), ("one_sample_mis", 3, 20), ("raw", 1, 2)):
            a, a2, sa = gs.render_sums(integ, 5, lo, hi, seed=5)
            b, b2, sb = m.render_sums(integ, 5, lo, hi, seed=5)
            assert np.abs(a - b).max() <= 1e-12 * np.abs(a).max()
            assert np.abs(a2 - b2).max() <= 1e-12 * np.abs(a2).max()
            for k in ("samples", "extend_rays", "shadow_rays", "shaded"):
                assert sa[k] == sb[k], (integ, k)
        c, none, _ = m.render_sums("mis", 5, 0, 4, seed=5, sumsq=False)
        d, _, _ = m.render_sums("mis", 5, 4, 9, seed=5, sumsq=False)
        a, _, _ = gs.render_sums("mis", 5, 0, 9, seed=5)
        assert none is None and np.abs(c + d - a).max() <= 1e-12 * np.abs(a).max()
    finally:
        m.close()
        gs.close()
