"""Live cross-check of the CPU restatement against the unmodified reference compiled into oracle/_ref/ (skipped where
that library is absent).  Larger and more varied than the committed golden vectors."""
import tempfile

import numpy as np
import pytest

from oracle import bindings as ob
from take_b200.sceneio import FlatScene

from conftest import all_pixel_rays


def test_rng_preload_trick(ref_lib):
    # the harness feeds the reference's std::mt19937 with our counter-based stream; check it really does
    for seed, pixel, sample in [(0, 0, 0), (123, 77, 5), (2**40 + 3, 2**31 + 1, 2**33 + 9)]:
        assert ref_lib.rng_selfcheck(seed, pixel, sample, 300) == 0


def test_scene_generators_match_reference_parser(ref_lib, small_scene):
    name, builder, flat = small_scene
    with tempfile.TemporaryDirectory() as d:
        rs = ref_lib.load(builder.write(d))
        rs.dump(d + "/ref.takescene")
        assert flat.same_as(FlatScene.load(d + "/ref.takescene")) == []
        assert rs.num_prims == flat.num_prims and rs.num_lights == len(flat.lights)
        rs.close()


def test_port_equals_reference(ref_lib, oracle_lib, small_scene):
    name, builder, flat = small_scene
    with tempfile.TemporaryDirectory() as d:
        rs = ref_lib.load(builder.write(d))
        sc = oracle_lib.load(flat)
        rb, rl = rs.bvh()
        ob_, ol, root = sc.bvh()
        assert root == rs.root and np.array_equal(rb, ob_) and np.array_equal(rl, ol)
        for jitter in (True, False):  # pixel-centre rays include exact-tie cases on the symmetric rooms
            rays = all_pixel_rays(sc, jitter=jitter)
            rp, rt, rrec = rs.intersect(rays, records=True)   # also asserts shim == scene_intersect bitwise
            p, t, _, rec = sc.intersect(rays, records=True)
            assert np.array_equal(p, rp) and np.array_equal(t, rt) and np.array_equal(rec, rrec)
        sec = ob.secondary_rays(rays, rt, rp, seed=3)
        rp, rt, rrec = rs.intersect(sec, records=True)
        p, t, _, rec = sc.intersect(sec, records=True)
        assert np.array_equal(p, rp) and np.array_equal(t, rt) and np.array_equal(rec, rrec)
        sec[:, 7] = 0.37 * np.abs(flat.positions).max()
        assert np.array_equal(sc.occluded(sec), rs.occluded(sec))
        for integ in ob.INTEGRATORS:
            for max_depth in (5, 0, -1):
                rsum, rsq = rs.render(integ, max_depth, 1, 3, seed=99)
                s, s2, st = sc.render(integ, max_depth, 1, 3, seed=99, stats=True)
                assert np.array_equal(s, rsum) and np.array_equal(s2, rsq), (name, integ, max_depth)
                assert st[5] == 0   # no light-aimed ray missed everything (the reference's UB case)
        rs.close()
        sc.close()


def test_default_max_depth_50(ref_lib, oracle_lib):
    from take_b200 import scenes
    b = scenes.cornell_box(16, 16, 1, materials="mixed")
    with tempfile.TemporaryDirectory() as d:
        rs = ref_lib.load(b.write(d))
        sc = oracle_lib.load(b.flat())
        for integ in ob.INTEGRATORS:
            a, _ = rs.render(integ, 50, 0, 2, seed=1)
            c, _ = sc.render(integ, 50, 0, 2, seed=1)
            assert np.array_equal(a, c)
