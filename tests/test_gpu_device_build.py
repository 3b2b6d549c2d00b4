"""The device builder of the fast tree (take_b200/csrc/bvh_device.cuh): structure of the tree it leaves on the
device, identical answers to a scene whose tree the host's SAH builder made, and the background build of the
reference-order tree."""
import numpy as np
import pytest

from oracle import bindings as ob
from take_b200 import api, scenes, sceneio

from conftest import all_pixel_rays

pytestmark = pytest.mark.gpu


def prim_bounds(flat):
    n = flat.num_prims
    sph = (flat.prim_flags & sceneio.PRIM_SPHERE) != 0
    lo = np.empty((n, 3)); hi = np.empty((n, 3))
    P = flat.positions[flat.indices[~sph]]
    lo[~sph], hi[~sph] = P.min(axis=1), P.max(axis=1)
    if sph.any():
        s = flat.spheres[flat.indices[sph, 0]]
        lo[sph], hi[sph] = s[:, :3] - s[:, 3:4], s[:, :3] + s[:, 3:4]
    return lo, hi


def check_tree(flat, wide, prims, max_leaf=4):
    """Every primitive sits in exactly one leaf slot, every leaf slot is reached exactly once, every child box strictly
    contains the FP64 bounds of what is below it, no node is reached twice.  Iterative (no recursion limit); returns the depth."""
    n = flat.num_prims
    assert len(prims) == n and np.array_equal(np.sort(prims), np.arange(n))
    lo, hi = prim_bounds(flat)
    if n == 0:
        assert len(wide) == 1 and (wide[0]["child"] == api.WIDE_EMPTY).all()
        return 1
    # bounds of leaf slot ranges via prefix structures: min / max over contiguous slots
    slo, shi = lo[prims], hi[prims]
    covered = np.zeros(n, np.int32)
    seen = np.zeros(len(wide), bool)
    # post-order accumulation of subtree bounds: first a pre-order list, then fold it backwards
    order, stack = [], [(0, 1)]
    depth = 0
    while stack:
        i, d = stack.pop()
        assert not seen[i]
        seen[i] = True
        depth = max(depth, d)
        order.append(i)
        for k in range(4):
            c = int(wide[i]["child"][k])
            if c != api.WIDE_EMPTY and c >= 0:
                stack.append((c, d + 1))
    blo = np.full((len(wide), 3), np.inf); bhi = np.full((len(wide), 3), -np.inf)
    for i in reversed(order):
        for k in range(4):
            c = int(wide[i]["child"][k])
            if c == api.WIDE_EMPTY:
                continue
            clo = np.array([wide[i][a][k] for a in ("lox", "loy", "loz")], np.float64)
            chi = np.array([wide[i][a][k] for a in ("hix", "hiy", "hiz")], np.float64)
            if c >= 0:
                l, h = blo[c], bhi[c]
            else:
                code = ~c
                first, count = code >> 3, (code & 7) + 1
                assert 1 <= count <= max_leaf and count == int(wide[i]["count"][k])
                covered[first:first + count] += 1
                l, h = slo[first:first + count].min(axis=0), shi[first:first + count].max(axis=0)
            assert (clo < l).all() and (chi > h).all(), (i, k)
            blo[i], bhi[i] = np.minimum(blo[i], l), np.maximum(bhi[i], h)
    assert seen.all() and (covered == 1).all()
    return depth


def test_device_tree_structure(small_scene, gpu_lib):
    name, _, flat = small_scene
    gs = api.GpuScene(flat)
    try:
        t = gs.create_timings()
        assert t["device_built"] == 1
        wide, prims = gs.debug_tree()
        depth = check_tree(flat, wide, prims)
        info = gs.info()
        assert info["fast_nodes"] == len(wide) and info["fast_tree_depth"] == depth
        assert gs.create_timings()["reference_tree_pending"] == 0      # info() joined the background build
    finally:
        gs.close()


@pytest.mark.parametrize("make", [lambda: scenes.heightfield(160, 64, 36, 1), lambda: scenes.instanced_spheres(64, 36, 1, copies_side=3, subdiv=4, n_lights=4),
                                  lambda: scenes.ibl_scene(32, 32, 1, n_objects=9, env_size=(16, 8))])
def test_device_tree_structure_medium(gpu_lib, make):
    flat = make().flat()
    gs = api.GpuScene(flat)
    try:
        wide, prims = gs.debug_tree()
        depth = check_tree(flat, wide, prims)
        assert depth <= 31 and len(wide) < flat.num_prims
    finally:
        gs.close()


def test_degenerate_inputs_build_on_the_device(gpu_lib, oracle_lib):
    """Empty scene, one primitive, all primitives identical (every Morton code equal), collinear centroids, two far-apart
    clusters of very different scale."""
    cases = []
    b = scenes.SceneBuilder(8, 8, (0, 0, 5), (0, 0, 0), background=(0.1, 0.2, 0.3))
    cases.append(b)
    b = scenes.SceneBuilder(8, 8, (0, 0, 5), (0, 0, 0))
    m = b.material(sceneio.MAT_DIFFUSE, (0.5, 0.5, 0.5))
    b.mesh([(-1, -1, 0), (1, -1, 0), (0, 1, 0)], [[0, 1, 2]], [(0, 0, 1)] * 3, None, m)
    cases.append(b)
    b = scenes.SceneBuilder(8, 8, (0, 0, 5), (0, 0, 0))
    m = b.material(sceneio.MAT_DIFFUSE, (0.5, 0.5, 0.5))
    b.mesh([(-1, -1, 0), (1, -1, 0), (0, 1, 0)], [[0, 1, 2]] * 37, [(0, 0, 1)] * 3, None, m)       # 37 copies of one triangle
    cases.append(b)
    b = scenes.SceneBuilder(8, 8, (0, 0, 5), (0, 0, 0))
    m = b.material(sceneio.MAT_DIFFUSE, (0.5, 0.5, 0.5))
    P = [(x + dx, dy, 0.0) for x in range(50) for dx, dy in ((0, 0), (0.5, 0), (0.25, 0.5))]
    b.mesh(P, [[3 * i, 3 * i + 1, 3 * i + 2] for i in range(50)], [(0, 0, 1)] * len(P), None, m)        # a row of triangles
    cases.append(b)
    b = scenes.SceneBuilder(8, 8, (0, 0, 5), (0, 0, 0))
    m = b.material(sceneio.MAT_DIFFUSE, (0.5, 0.5, 0.5))
    rng = np.random.default_rng(0)
    P = np.concatenate([rng.uniform(-1e-3, 1e-3, (300, 3)), rng.uniform(-1e-3, 1e-3, (300, 3)) + 1e4])
    b.mesh(P, np.arange(600).reshape(-1, 3), np.tile([0, 0, 1.0], (600, 1)), None, m)
    cases.append(b)
    for b in cases:
        flat = b.flat()
        gs, sc = api.GpuScene(flat), oracle_lib.load(flat)
        try:
            wide, prims = gs.debug_tree()
            check_tree(flat, wide, prims)
            rays = all_pixel_rays(sc, seed=3)
            op, ot, _ = sc.intersect(rays)
            hit = op >= 0
            p, t, _ = gs.intersect(rays, exact=True)          # the reference's own traversal: bit-identical, ties included
            assert np.array_equal(p, op) and np.array_equal(t[hit], ot[hit])
            p, t, _ = gs.intersect(rays, exact=False)
            assert np.array_equal(p >= 0, hit) and np.array_equal(t[hit], ot[hit])
            if flat.num_prims != 37:      # (37 coincident triangles: every hit is a 37-way exact tie, where the reference's
                assert np.array_equal(p, op)   # answer also depends on its 1-ulp box culling -- SURVEY.md 7.2 -- so only t is compared)
        finally:
            gs.close()
            sc.close()


def test_device_built_scene_equals_host_built_scene(small_scene, gpu_lib, monkeypatch):
    """Any tree with conservative boxes gives the same closest hits, so a scene whose fast tree the device built answers
    bit for bit like one built by the host's SAH builder -- intersections, occlusion and whole renders."""
    name, _, flat = small_scene
    dev = api.GpuScene(flat)
    monkeypatch.setenv("TAKE_DEVICE_BUILD", "0")
    host = api.GpuScene(flat)
    monkeypatch.delenv("TAKE_DEVICE_BUILD")
    try:
        assert dev.create_timings()["device_built"] == 1 and host.create_timings()["device_built"] == 0
        sc = ob.OracleLib().load(flat)
        rays = all_pixel_rays(sc, seed=9)
        a, b = dev.intersect(rays), host.intersect(rays)
        assert all(np.array_equal(x, y) for x, y in zip(a, b))
        sec = ob.secondary_rays(rays, a[1], a[0], seed=4)
        assert all(np.array_equal(x, y) for x, y in zip(dev.intersect(sec), host.intersect(sec)))
        sec[:, 7] = 0.4 * np.abs(flat.positions).max()
        assert np.array_equal(dev.occluded(sec), host.occluded(sec))
        for integ in api.INTEGRATORS:
            s1, q1, st1 = dev.render_sums(integ, 5, 0, 3, seed=2)
            s2, q2, st2 = host.render_sums(integ, 5, 0, 3, seed=2)
            assert np.array_equal(s1, s2) and np.array_equal(q1, q2)
            for k in ("extend_rays", "shadow_rays", "shaded"):
                assert st1[k] == st2[k]
        sc.close()
    finally:
        dev.close()
        host.close()


def test_scene_create_does_not_wait_for_the_reference_tree(gpu_lib):
    """config 2: take_gpu_scene_create returns while the reference-order tree is still being built by the host thread; the
    first query joins it.  Prints the phase timings (the figures DESIGN.md quotes)."""
    flat = scenes.heightfield().flat()
    gs = api.GpuScene(flat)
    try:
        t = gs.create_timings()
        print("scene_create phases (ms):", {k: round(v, 1) for k, v in t.items()})
        assert t["device_built"] == 1
        rays = api.make_rays([[0, 160, 240]], [[0.0137, -0.5547, -0.8262]])   # (not along a mesh edge: those rays slip through, shape.cpp:65,71)
        p, _, _ = gs.intersect(rays, exact=True)          # needs the reference tree
        assert gs.create_timings()["reference_tree_pending"] == 0 and p[0] >= 0
        info = gs.info()
        print("fast tree:", int(info["fast_nodes"]), "wide nodes, depth", int(info["fast_tree_depth"]), "SAH cost", round(info["sah_cost"], 2),
              "; reference-order tree", round(info["build_ms_reference_tree"], 1), "ms in the background")
    finally:
        gs.close()


def test_render_ahead_of_the_reference_tree(gpu_lib, monkeypatch):
    """take_gpu_render does not wait for the background reference-order tree: it renders at once, counts the leaf tests in which
    a tie-break rank decided anything, and repeats the call after the join only if there were any.  TAKE_REF_DELAY_MS keeps the
    tree 'under construction' long enough to observe both outcomes."""
    import time
    # (a) an ordinary scene: no exact ties -> the provisional image stands, and it is the image of a host-built scene
    flat = scenes.cornell_box(64, 64, 4, materials="mixed").flat()
    monkeypatch.setenv("TAKE_DEVICE_BUILD", "0")
    host = api.GpuScene(flat)
    monkeypatch.delenv("TAKE_DEVICE_BUILD")
    want = {integ: host.render_sums(integ, 5, 0, 4, seed=11) for integ in ("mis", "one_sample_mis")}
    host.close()
    monkeypatch.setenv("TAKE_REF_DELAY_MS", "1500")
    gs = api.GpuScene(flat)
    t0 = time.perf_counter()
    got = {integ: gs.render_sums(integ, 5, 0, 4, seed=11) for integ in ("mis", "one_sample_mis")}
    dt = time.perf_counter() - t0
    st = gs.provisional_stats()
    assert gs.create_timings()["reference_tree_pending"] == 1 and dt < 1.0        # still pending: nothing waited for it
    assert st == {"provisional_renders": 2, "provisional_reruns": 0}
    for integ in want:
        assert np.array_equal(got[integ][0], want[integ][0]) and np.array_equal(got[integ][1], want[integ][1])
        assert got[integ][2]["extend_rays"] == want[integ][2]["extend_rays"]
    # accumulating entry point: the snapshot / restore must leave earlier contents alone
    p, _, _ = gs.intersect(api.make_rays([[0, 1, 3.8]], [[0.01, 0.02, -1.0]]), exact=True)   # joins the tree
    assert gs.create_timings()["reference_tree_pending"] == 0
    again = gs.render_sums("mis", 5, 0, 4, seed=11)
    assert np.array_equal(again[0], want["mis"][0]) and gs.provisional_stats()["provisional_renders"] == 2
    gs.close()
    # (b) 37 coincident triangles: every hit is an exact tie -> the provisional render is thrown away and repeated
    b = scenes.SceneBuilder(16, 16, (0, 0, 5), (0, 0, 0), background=(0.1, 0.1, 0.1))
    m = b.material(sceneio.MAT_DIFFUSE, (0.5, 0.5, 0.5))
    black = b.material(sceneio.MAT_DIFFUSE, (0, 0, 0))
    b.mesh([(-1, -1, 0), (1, -1, 0), (0, 1, 0)], [[0, 1, 2]] * 37, [(0, 0, 1)] * 3, None, m)
    b.quad((-1, 3, -1), (1, 3, -1), (1, 3, 1), (-1, 3, 1), black, radiance=(5, 5, 5))
    flat = b.flat()
    monkeypatch.delenv("TAKE_REF_DELAY_MS")
    monkeypatch.setenv("TAKE_DEVICE_BUILD", "0")
    host = api.GpuScene(flat)
    monkeypatch.delenv("TAKE_DEVICE_BUILD")
    want = host.render_sums("mis", 5, 0, 3, seed=2)
    host.close()
    monkeypatch.setenv("TAKE_REF_DELAY_MS", "300")
    gs = api.GpuScene(flat)
    got = gs.render_sums("mis", 5, 0, 3, seed=2)
    assert gs.provisional_stats() == {"provisional_renders": 1, "provisional_reruns": 1}
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
    gs.close()
