"""CPU emulation of the fast traversal's conservative FP32 slab test (take_b200/csrc/traverse.cuh: trace_fast) on the
host-built tree: for every ray the oracle says hits primitive P at distance t, every box on the root-to-leaf path of P
must pass the emulated test with the best distance already shrunk to t -- i.e. the traversal can never cull the
primitive the FP64 leaf test would accept.  Includes axis-parallel rays (zero direction components) and rays with
denormal-small components, where a naive reciprocal would produce inf - inf."""
import numpy as np
import pytest

from oracle import bindings as ob
from take_b200 import api, scenes, sceneio

from conftest import all_pixel_rays

F = np.float32
SLACK = F(1.00000191)


def fma32(a, b, c):
    # float32 fma emulated through float64: the product of two float32 is exact in float64
    with np.errstate(invalid="ignore", over="ignore"):
        return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(F)


def ray_setup(rays, abs_max):
    o = rays[:, 0:3].astype(F)
    with np.errstate(divide="ignore"):
        idir = np.clip(F(1) / rays[:, 3:6].astype(F), F(-1e18), F(1e18))
    delta = F(1.9073486e-6) * np.maximum(np.abs(o).max(axis=1), F(abs_max))
    ol = -(o + delta[:, None]) * idir
    oh = -(o - delta[:, None]) * idir
    return idir, ol, oh


def box_pass(lo, hi, idir, ol, oh, tmin_f, tbest_f):
    """lo, hi: [n,3] float32 child boxes.  Mirrors the kernel: NaN-ignoring fminf/fmaxf."""
    a = fma32(lo, idir, ol)
    b = fma32(hi, idir, oh)
    tn = np.fmax.reduce(np.fmin(a, b), axis=1)
    tf = np.fmin.reduce(np.fmax(a, b), axis=1)
    tn = np.fmax(tn, tmin_f)
    tf = np.fmin(tf, tbest_f)
    with np.errstate(invalid="ignore"):
        return tn <= tf * SLACK


def float_round_up(x):
    f = x.astype(F)
    return np.where(f.astype(np.float64) < x, np.nextafter(f, F(np.inf)), f)


def subtree_ranges(nodes):
    """(first, end) leaf-slot range below each child of each inner node."""
    n = len(nodes)
    rng = np.zeros((n, 2, 2), np.int64)
    order, stack = [], [0]
    while stack:
        i = stack.pop()
        order.append(i)
        for k in (0, 1):
            c = int(nodes[i][f"child{k}"])
            if c >= 0:
                stack.append(c)
    for i in reversed(order):
        for k in (0, 1):
            c = int(nodes[i][f"child{k}"])
            if c >= 0:
                rng[i, k] = (min(rng[c, 0, 0], rng[c, 1, 0]), max(rng[c, 0, 1], rng[c, 1, 1]))
            elif nodes[i][f"c{k}lox"] <= nodes[i][f"c{k}hix"]:
                code = ~c
                rng[i, k] = (code >> 3, (code >> 3) + (code & 7) + 1)
            else:
                rng[i, k] = (1 << 40, -1)       # empty child
    return rng


def check_paths(hb, rays, prim, t):
    nodes = hb["fast_nodes"]
    slot_of = np.empty(len(hb["leaf_prims"]), np.int64)
    slot_of[hb["leaf_prims"]] = np.arange(len(slot_of))
    ranges = subtree_ranges(nodes)
    hit = prim >= 0
    rays, slot, t = rays[hit], slot_of[prim[hit]], t[hit]
    idir, ol, oh = ray_setup(rays, hb["abs_max"])
    tmin_f = np.nextafter(rays[:, 6].astype(F), F(-np.inf))
    tbest_f = float_round_up(t)
    node = np.zeros(len(rays), np.int64)
    alive = np.ones(len(rays), bool)
    lo_names = [["c0lox", "c0loy", "c0loz"], ["c1lox", "c1loy", "c1loz"]]
    hi_names = [["c0hix", "c0hiy", "c0hiz"], ["c1hix", "c1hiy", "c1hiz"]]
    steps = 0
    while alive.any():
        idx = np.nonzero(alive)[0]
        nd = node[idx]
        k = np.where((slot[idx] >= ranges[nd, 0, 0]) & (slot[idx] < ranges[nd, 0, 1]), 0, 1)
        assert ((slot[idx] >= ranges[nd, k, 0]) & (slot[idx] < ranges[nd, k, 1])).all()
        lo = np.stack([np.where(k == 0, nodes[lo_names[0][a]][nd], nodes[lo_names[1][a]][nd]) for a in range(3)], axis=1)
        hi = np.stack([np.where(k == 0, nodes[hi_names[0][a]][nd], nodes[hi_names[1][a]][nd]) for a in range(3)], axis=1)
        ok = box_pass(lo, hi, idir[idx], ol[idx], oh[idx], tmin_f[idx], tbest_f[idx])
        assert ok.all(), f"{(~ok).sum()} boxes on the path to the true hit would be culled"
        child = np.where(k == 0, nodes["child0"][nd], nodes["child1"][nd])
        node[idx] = child
        alive[idx] = child >= 0
        steps += 1
        assert steps < 200
    return int(hit.sum())


def check_paths_wide(hb, rays, prim, t):
    """Same property on the 4-wide collapse: the child whose subtree holds the true hit passes the emulated test."""
    wide = hb["wide_nodes"]
    n = len(wide)
    # leaf-slot range below each child
    rng_ = np.zeros((n, 4, 2), np.int64)
    order, stack = [], [0]
    while stack:
        i = stack.pop()
        order.append(i)
        stack.extend(int(c) for c in wide[i]["child"] if 0 <= c < 0x7fffffff)
    for i in reversed(order):
        for k in range(4):
            c = int(wide[i]["child"][k])
            if c == 0x7fffffff:
                rng_[i, k] = (1 << 40, -1)
            elif c >= 0:
                sub = rng_[c][rng_[c][:, 1] >= 0]
                rng_[i, k] = (sub[:, 0].min(), sub[:, 1].max())
            else:
                code = ~c
                rng_[i, k] = (code >> 3, (code >> 3) + (code & 7) + 1)
    slot_of = np.empty(len(hb["leaf_prims"]), np.int64)
    slot_of[hb["leaf_prims"]] = np.arange(len(slot_of))
    hit = prim >= 0
    rays, slot, t = rays[hit], slot_of[prim[hit]], t[hit]
    idir, ol, oh = ray_setup(rays, hb["abs_max"])
    tmin_f = np.nextafter(rays[:, 6].astype(F), F(-np.inf))
    tbest_f = float_round_up(t)
    node = np.zeros(len(rays), np.int64)
    alive = np.ones(len(rays), bool)
    steps = 0
    while alive.any():
        idx = np.nonzero(alive)[0]
        nd = node[idx]
        inside = (slot[idx, None] >= rng_[nd, :, 0]) & (slot[idx, None] < rng_[nd, :, 1])
        assert (inside.sum(axis=1) == 1).all()
        k = inside.argmax(axis=1)
        pick = lambda name: wide[name][nd, k]
        lo = np.stack([pick("lox"), pick("loy"), pick("loz")], axis=1)
        hi = np.stack([pick("hix"), pick("hiy"), pick("hiz")], axis=1)
        ok = box_pass(lo, hi, idir[idx], ol[idx], oh[idx], tmin_f[idx], tbest_f[idx])
        assert ok.all(), f"{(~ok).sum()} wide boxes on the path to the true hit would be culled"
        child = wide["child"][nd, k]
        node[idx] = child
        alive[idx] = child >= 0
        steps += 1
        assert steps < 200
    return int(hit.sum())


def axis_parallel_rays(flat, rng, n):
    lo, hi = flat.positions.min(axis=0), flat.positions.max(axis=0)
    ext = hi - lo
    rays = np.zeros((n, 8))
    rays[:, 0:3] = lo - 0.3 * ext + rng.uniform(0, 1.6, (n, 3)) * ext
    kind = rng.integers(0, 4, n)
    axis = rng.integers(0, 3, n)
    sign = rng.choice([-1.0, 1.0], n)
    d = np.zeros((n, 3))
    d[np.arange(n), axis] = sign
    other = (axis + 1) % 3
    tiny = np.choose(kind, [0.0, 1e-30, 1e-20, -1e-42])       # exact zero, denormal-in-float32, tiny, negative denormal
    d[np.arange(n), other] = tiny
    d[kind == 3, (axis[kind == 3] + 2) % 3] = -0.0
    rays[:, 3:6] = d
    rays[:, 6], rays[:, 7] = 1e-7, np.inf
    return rays


@pytest.mark.parametrize("name", ["cornell", "heightfield", "spheres", "multi_light"])
def test_true_hit_is_never_culled(oracle_lib, name):
    make = {"cornell": lambda: scenes.cornell_box(64, 64, 1), "heightfield": lambda: scenes.heightfield(96, 96, 54, 1),
            "spheres": lambda: scenes.sphere_room(64, 64, 1), "multi_light": lambda: scenes.multi_light(64, 40, 1, n_side=5)}[name]
    flat = make().flat()
    hb = api.host_build(flat)
    sc = oracle_lib.load(flat)
    rng = np.random.default_rng(11)
    total = 0
    for jitter in (True, False):
        rays = all_pixel_rays(sc, jitter=jitter)
        p, t, _ = sc.intersect(rays)
        total += check_paths(hb, rays, p, t)
        assert check_paths_wide(hb, rays, p, t) == int((p >= 0).sum())
        sec = ob.secondary_rays(rays, t, p, seed=4)
        p2, t2, _ = sc.intersect(sec)
        total += check_paths(hb, sec, p2, t2)
        check_paths_wide(hb, sec, p2, t2)
    ap = axis_parallel_rays(flat, rng, 20000)
    p, t, _ = sc.intersect(ap)
    n_ap = check_paths(hb, ap, p, t)
    check_paths_wide(hb, ap, p, t)
    assert n_ap > 100, "axis-parallel set must actually hit something"
    # far-away origins: the per-ray padding scales with |origin|
    far = all_pixel_rays(sc, jitter=True)
    shift = 1000.0 * np.abs(flat.positions).max()
    far[:, 0:3] -= far[:, 3:6] * shift
    p, t, _ = sc.intersect(far)
    total += check_paths(hb, far, p, t)
    assert total > 1000
