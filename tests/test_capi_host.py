"""Host-side checks that need no GPU: the C-ABI library loads, exports every symbol include/take_gpu.h declares, refuses
to run without a device (no silent CPU path), and builds the acceleration structures the kernels rely on."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from take_b200 import api, scenes, sceneio

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "take_gpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(take_gpu_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol(gpu_lib):
    names = declared_functions()
    assert len(names) >= 15
    missing = [n for n in names if not hasattr(gpu_lib, n)]
    assert missing == []
    assert set(api.EXPORTS) <= set(names)
    assert b"sm_100a" in gpu_lib.take_gpu_version()


def test_struct_layouts_match_header():
    assert C.sizeof(sceneio.TakeCamera) == 8 + 10 * 8
    assert sceneio.MAT_DTYPE.itemsize == 80 and sceneio.LIGHT_DTYPE.itemsize == 56
    assert C.sizeof(api.TakeRenderOpts) == 40
    assert C.sizeof(api.TakeStats) == 7 * 8 + 7 * 8 + 4 * 8
    assert api.RAY_DTYPE.itemsize == 64 and api.HIT_DTYPE.itemsize == 32


def _cuda_available():
    try:
        return api.device_count() > 0
    except api.TakeGpuError:
        return False


def test_no_silent_cpu_fallback(gpu_lib):
    """Without a CUDA device every compute entry point must fail loudly."""
    if _cuda_available():
        pytest.skip("a CUDA device is present")
    flat = scenes.cornell_box(8, 8, 1).flat()
    with pytest.raises(api.TakeGpuError):
        api.GpuScene(flat)
    with pytest.raises(api.TakeGpuError):
        api.device_count()


def test_invalid_scene_is_rejected(gpu_lib):
    flat = scenes.cornell_box(8, 8, 1).flat()
    flat.indices[3, 1] = 10 ** 6
    with pytest.raises(api.TakeGpuError, match="vertex index"):
        api.host_build(flat)
    flat = scenes.cornell_box(8, 8, 1).flat()
    flat.prim_material[0] = 99
    with pytest.raises(api.TakeGpuError, match="material"):
        api.host_build(flat)
    flat = scenes.cornell_box(8, 8, 1).flat()
    flat.positions[5, 1] = np.inf          # would make the conservative padding of every box test infinite
    with pytest.raises(api.TakeGpuError, match="non-finite"):
        api.host_build(flat)
    flat = scenes.cornell_box(8, 8, 1).flat()
    flat.lights["kind"][0] = 7
    with pytest.raises(api.TakeGpuError, match="light kind"):
        api.host_build(flat)


def test_reference_order_tree_matches_oracle(oracle_lib, small_scene):
    name, _, flat = small_scene
    hb = api.host_build(flat)
    sc = oracle_lib.load(flat)
    box, links, root = sc.bvh()
    r = hb["ref_nodes"]
    assert hb["ref_root"] == root
    assert np.array_equal(np.concatenate([r["lo"], r["hi"]], axis=1), box)
    assert np.array_equal(np.stack([r["left"], r["right"], r["prim"]], axis=1), links)
    assert np.array_equal(hb["dfs_rank"], sc.dfs_rank())


def _walk_fast_tree(hb):
    """Returns per-node (first_slot, end_slot) of inner nodes and checks structural invariants."""
    nodes, n_prims = hb["fast_nodes"], len(hb["leaf_prims"])
    covered = np.zeros(n_prims, np.int32)
    visited = np.zeros(len(nodes), bool)
    stack = [0]
    while stack:
        i = stack.pop()
        assert not visited[i]
        visited[i] = True
        for k in (0, 1):
            c, cnt = int(nodes[i][f"child{k}"]), int(nodes[i][f"count{k}"])
            lo = [nodes[i][f"c{k}lo{a}"] for a in "xyz"]
            hi = [nodes[i][f"c{k}hi{a}"] for a in "xyz"]
            if c >= 0:
                assert cnt == 0
                stack.append(c)
            elif lo[0] <= hi[0]:           # a real leaf (empty children have inverted boxes)
                code = ~c
                first, count = code >> 3, (code & 7) + 1
                assert count == cnt and 1 <= count <= 8
                covered[first:first + count] += 1
    assert visited.all()
    assert (covered == 1).all()            # every leaf slot belongs to exactly one leaf
    assert sorted(hb["leaf_prims"].tolist()) == list(range(n_prims))


def test_fast_tree_structure(small_scene):
    _, _, flat = small_scene
    hb = api.host_build(flat)
    _walk_fast_tree(hb)
    # leaf records: slot 3 carries (rank << 32 | prim), e1 / e2 are exact differences of the FP64 vertices
    recs, prims = hb["leaf_records"], hb["leaf_prims"]
    bits = recs[:, 3].copy().view(np.int64)
    assert np.array_equal((bits & 0xffffffff).astype(np.int32), prims)
    assert np.array_equal((bits >> 32).astype(np.int32), hb["dfs_rank"][prims])
    tri = (flat.prim_flags[prims] & sceneio.PRIM_SPHERE) == 0
    idx = flat.indices[prims[tri]]
    v0, v1, v2 = (flat.positions[idx[:, k]] for k in range(3))
    assert np.array_equal(recs[tri, 0:3], v0)
    assert np.array_equal(recs[tri, 4:7], v1 - v0)
    assert np.array_equal(recs[tri, 8:11], v2 - v0)


def test_fast_boxes_contain_their_primitives(small_scene):
    """Every child box (FP32, rounded outward) must contain the FP64 bounds of all primitives below it."""
    _, _, flat = small_scene
    hb = api.host_build(flat)
    nodes, prims = hb["fast_nodes"], hb["leaf_prims"]
    sph = (flat.prim_flags & sceneio.PRIM_SPHERE) != 0
    lo = np.empty((flat.num_prims, 3)); hi = np.empty((flat.num_prims, 3))
    tri_idx = flat.indices[~sph]
    P = flat.positions[tri_idx]            # [n,3,3]
    lo[~sph], hi[~sph] = P.min(axis=1), P.max(axis=1)
    if sph.any():
        s = flat.spheres[flat.indices[sph, 0]]
        lo[sph], hi[sph] = s[:, :3] - s[:, 3:4], s[:, :3] + s[:, 3:4]

    def bounds(i):  # FP64 bounds of everything below inner node i, checking on the way down
        blo, bhi = np.full(3, np.inf), np.full(3, -np.inf)
        for k in (0, 1):
            c = int(nodes[i][f"child{k}"])
            clo = np.array([nodes[i][f"c{k}lo{a}"] for a in "xyz"], np.float64)
            chi = np.array([nodes[i][f"c{k}hi{a}"] for a in "xyz"], np.float64)
            if c >= 0:
                l, h = bounds(c)
            elif clo[0] <= chi[0]:
                code = ~c
                sl = prims[(code >> 3):(code >> 3) + (code & 7) + 1]
                l, h = lo[sl].min(axis=0), hi[sl].max(axis=0)
            else:
                continue
            assert (clo < l).all() and (chi > h).all()    # strictly outside: outward rounding + one ulp
            blo, bhi = np.minimum(blo, l), np.maximum(bhi, h)
        return blo, bhi

    import sys
    sys.setrecursionlimit(10000)
    bounds(0)


def test_degenerate_scenes_build():
    empty = scenes.SceneBuilder(8, 8, (0, 0, 5), (0, 0, 0)).flat()
    hb = api.host_build(empty)
    assert len(hb["fast_nodes"]) == 1 and len(hb["ref_nodes"]) == 0 and hb["ref_root"] == -1
    b = scenes.SceneBuilder(8, 8, (0, 0, 5), (0, 0, 0))
    m = b.material(sceneio.MAT_DIFFUSE, (0.5, 0.5, 0.5))
    b.mesh([(0, 0, 0), (1, 0, 0), (0, 1, 0)], [[0, 1, 2]], [(0, 0, 1)] * 3, None, m)
    hb = api.host_build(b.flat())
    assert len(hb["fast_nodes"]) == 1 and len(hb["ref_nodes"]) == 1
    _walk_fast_tree(hb)
    # many identical triangles: centroids coincide, the builder must still terminate with legal leaves
    b = scenes.SceneBuilder(8, 8, (0, 0, 5), (0, 0, 0))
    m = b.material(sceneio.MAT_DIFFUSE, (0.5, 0.5, 0.5))
    for _ in range(37):
        b.mesh([(0, 0, 0), (1, 0, 0), (0, 1, 0)], [[0, 1, 2]], [(0, 0, 1)] * 3, None, m)
    _walk_fast_tree(api.host_build(b.flat()))


def test_wide_tree_structure(small_scene):
    """The 4-wide collapse covers every leaf slot exactly once and every child box contains what is below it."""
    _, _, flat = small_scene
    hb = api.host_build(flat)
    wide, prims = hb["wide_nodes"], hb["leaf_prims"]
    n_prims = len(prims)
    sph = (flat.prim_flags & sceneio.PRIM_SPHERE) != 0
    lo = np.empty((n_prims, 3)); hi = np.empty((n_prims, 3))
    P = flat.positions[flat.indices[~sph]]
    lo[~sph], hi[~sph] = P.min(axis=1), P.max(axis=1)
    if sph.any():
        s = flat.spheres[flat.indices[sph, 0]]
        lo[sph], hi[sph] = s[:, :3] - s[:, 3:4], s[:, :3] + s[:, 3:4]
    covered = np.zeros(n_prims, np.int32)
    seen = np.zeros(len(wide), bool)

    def bounds(i):
        assert not seen[i]
        seen[i] = True
        blo, bhi = np.full(3, np.inf), np.full(3, -np.inf)
        for k in range(4):
            c = int(wide[i]["child"][k])
            if c == api.WIDE_EMPTY:
                continue
            clo = np.array([wide[i][a][k] for a in ("lox", "loy", "loz")], np.float64)
            chi = np.array([wide[i][a][k] for a in ("hix", "hiy", "hiz")], np.float64)
            if c >= 0:
                l, h = bounds(c)
            else:
                code = ~c
                first, count = code >> 3, (code & 7) + 1
                assert count == int(wide[i]["count"][k])
                covered[first:first + count] += 1
                sl = prims[first:first + count]
                l, h = lo[sl].min(axis=0), hi[sl].max(axis=0)
            assert (clo < l).all() and (chi > h).all()
            blo, bhi = np.minimum(blo, l), np.maximum(bhi, h)
        return blo, bhi

    import sys
    sys.setrecursionlimit(10000)
    if n_prims:
        bounds(0)
        assert seen.all() and (covered == 1).all()
        assert len(wide) <= max(1, len(hb["fast_nodes"]))   # collapsing never adds nodes


def test_large_tree_takes_the_parallel_paths():
    """180 k primitives: the chunk-parallel top-level partitions and the per-subtree parallel flatten of both tree images
    (bvh_build.cpp) only start at this size.  Every node is reached exactly once, every leaf slot is covered exactly once,
    children come after their parents, and every child box contains the FP64 bounds of what is below it."""
    flat = scenes.heightfield(n=300, width=64, height=32).flat()
    hb = api.host_build(flat)
    prims = hb["leaf_prims"]
    n = flat.num_prims
    assert n > (1 << 17) and sorted(prims.tolist()) == list(range(n))
    P = flat.positions[flat.indices]
    plo, phi = P.min(axis=1)[prims], P.max(axis=1)[prims]           # bounds by leaf slot

    def check(n_nodes, children_of):
        lo = np.full((n_nodes, 3), np.inf); hi = np.full((n_nodes, 3), -np.inf)
        refs = np.zeros(n_nodes, np.int32)
        covered = np.zeros(n, np.int32)
        for i in range(n_nodes - 1, -1, -1):                         # children have larger indices than their parent
            for c, clo, chi in children_of(i):
                if c >= 0:
                    assert c > i
                    refs[c] += 1
                    l, h = lo[c], hi[c]
                else:
                    code = ~c
                    first, count = code >> 3, (code & 7) + 1
                    covered[first:first + count] += 1
                    l, h = plo[first:first + count].min(axis=0), phi[first:first + count].max(axis=0)
                assert (clo < l).all() and (chi > h).all()
                lo[i] = np.minimum(lo[i], l); hi[i] = np.maximum(hi[i], h)
        assert (refs[1:] == 1).all() and refs[0] == 0 and (covered == 1).all()

    wide = hb["wide_nodes"]
    wl = np.stack([wide["lox"], wide["loy"], wide["loz"]], axis=2).astype(np.float64)   # [node, child, axis]
    wh = np.stack([wide["hix"], wide["hiy"], wide["hiz"]], axis=2).astype(np.float64)
    wc = wide["child"]
    check(len(wide), lambda i: [(int(wc[i, k]), wl[i, k], wh[i, k]) for k in range(4) if wc[i, k] != api.WIDE_EMPTY])
    fast = hb["fast_nodes"]
    fl = [np.stack([fast[f"c{k}lo{a}"] for a in "xyz"], axis=1).astype(np.float64) for k in (0, 1)]
    fh = [np.stack([fast[f"c{k}hi{a}"] for a in "xyz"], axis=1).astype(np.float64) for k in (0, 1)]
    fc = [fast["child0"], fast["child1"]]
    check(len(fast), lambda i: [(int(fc[k][i]), fl[k][i], fh[k][i]) for k in (0, 1) if fl[k][i][0] <= fh[k][i][0]])


def test_host_build_save_load_roundtrip(tmp_path, small_scene):
    """Build once, create many: a saved build loads back bit for bit, a truncated or foreign file is refused, and a build is
    refused for a scene it was not made from (take_gpu_scene_create_prebuilt checks the geometry hash BEFORE touching CUDA)."""
    import ctypes as C
    name, _, flat = small_scene
    hb = api.HostBuild(flat)
    path = str(tmp_path / "build.bin")
    hb.save(path)
    assert not os.path.exists(path + ".part")
    a, b = hb.arrays(), api.HostBuild(path=path).arrays()
    for k in ("ref_nodes", "dfs_rank", "fast_nodes", "wide_nodes", "leaf_prims", "leaf_records"):
        assert np.array_equal(a[k], b[k]), k
    assert (a["ref_root"], a["depth"], a["abs_max"]) == (b["ref_root"], b["depth"], b["abs_max"])
    raw = open(path, "rb").read()
    flipped = bytearray(raw)
    flipped[len(raw) // 2] ^= 0x40                                   # one bit inside an array: the payload hash notices
    # header scalars the device code trusts (ref_root at byte 72, abs_max at 96, the counts from 24): covered by the header hash,
    # range-checked, and a count that does not fit the file length is refused before anything is allocated from it
    hdr_cases = []
    for off, val in ((72, b"\x07"), (96 + 7, b"\x7f"), (96 + 6, b"\xf8"), (40, b"\xff\xff\xff\x7f"), (76, b"\x63")):
        c = bytearray(raw)
        c[off:off + len(val)] = val
        if bytes(c) != raw:
            hdr_cases.append(bytes(c))
    assert len(hdr_cases) >= 4
    for bad in [raw[:len(raw) // 2], b"TAKEHB01" + raw[8:], raw + b"x", b"", bytes(flipped)] + hdr_cases:
        p2 = str(tmp_path / "bad.bin")
        open(p2, "wb").write(bad)
        with pytest.raises(api.TakeGpuError):
            api.HostBuild(path=p2)
    with pytest.raises(api.TakeGpuError):
        api.HostBuild(path=str(tmp_path / "missing.bin"))
    # a build of a different scene: refused with TAKE_E_INVALID (-1), whether or not a GPU is present
    other = next(o for o in (scenes.cornell_box(16, 16, 1).flat(), scenes.multi_light(16, 16, 1, n_side=3).flat())
                 if o.num_prims != flat.num_prims)
    lib = api.load_library()
    desc = other.to_desc()
    h = C.c_void_p()
    rc = lib.take_gpu_scene_create_prebuilt(0, C.byref(desc), hb.h, C.byref(h))
    assert rc == -1 and b"do not belong" in lib.take_gpu_last_error()
    # same counts, one vertex moved: the hash notices
    import copy
    moved = copy.copy(flat)
    moved.positions = flat.positions.copy()
    moved.positions[flat.indices[0, 0] if not (flat.prim_flags[0] & sceneio.PRIM_SPHERE) else 0] += 0.5
    desc2 = moved.to_desc()
    if flat.positions.size:
        rc = lib.take_gpu_scene_create_prebuilt(0, C.byref(desc2), hb.h, C.byref(h))
        assert rc == -1 and b"do not belong" in lib.take_gpu_last_error()
    hb.close()


@pytest.mark.parametrize("pattern,distinct", [(0, 0), (0, 1000), (0, 7), (1, 0), (2, 0), (3, 0), (4, 0), (4, 50), (1, 16)])
def test_forked_sort_is_std_sort(pattern, distinct):
    """The reference-order tree depends on the order libstdc++'s std::sort leaves EQUAL keys in (src/bvh.cpp:25-31), so the
    multi-threaded sort used near the root must return std::sort's exact permutation -- on random keys, heavy ties,
    presorted / reversed / constant / organ-pipe inputs."""
    lib = api.load_library()
    lib.take_gpu_selftest_sort.restype = C.c_int64
    lib.take_gpu_selftest_sort.argtypes = [C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_uint64]
    for n in (70_000, 300_001):
        assert lib.take_gpu_selftest_sort(n, distinct, pattern, 4, 12345 + n) == 0
