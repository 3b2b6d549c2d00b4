"""take_gpu_builder_* (SURVEY.md 8f-2): meshes straight into the flat scene arrays on all host threads, bit-identical to
flattening the Scene the UNMODIFIED reference parser builds (parse_ply + compute_normals + per-face Shape / light
expansion: src/parse/parse_ply.cpp:16-120, src/compute_normals.cpp:12-47, src/parse/parse_scene.cpp:934-945).
Host-only: no GPU needed.  The live comparisons need oracle/_ref (skipped without it); the golden case does not."""
import os
import tempfile
import time

import numpy as np
import pytest

from take_b200 import api, scenes, sceneio
from take_b200.sceneio import FlatScene

KEYS = ("positions", "normals", "uvs", "indices", "prim_material", "prim_light", "prim_flags", "spheres")
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_like_the_parser(builder: scenes.SceneBuilder, directory: str) -> dict:
    """What a host does with the builder API for a scene written by SceneBuilder.write(): shapes in file order."""
    b = api.DescBuilder()
    for pos, inten in builder.point_lights:
        b.add_point_light(pos, inten)     # (the parser's swapped PointLight fields, see SceneBuilder.flat)
    for i, m in enumerate(builder.meshes):
        if m.get("sphere"):
            b.add_sphere(m["center"], m["radius"], m["material"], m["radiance"])
        else:
            b.add_ply(os.path.join(directory, f"mesh{i}.ply"), m["material"], radiance=m["radiance"])
    out = b.arrays()
    b.close()
    return out


def assert_same(got: dict, flat: FlatScene):
    for k in KEYS:
        want = getattr(flat, k)
        assert got[k].shape == want.shape and got[k].tobytes() == want.tobytes(), k      # bit for bit, -0.0 vs 0.0 included
    assert got["lights"].tobytes() == flat.lights.tobytes()


def test_builder_matches_reference_parser(ref_lib, small_scene):
    name, builder, flat = small_scene
    with tempfile.TemporaryDirectory() as d:
        rs = ref_lib.load(builder.write(d))
        rs.dump(d + "/ref.takescene")
        ref = FlatScene.load(d + "/ref.takescene")
        rs.close()
        assert_same(load_like_the_parser(builder, d), ref)


def write_ply_variant(path, P, T, N=None, UV=None, fmt="binary_little_endian", vtype="float", itype=("uchar", "int"), extra=False):
    """PLY writer covering what parse_ply accepts: ascii / little / big endian, float or double vertices, every index width,
    optional extra properties and an extra element that must be skipped."""
    nv, nf = len(P), len(T)
    big = fmt == "binary_big_endian"
    e = ">" if big else "<"
    vt = {"float": "f4", "double": "f8"}[vtype]
    it = {"char": "i1", "uchar": "u1", "short": "i2", "ushort": "u2", "int": "i4", "uint": "u4"}
    props = ["x", "y", "z"] + (["nx", "ny", "nz"] if N is not None else []) + (["u", "v"] if UV is not None else [])
    hdr = ["ply", f"format {fmt} 1.0", "comment made by tests/test_mesh_load.py"]
    if extra:
        hdr += ["element junk 2", "property uchar a", "property list uchar short b"]
    hdr += [f"element vertex {nv}"] + [f"property {vtype} {p}" for p in props] + (["property uchar red"] if extra else [])
    hdr += [f"element face {nf}", f"property list {itype[0]} {itype[1]} vertex_indices", "end_header"]
    cols = [P] + ([N] if N is not None else []) + ([UV] if UV is not None else [])
    V = np.concatenate(cols, axis=1)
    with open(path, "wb") as f:
        f.write(("\n".join(hdr) + "\n").encode())
        if fmt == "ascii":
            if extra:
                f.write(b"7 2 5 6\n9 0\n")
            for row in V:
                f.write((" ".join(repr(float(np.float32(x))) if vtype == "float" else repr(float(x)) for x in row) + (" 200" if extra else "") + "\n").encode())
            for t in T:
                f.write(f"3 {t[0]} {t[1]} {t[2]}\n".encode())
        else:
            if extra:
                f.write(np.array([7, 2], "u1").tobytes() + np.array([5, 6], e + "i2").tobytes() + np.array([9, 0], "u1").tobytes())
            dt = [("v", e + vt, V.shape[1])] + ([("red", "u1")] if extra else [])
            rec = np.zeros(nv, np.dtype(dt))
            rec["v"] = V
            if extra:
                rec["red"] = 200
            f.write(rec.tobytes())
            fr = np.zeros(nf, np.dtype([("n", e + it[itype[0]]), ("i", e + it[itype[1]], 3)]))
            fr["n"] = 3
            fr["i"] = T
            f.write(fr.tobytes())


def xml_for(shapes, spp=1):
    """Minimal scene: one diffuse bsdf, `shapes` = list of (file, matrix16 or None, face_normals, radiance or None)."""
    x = ['<?xml version="1.0" encoding="utf-8"?>', '<scene version="0.5.0">',
         '<sensor type="perspective"><float name="fov" value="45"/><transform name="toWorld"><lookat origin="0, 0, 5" target="0, 0, 0" up="0, 1, 0"/>'
         f'</transform><sampler type="independent"><integer name="sampleCount" value="{spp}"/></sampler>'
         '<film type="hdrfilm"><integer name="width" value="8"/><integer name="height" value="8"/></film></sensor>',
         '<bsdf type="diffuse" id="m0"><rgb name="reflectance" value="0.5, 0.5, 0.5"/></bsdf>',
         '<bsdf type="diffuse" id="m1"><rgb name="reflectance" value="0.2, 0.5, 0.7"/></bsdf>']
    for i, (fn, M, face_normals, rad) in enumerate(shapes):
        x.append(f'<shape type="ply"><string name="filename" value="{fn}"/><ref id="m{i % 2}"/>')
        if M is not None:
            x.append('<transform name="toWorld"><matrix value="' + " ".join("%.9g" % float(np.float32(v)) for v in np.ravel(M)) + '"/></transform>')
        if face_normals:
            x.append('<boolean name="faceNormals" value="true"/>')
        if rad is not None:
            x.append('<emitter type="area"><rgb name="radiance" value="%g, %g, %g"/></emitter>' % tuple(rad))
        x.append('</shape>')
    x.append('</scene>')
    return "\n".join(x) + "\n"


def bumpy_mesh(subdiv, seed):
    sv, sf = scenes._icosphere(subdiv)
    rng = np.random.default_rng(seed)
    P = (sv * (1 + 0.2 * rng.normal(size=(len(sv), 1)))).astype(np.float32).astype(np.float64)
    return P, sf


def test_ply_variants_transforms_and_computed_normals(ref_lib):
    """Files WITHOUT normals (the angle-weighted vertex normals of compute_normals, accumulated in face order), every PLY
    encoding parse_ply accepts, arbitrary toWorld matrices on normal-free meshes, exactly invertible ones on meshes with
    normals (the inverse is the host's: the test passes numpy's, which is exact for power-of-two scales and integer
    translations), faceNormals, emitters, degenerate faces, an unreferenced vertex."""
    rng = np.random.default_rng(5)
    with tempfile.TemporaryDirectory() as d:
        shapes, calls = [], []
        P, T = bumpy_mesh(3, 1)
        T = np.concatenate([T, [[0, 0, 5], [3, 7, 3]]]).astype(np.int32)            # zero-area faces: skipped by compute_normals
        P = np.concatenate([P, [[9.0, 9.0, 9.0]]])                                 # a vertex no face uses: normal (0,0,0)
        rot = np.array([[0.36, 0.48, -0.8, 1.5], [-0.8, 0.6, 0.0, -2.25], [0.48, 0.64, 0.6, 0.125], [0, 0, 0, 1]])
        rot = rot.astype(np.float32).astype(np.float64)
        pow2 = np.array([[2, 0, 0, 3], [0, 0.5, 0, -1], [0, 0, 4, 2], [0, 0, 0, 1.0]])
        N = (P / np.maximum(np.linalg.norm(P, axis=1, keepdims=True), 1e-9)).astype(np.float32).astype(np.float64)
        UV = rng.uniform(size=(len(P), 2)).astype(np.float32).astype(np.float64)
        cases = [
            ("a.ply", dict(), None, False, None),
            ("b.ply", dict(fmt="binary_big_endian", vtype="double", itype=("uchar", "uint")), rot, False, (3.0, 2.0, 1.0)),
            ("c.ply", dict(extra=True, itype=("char", "uint")), rot, False, None),
            ("d.ply", dict(N=N, UV=UV, itype=("int", "ushort"), extra=True), pow2, False, None),
            ("e.ply", dict(N=N, fmt="binary_big_endian", vtype="double"), None, True, (1.0, 1.0, 1.0)),
            ("f.ply", dict(UV=UV, itype=("ushort", "short")), None, False, None),
        ]
        for fn, kw, M, face_normals, rad in cases:
            write_ply_variant(os.path.join(d, fn), P, T, **kw)
            shapes.append((fn, M, face_normals, rad))
        open(os.path.join(d, "scene.xml"), "w").write(xml_for(shapes))
        rs = ref_lib.load(os.path.join(d, "scene.xml"))
        rs.dump(d + "/ref.takescene")
        ref = FlatScene.load(d + "/ref.takescene")
        rs.close()
        b = api.DescBuilder()
        for i, (fn, M, face_normals, rad) in enumerate(shapes):
            b.add_ply(os.path.join(d, fn), i % 2, to_world=M, inv_to_world=None if M is None else np.linalg.inv(M),
                      face_normals=face_normals, radiance=rad)
        got = b.arrays()
        b.close()
        assert_same(got, ref)
        assert (ref.prim_light >= 0).sum() == 2 * len(T) and len(ref.lights) == 2 * len(T)


def test_ascii_ply_equals_binary(tmp_path):
    """ascii PLY is accepted as well.  It cannot be pinned: the reference's vendored tinyply throws "unexpected EOF" on any
    ascii file with more than three faces (the offset of its list-size read is never reset, 3rdparty/tinyply.h:891), so the
    check is that the ascii and the binary encoding of the same float data give identical arrays."""
    P, T = bumpy_mesh(2, 4)
    rng = np.random.default_rng(1)
    N = (P / np.linalg.norm(P, axis=1, keepdims=True)).astype(np.float32).astype(np.float64)
    UV = rng.uniform(size=(len(P), 2)).astype(np.float32).astype(np.float64)
    M = np.array([[2, 0, 0, 3], [0, 0.5, 0, -1], [0, 0, 4, 2], [0, 0, 0, 1.0]])
    out = []
    for fmt, vtype, extra in (("ascii", "float", True), ("binary_little_endian", "float", False), ("ascii", "double", False),
                              ("binary_big_endian", "double", True)):
        p = str(tmp_path / f"{fmt}_{vtype}.ply")
        write_ply_variant(p, P, T, N=N, UV=UV, fmt=fmt, vtype=vtype, extra=extra)
        b = api.DescBuilder()
        b.add_ply(p, 0, to_world=M, inv_to_world=np.linalg.inv(M), radiance=(1, 2, 3))
        out.append(b.arrays())
        b.close()
    for o in out[1:]:
        for k in KEYS + ("lights",):
            assert o[k].tobytes() == out[0][k].tobytes(), k


def test_builder_refuses_bad_files(tmp_path):
    P, T = bumpy_mesh(1, 2)
    b = api.DescBuilder()
    with pytest.raises(api.TakeGpuError, match="cannot open"):
        b.add_ply(str(tmp_path / "missing.ply"), 0)
    p = str(tmp_path / "quad.ply")
    write_ply_variant(p, P, T)
    raw = bytearray(open(p, "rb").read())
    off = raw.index(b"end_header\n") + len(b"end_header\n") + len(P) * 12
    raw[off] = 4                                                   # first face claims four vertices
    open(p, "wb").write(raw)
    with pytest.raises(api.TakeGpuError, match="triangles"):
        b.add_ply(p, 0)
    p = str(tmp_path / "oob.ply")
    T2 = T.copy(); T2[3, 1] = len(P) + 5
    write_ply_variant(p, P, T2)
    with pytest.raises(api.TakeGpuError, match="out of range"):
        b.add_ply(p, 0)
    p = str(tmp_path / "short.ply")
    write_ply_variant(p, P, T)
    data = open(p, "rb").read()
    open(p, "wb").write(data[:-7])
    with pytest.raises(api.TakeGpuError, match="truncated"):
        b.add_ply(p, 0)
    p = str(tmp_path / "nopos.ply")
    open(p, "wb").write(b"ply\nformat ascii 1.0\nelement vertex 1\nproperty float x\nproperty float y\nelement face 0\n"
                        b"property list uchar int vertex_indices\nend_header\n0 0\n")
    with pytest.raises(api.TakeGpuError, match="positions not found"):
        b.add_ply(p, 0)
    assert b.arrays()["positions"].shape == (0, 3)                 # nothing was appended by the failed calls
    b.close()


def test_corrupted_files_are_refused_not_crashed_on(tmp_path):
    """Nothing may propagate through the C boundary: a header that promises 10^11 faces (a std::length_error inside the
    loader before the fix), truncations, flipped bytes, absurd counts -- every case is either refused with an error code or
    accepted with all indices in range, in all three formats."""
    rng = np.random.default_rng(5)
    P, T = bumpy_mesh(2, 2)
    for fmt in ("binary_little_endian", "binary_big_endian", "ascii"):
        good = str(tmp_path / f"good_{fmt}.ply")
        write_ply_variant(good, P, T, fmt=fmt)
        data = open(good, "rb").read()
        cases = [data.replace(b"element face %d" % len(T), b"element face 100000000000"),
                 data.replace(b"element vertex %d" % len(P), b"element vertex 4000000000"),
                 data.replace(b"element face %d" % len(T), b"element face -3")]
        for _ in range(40):
            b_ = bytearray(data)
            if rng.random() < 0.5:
                b_ = b_[:int(rng.integers(0, len(b_)))]
            else:
                for _ in range(int(rng.integers(1, 8))):
                    b_[int(rng.integers(0, len(b_)))] = int(rng.integers(0, 256))
            cases.append(bytes(b_))
        for i, c in enumerate(cases):
            p = str(tmp_path / f"case_{fmt}_{i}.ply")
            open(p, "wb").write(c)
            b = api.DescBuilder()
            try:
                b.add_ply(p, 0)
                a = b.arrays()
                assert a["indices"].size == 0 or (a["indices"].min() >= 0 and a["indices"].max() < len(a["positions"]))
            except api.TakeGpuError:
                pass
            finally:
                b.close()


def test_large_mesh_parallel_paths_equal_small_path(tmp_path):
    """A mesh big enough for every loop to run in chunks on several threads gives the same arrays as the add_mesh route fed
    with numpy data (same code, single chunk sizes differ) -- and computed normals do not depend on the thread count."""
    P, T = bumpy_mesh(6, 3)                                         # 40 962 vertices, 81 920 faces
    p = str(tmp_path / "big.ply")
    write_ply_variant(p, P, T)
    b = api.DescBuilder()
    t0 = time.perf_counter()
    b.add_ply(p, 0)
    dt = time.perf_counter() - t0
    a = b.arrays()
    b.close()
    # brute-force restatement of compute_normals in numpy, face order kept by np.add.at being sequential per index
    nrm = np.zeros_like(P)
    for f in T:                                                     # (python loop: 82 k faces, a second or two)
        v = P[f]
        n = np.cross(v[1] - v[0], v[2] - v[0])
        l = np.sqrt((n[0] * n[0] + n[1] * n[1]) + n[2] * n[2])
        if l == 0:
            continue
        n = n * (1.0 / l)
        for c in range(3):
            s1, s2 = v[(c + 1) % 3] - v[c], v[(c + 2) % 3] - v[c]
            u1 = s1 * (1.0 / np.sqrt((s1[0] * s1[0] + s1[1] * s1[1]) + s1[2] * s1[2]))
            u2 = s2 * (1.0 / np.sqrt((s2[0] * s2[0] + s2[1] * s2[1]) + s2[2] * s2[2]))
            dd = (u1[0] * u2[0] + u1[1] * u2[1]) + u1[2] * u2[2]
            w = u2 + u1 if dd < 0 else u2 - u1
            h = 0.5 * np.sqrt((w[0] * w[0] + w[1] * w[1]) + w[2] * w[2])
            ang = (np.pi - 2) * np.arcsin(h) if dd < 0 else 2 * np.arcsin(h)
            nrm[f[c]] = nrm[f[c]] + n * ang
    l = np.sqrt((nrm[:, 0] * nrm[:, 0] + nrm[:, 1] * nrm[:, 1]) + nrm[:, 2] * nrm[:, 2])
    nrm = nrm * (1.0 / l)[:, None]
    assert np.abs(a["normals"] - nrm).max() < 1e-14                 # numpy's arcsin may differ from glibc's asin in the last bit
    assert a["indices"].shape == (len(T), 3) and (a["prim_flags"] == sceneio.PRIM_HAS_NORMALS).all()
    print(f"81 920-face PLY -> arrays in {dt * 1e3:.1f} ms")
