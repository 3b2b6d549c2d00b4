"""Parser throughput for big meshes: a PLY of N triangles WITHOUT normals (so that the angle-weighted vertex
normals of compute_normals are part of the job) -> flat scene arrays, through take_gpu_builder_add_ply and -- with --ref --
through the unmodified reference parser (parse_ply + compute_normals + one Shape per triangle), arrays compared bit for bit.
Host-only.  Usage: python tools/mesh_load_bench.py [quads_per_side=2236] [--ref]"""
import json, os, sys, tempfile, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from take_b200 import api, scenes
from take_b200.sceneio import FlatScene

n = int(([a for a in sys.argv[1:] if not a.startswith("-")] or ["2236"])[0])
rng = np.random.default_rng(1)
noise = rng.normal(0.0, 1.0, size=(n + 1, n + 1))
P, T, N, UV = scenes._grid_mesh(n, 100.0, lambda X, Z: 15.0 * np.sin(0.06 * X) * np.cos(0.05 * Z) + noise)
with tempfile.TemporaryDirectory() as d:
    path = os.path.join(d, "mesh0.ply")
    nv, nf = len(P), len(T)
    hdr = ["ply", "format binary_little_endian 1.0", f"element vertex {nv}", "property float x", "property float y", "property float z",
           f"element face {nf}", "property list uchar int vertex_indices", "end_header"]
    faces = np.zeros(nf, dtype=np.dtype([("n", "u1"), ("i", "<i4", 3)]))
    faces["n"] = 3
    faces["i"] = T
    with open(path, "wb") as f:
        f.write(("\n".join(hdr) + "\n").encode())
        P.astype("<f4").tofile(f)
        faces.tofile(f)
    out = {"triangles": nf, "vertices": nv, "file_MB": round(os.path.getsize(path) / 1e6, 1), "host_threads": os.cpu_count()}
    best = None
    for rep in range(3):
        b = api.DescBuilder()
        t0 = time.perf_counter()
        b.add_ply(path, 0)
        desc = b.desc()
        dt = time.perf_counter() - t0
        if best is None or dt < best[0]:
            best = (dt, b.timings())
        if rep < 2:
            b.close()
    out["builder_s"] = round(best[0], 3)
    out["builder_ms"] = {k: round(v, 1) for k, v in best[1].items()}
    if "--ref" in sys.argv:
        from oracle import bindings as ob
        open(os.path.join(d, "scene.xml"), "w").write(
            '<?xml version="1.0"?><scene version="0.5.0"><sensor type="perspective"><film type="hdrfilm"><integer name="width" value="8"/>'
            '<integer name="height" value="8"/></film></sensor><bsdf type="diffuse" id="m0"><rgb name="reflectance" value="0.5,0.5,0.5"/></bsdf>'
            '<shape type="ply"><string name="filename" value="mesh0.ply"/><ref id="m0"/></shape></scene>')
        cwd = os.getcwd()
        t0 = time.perf_counter()
        rs = ob.RefLib().load(os.path.join(d, "scene.xml"))      # parse_scene + build_bvh (the harness builds the tree on load)
        out["reference_parse_and_bvh_s"] = round(time.perf_counter() - t0, 2)
        os.chdir(cwd)
        rs.dump(os.path.join(d, "ref.takescene"))
        ref = FlatScene.load(os.path.join(d, "ref.takescene"))
        got = b.arrays()
        out["identical"] = all(got[k].tobytes() == getattr(ref, k).tobytes() for k in
                               ("positions", "normals", "uvs", "indices", "prim_material", "prim_light", "prim_flags"))
    b.close()
print(json.dumps(out))
