"""The C++ adapter's scene loading (take_b200/host/render_gpu.cpp: load_scene_fast): the <scene> element
walked with the reference's own parser functions, mesh payloads routed through take_gpu_builder_* -- against the same adapter
flattening the Scene of the reference's parse_scene() (-ref_parse).  The two TAKESCN1 dumps must be byte-identical.
Needs oracle/_ref/take_gpu (the adapter linked with the reference's unmodified front end); no GPU."""
import os
import subprocess

import numpy as np
import pytest

from take_b200 import scenes
from take_b200.sceneio import FlatScene

import test_mesh_load as tm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ADAPTER = os.path.join(ROOT, "oracle", "_ref", "take_gpu")
pytestmark = pytest.mark.skipif(not os.path.exists(ADAPTER), reason="oracle/_ref/take_gpu not built (needs /root/reference)")


def dumps(xml, tmp_path):
    out = []
    for tag, extra in (("fast", []), ("ref", ["-ref_parse"])):
        p = str(tmp_path / f"{tag}.takescene")
        r = subprocess.run([ADAPTER, xml, *extra, "-dump_scene", p], capture_output=True, text=True, cwd=str(tmp_path), timeout=300)
        assert r.returncode == 0, r.stdout + r.stderr
        out.append(open(p, "rb").read())
    return out


def test_fast_load_equals_reference_parse(small_scene, tmp_path):
    name, builder, flat = small_scene
    xml = builder.write(str(tmp_path / "sc"))
    fast, ref = dumps(xml, tmp_path)
    assert fast == ref
    assert FlatScene.load(str(tmp_path / "fast.takescene")).same_as(flat) == []


def test_fast_load_mixed_shapes_and_transforms(tmp_path):
    """Rectangles, a sphere, a nested <bsdf>, <default> substitution, a scene-level point emitter, PLY with and without
    normals under a general (rotation + non-uniform scale + translation) toWorld -- the inverse is the reference's own."""
    d = tmp_path / "sc"
    d.mkdir()
    P, T = tm.bumpy_mesh(3, 7)
    N = (P / np.linalg.norm(P, axis=1, keepdims=True)).astype(np.float32).astype(np.float64)
    UV = np.random.default_rng(2).uniform(size=(len(P), 2)).astype(np.float32).astype(np.float64)
    tm.write_ply_variant(str(d / "a.ply"), P, T, N=N, UV=UV)
    tm.write_ply_variant(str(d / "b.ply"), P, T, fmt="binary_big_endian", vtype="double", itype=("uchar", "uint"))
    xml = """<?xml version="1.0" encoding="utf-8"?>
<scene version="0.5.0">
<default name="spp" value="3"/>
<sensor type="perspective"><float name="fov" value="40"/><transform name="toWorld"><lookat origin="0, 1, 6" target="0, 1, 0" up="0, 1, 0"/></transform>
<sampler type="independent"><integer name="sampleCount" value="$spp"/></sampler><film type="hdrfilm"><integer name="width" value="12"/><integer name="height" value="10"/></film></sensor>
<background><rgb name="radiance" value="0.1, 0.2, 0.3"/></background>
<bsdf type="diffuse" id="white"><rgb name="reflectance" value="0.7, 0.7, 0.7"/></bsdf>
<bsdf type="blinn_microfacet" id="gloss"><rgb name="reflectance" value="0.6, 0.5, 0.4"/><float name="exponent" value="30"/></bsdf>
<emitter type="point"><point name="position" x="1" y="2" z="3"/><rgb name="intensity" value="4, 5, 6"/></emitter>
<shape type="rectangle"><transform name="toWorld"><scale x="3" y="3"/><rotate x="1" angle="-90"/></transform><ref id="white"/></shape>
<shape type="ply"><string name="filename" value="a.ply"/><transform name="toWorld"><scale x="0.5" y="0.8" z="0.6"/><rotate y="1" angle="33"/><rotate x="1" angle="-12"/><translate x="-1" y="1" z="0.25"/></transform><ref id="gloss"/></shape>
<shape type="sphere"><point name="center" x="1.2" y="0.5" z="0"/><float name="radius" value="0.5"/><bsdf type="mirror"><rgb name="reflectance" value="0.9, 0.9, 0.9"/></bsdf>
  <emitter type="area"><rgb name="radiance" value="3, 3, 3"/></emitter></shape>
<shape type="ply"><string name="filename" value="b.ply"/><transform name="toWorld"><rotate z="1" angle="70"/><translate x="1" y="2" z="-1"/></transform><ref id="white"/>
  <emitter type="area"><rgb name="radiance" value="1, 2, 1"/></emitter></shape>
<shape type="rectangle"><transform name="toWorld"><scale x="0.4" y="0.4"/><rotate x="1" angle="90"/><translate y="2.9"/></transform><ref id="white"/>
  <emitter type="area"><rgb name="radiance" value="9, 9, 9"/></emitter></shape>
<shape type="ply"><string name="filename" value="a.ply"/><boolean name="faceNormals" value="true"/><ref id="gloss"/></shape>
</scene>
"""
    (d / "scene.xml").write_text(xml)
    fast, ref = dumps(str(d / "scene.xml"), tmp_path)
    assert fast == ref
    f = FlatScene.load(str(tmp_path / "fast.takescene"))
    assert f.num_prims == 2 + len(T) + 1 + len(T) + 2 + len(T) and len(f.lights) == 1 + 1 + len(T) + 2 and f.spp == 3
    assert f.lights["kind"][0] == 0 and len(f.materials) == 3
