"""Procedural scenes for the BASELINE.json configs.

Every generator produces the SAME scene twice:
  * `SceneBuilder.flat()`    -> `FlatScene` (numpy arrays) for the GPU core's C-ABI, and
  * `SceneBuilder.write(dir)` -> Mitsuba-style `scene.xml` + binary PLY meshes that the reference's own
    front end parses (src/parse/parse_scene.cpp, src/parse/parse_ply.cpp).
The two are bit-identical by construction (tests/test_scenes.py checks it against the reference
parser): all XML scalars are float32 values because the reference reads them with std::stof
(parse_scene.cpp:52-58,98-104); mesh data are float32 in world space under an identity toWorld, so
xform_point (src/transform.cpp:79-87) is exact; normals go through xform_normal = normalize() in
double (transform.cpp:95-100, src/vector.h:249-257), restated below in numpy with the same op order.
"""
from __future__ import annotations

import os

import numpy as np

from . import sceneio as sio
from .sceneio import FlatScene

_BSDF_XML = {
    sio.MAT_DIFFUSE: ("diffuse", "reflectance"), sio.MAT_MIRROR: ("mirror", "reflectance"),
    sio.MAT_PLASTIC: ("plastic", "reflectance"), sio.MAT_PHONG: ("phong", "reflectance"),
    sio.MAT_BLINN_PHONG: ("blinn", "reflectance"), sio.MAT_BLINN_MICROFACET: ("blinn_microfacet", "reflectance"),
    sio.MAT_DISNEY_DIFFUSE: ("disneydiffuse", "baseColor"), sio.MAT_DISNEY_METAL: ("disneymetal", "baseColor"),
    sio.MAT_DISNEY_GLASS: ("disneyglass", "baseColor"), sio.MAT_DISNEY_SHEEN: ("disneysheen", "baseColor"),
    sio.MAT_DISNEY_BSDF: ("disneybsdf", "baseColor"),
    sio.MAT_DISNEY_CLEARCOAT: ("disneyclearcoat", None),   # no colour at all (parse_scene.cpp:613-621)
}


def f32(x):
    """Round to float32 and widen back: what std::stof / a float PLY property hands the reference."""
    return np.asarray(x, dtype=np.float32).astype(np.float64)


def _fmt(x) -> str:
    return "%.9g" % float(np.float32(x))


def _vec(v) -> str:
    return ", ".join(_fmt(c) for c in v)


def ref_normalize(v: np.ndarray) -> np.ndarray:
    """normalize() of src/vector.h:249-257: l = sqrt((x*x + y*y) + z*z); l <= 0 -> 0; else v * (1/l)."""
    v = np.asarray(v, dtype=np.float64)
    l = np.sqrt((v[:, 0] * v[:, 0] + v[:, 1] * v[:, 1]) + v[:, 2] * v[:, 2])
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = 1.0 / l
        out = v * inv[:, None]
    out[l <= 0] = 0.0
    return out


class SceneBuilder:
    def __init__(self, width, height, lookfrom, lookat, up=(0, 1, 0), vfov=45.0, spp=16, background=(0, 0, 0)):
        self.width, self.height, self.spp = int(width), int(height), int(spp)
        self.lookfrom, self.lookat, self.up = f32(lookfrom), f32(lookat), f32(up)
        self.vfov = float(f32(vfov))
        self.background = f32(background)
        self.materials = []   # (type, color|None, tex|None, params dict)
        self.textures = []    # (name, rgbe uint8 [h,w,4])
        self.meshes = []      # shapes in the order the reference would parse them: mesh dicts and sphere dicts
        self.env = None       # optional environment map (extension: the reference has none)
        self.env_sample = False
        self.point_lights = []  # top-level <emitter type="point">: (position, intensity), parsed before any shape

    # -- materials ----------------------------------------------------------------------------
    def material(self, mtype, color=(0.5, 0.5, 0.5), texture=None, uvscale=(1, 1), uvoffset=(0, 0), **params) -> int:
        """params: eta (plastic), exponent (phong/blinn/blinn_microfacet), roughness/subsurface (disneydiffuse)."""
        self.materials.append(dict(type=mtype, color=f32(color), texture=texture, uvscale=f32(uvscale),
                                   uvoffset=f32(uvoffset), params={k: float(f32(v)) for k, v in params.items()}))
        return len(self.materials) - 1

    def texture_rgbe(self, rgbe: np.ndarray) -> int:
        """Image texture given as Radiance RGBE bytes [h,w,4]; texel = mantissa * 2^(e-136) exactly
        (stb_image's stbi__hdr_convert, which the reference's imread3 calls for .hdr: src/image.cpp:92-104)."""
        rgbe = np.ascontiguousarray(rgbe, dtype=np.uint8)
        assert rgbe.ndim == 3 and rgbe.shape[2] == 4
        self.textures.append(rgbe)
        return len(self.textures) - 1

    def environment(self, rgb, sample=True):
        """Lat-long environment map [h,w,3] (row 0 = up).  EXTENSION: the reference's only environment is the constant
        <background>; `write()` therefore refuses such a scene unless the map is constant."""
        self.env = np.ascontiguousarray(rgb, dtype=np.float64)
        self.env_sample = bool(sample)

    def point_light(self, position, intensity=(1, 1, 1)):
        """Top-level <emitter type="point"> (src/parse/parse_scene.cpp:701-727).  The integrators ignore point lights
        (path_tracing.h:33: get_if<DiffuseAreaLight> fails) but they take part in the uniform light pick and dilute it
        (src/light.cpp:5-7)."""
        self.point_lights.append((f32(position), f32(intensity)))
        return len(self.point_lights) - 1

    # -- geometry -----------------------------------------------------------------------------
    def mesh(self, positions, indices, normals, uvs=None, material=0, radiance=None):
        """World-space triangle mesh.  `normals` are required (the reference would otherwise run
        compute_normals, and emitters need them: src/shape.cpp:163-165)."""
        self.meshes.append(dict(
            positions=np.asarray(positions, dtype=np.float32).reshape(-1, 3),
            indices=np.asarray(indices, dtype=np.int32).reshape(-1, 3),
            normals=np.asarray(normals, dtype=np.float32).reshape(-1, 3),
            uvs=None if uvs is None else np.asarray(uvs, dtype=np.float32).reshape(-1, 2),
            material=int(material), radiance=None if radiance is None else f32(radiance)))
        return len(self.meshes) - 1

    def sphere(self, center, radius, material=0, radiance=None):
        """<shape type="sphere"> (src/parse/parse_scene.cpp:785-808)."""
        self.meshes.append(dict(sphere=True, center=f32(center), radius=float(f32(radius)), material=int(material),
                                radiance=None if radiance is None else f32(radiance)))
        return len(self.meshes) - 1

    def quad(self, p0, p1, p2, p3, material=0, radiance=None, normal=None):
        """Quad p0..p3 (counter-clockwise seen from the side the normal points to), split (0,1,2),(0,2,3)."""
        p = np.array([p0, p1, p2, p3], dtype=np.float64)
        if normal is None:
            normal = np.cross(p[1] - p[0], p[3] - p[0])
            normal = normal / np.linalg.norm(normal)
        n = np.tile(np.asarray(normal, dtype=np.float64), (4, 1))
        uv = np.array([[0, 0], [1, 0], [1, 1], [0, 1]], dtype=np.float64)
        return self.mesh(p, [[0, 1, 2], [0, 2, 3]], n, uv, material, radiance)

    def box(self, center, half, yaw_deg, material=0, bottom=False):
        """Axis box rotated about +y, 5 faces (6 with bottom), outward normals."""
        c, s = np.cos(np.radians(yaw_deg)), np.sin(np.radians(yaw_deg))
        R = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])
        hx, hy, hz = half
        corners = lambda pts: [R @ np.array(p) + np.array(center) for p in pts]
        faces = [
            [(-hx, -hy, hz), (hx, -hy, hz), (hx, hy, hz), (-hx, hy, hz)],        # +z
            [(hx, -hy, -hz), (-hx, -hy, -hz), (-hx, hy, -hz), (hx, hy, -hz)],    # -z
            [(hx, -hy, hz), (hx, -hy, -hz), (hx, hy, -hz), (hx, hy, hz)],        # +x
            [(-hx, -hy, -hz), (-hx, -hy, hz), (-hx, hy, hz), (-hx, hy, -hz)],    # -x
            [(-hx, hy, hz), (hx, hy, hz), (hx, hy, -hz), (-hx, hy, -hz)],        # +y
        ]
        if bottom:
            faces.append([(-hx, -hy, -hz), (hx, -hy, -hz), (hx, -hy, hz), (-hx, -hy, hz)])
        for f in faces:
            self.quad(*corners(f), material=material)

    # -- outputs ------------------------------------------------------------------------------
    def flat(self) -> FlatScene:
        pos, nrm, uv, idx, pmat, plight, pflags, lights, spheres = [], [], [], [], [], [], [], [], []
        base = 0
        n_prims = 0
        for pos_, inten in self.point_lights:
            # the parser brace-initialises PointLight{position, intensity} into a struct declared {intensity, position}
            # (parse_scene.cpp:723 vs light.h:9-12): the fields arrive swapped, and that is what the reference holds
            lights.append((sio.LIGHT_POINT, -1, tuple(pos_), tuple(inten)))
        for m in self.meshes:
            if m.get("sphere"):
                idx.append(np.array([[len(spheres), 0, 0]], np.int32))
                spheres.append([*m["center"], m["radius"]])
                pmat.append(np.array([m["material"]], np.int32))
                pflags.append(np.array([sio.PRIM_SPHERE], np.uint8))
                if m["radiance"] is not None:
                    plight.append(np.array([len(lights)], np.int32))
                    lights.append((sio.LIGHT_AREA, n_prims, tuple(m["radiance"]), (0.0, 0.0, 0.0)))
                else:
                    plight.append(np.array([-1], np.int32))
                n_prims += 1
                continue
            nv, nf = len(m["positions"]), len(m["indices"])
            pos.append(m["positions"].astype(np.float64))
            nrm.append(ref_normalize(m["normals"].astype(np.float64)))
            uv.append(np.zeros((nv, 2)) if m["uvs"] is None else m["uvs"].astype(np.float64))
            idx.append(m["indices"] + base)
            pmat.append(np.full(nf, m["material"], np.int32))
            pflags.append(np.full(nf, sio.PRIM_HAS_NORMALS | (0 if m["uvs"] is None else sio.PRIM_HAS_UVS), np.uint8))
            if m["radiance"] is not None:
                # one DiffuseAreaLight per face, in shape order (parse_scene.cpp:940-943)
                first = len(lights)
                for f in range(nf):
                    lights.append((sio.LIGHT_AREA, n_prims + f, tuple(m["radiance"]), (0.0, 0.0, 0.0)))
                plight.append(np.arange(first, first + nf, dtype=np.int32))
            else:
                plight.append(np.full(nf, -1, np.int32))
            base += nv
            n_prims += nf
        mats = np.zeros(len(self.materials), sio.MAT_DTYPE)
        for i, m in enumerate(self.materials):
            mats[i]["type"] = m["type"]
            mats[i]["tex_id"] = -1 if m["texture"] is None else m["texture"]
            mats[i]["color"] = 0.0 if (m["texture"] is not None or m["type"] == sio.MAT_DISNEY_CLEARCOAT) else m["color"]
            mats[i]["uv"] = [1, 1, 0, 0] if m["texture"] is None else [*m["uvscale"], *m["uvoffset"]]
            p = m["params"]
            t = m["type"]
            if t in (sio.MAT_MIRROR,):
                mats[i]["p"][0] = 1.0                      # Mirror::eta default (material.h:11-14); parser never sets it
            elif t == sio.MAT_PLASTIC:
                mats[i]["p"][0] = p.get("eta", 1.5)
            elif t in (sio.MAT_PHONG, sio.MAT_BLINN_PHONG, sio.MAT_BLINN_MICROFACET):
                mats[i]["p"][0] = p.get("exponent", 5.0)
            elif t == sio.MAT_DISNEY_DIFFUSE:
                mats[i]["p"][0] = p.get("roughness", 0.5)
                mats[i]["p"][1] = p.get("subsurface", 0.0)
            elif t == sio.MAT_GGX:
                mats[i]["p"][0] = p.get("alpha", 0.14)
        textures = []
        for rgbe in self.textures:
            scale = np.ldexp(np.float32(1.0), rgbe[..., 3].astype(np.int32) - 136).astype(np.float32)
            rgb = (rgbe[..., :3].astype(np.float32) * scale[..., None]).astype(np.float64)
            rgb[rgbe[..., 3] == 0] = 0.0
            textures.append(rgb)
        cat = lambda xs, shape, dt: np.concatenate(xs).astype(dt) if xs else np.zeros(shape, dt)
        return FlatScene(
            self.width, self.height, self.lookfrom, self.lookat, self.up, self.vfov, self.background,
            cat(pos, (0, 3), np.float64), cat(nrm, (0, 3), np.float64), cat(uv, (0, 2), np.float64),
            cat(idx, (0, 3), np.int32), cat(pmat, (0,), np.int32), cat(plight, (0,), np.int32),
            cat(pflags, (0,), np.uint8), np.array(spheres, np.float64).reshape(-1, 4), mats,
            np.array(lights, dtype=sio.LIGHT_DTYPE) if lights else np.zeros(0, sio.LIGHT_DTYPE), textures,
            self.spp, self.env, self.env_sample)._canon()

    def write(self, directory) -> str:
        """Write scene.xml + meshes + textures under `directory`; returns the XML path."""
        os.makedirs(directory, exist_ok=True)
        if self.env is not None:
            raise ValueError("the reference has no environment emitter: scenes with an environment map cannot be written as XML")
        if any(m["type"] == sio.MAT_GGX for m in self.materials):
            raise ValueError("the reference has no GGX BSDF: scenes using the GGX extension cannot be written as XML")
        x = ['<?xml version="1.0" encoding="utf-8"?>', '<scene version="0.5.0">']
        x.append('<sensor type="perspective">')
        x.append(f'  <float name="fov" value="{_fmt(self.vfov)}"/><string name="fovAxis" value="y"/>')
        x.append(f'  <transform name="toWorld"><lookat origin="{_vec(self.lookfrom)}" target="{_vec(self.lookat)}" '
                 f'up="{_vec(self.up)}"/></transform>')
        x.append(f'  <sampler type="independent"><integer name="sampleCount" value="{self.spp}"/></sampler>')
        x.append(f'  <film type="hdrfilm"><integer name="width" value="{self.width}"/>'
                 f'<integer name="height" value="{self.height}"/></film>')
        x.append('</sensor>')
        x.append(f'<background><rgb name="radiance" value="{_vec(self.background)}"/></background>')
        for i, rgbe in enumerate(self.textures):
            write_hdr(os.path.join(directory, f"tex{i}.hdr"), rgbe)
        for i, m in enumerate(self.materials):
            tname, cname = _BSDF_XML[m["type"]]
            x.append(f'<bsdf type="{tname}" id="m{i}">')
            if cname is None:
                pass
            elif m["texture"] is None:
                x.append(f'  <rgb name="{cname}" value="{_vec(m["color"])}"/>')
            else:
                x.append(f'  <texture type="bitmap" name="{cname}"><string name="filename" value="tex{m["texture"]}.hdr"/>'
                         f'<float name="uscale" value="{_fmt(m["uvscale"][0])}"/><float name="vscale" value="{_fmt(m["uvscale"][1])}"/>'
                         f'<float name="uoffset" value="{_fmt(m["uvoffset"][0])}"/><float name="voffset" value="{_fmt(m["uvoffset"][1])}"/></texture>')
            for k, v in m["params"].items():
                x.append(f'  <float name="{k}" value="{_fmt(v)}"/>')
            x.append('</bsdf>')
        for pos_, inten in self.point_lights:     # before the shapes: they take the first light indices
            x.append(f'<emitter type="point"><point name="position" x="{_fmt(pos_[0])}" y="{_fmt(pos_[1])}" z="{_fmt(pos_[2])}"/>'
                     f'<rgb name="intensity" value="{_vec(inten)}"/></emitter>')
        for i, m in enumerate(self.meshes):
            if m.get("sphere"):
                c = m["center"]
                x.append(f'<shape type="sphere"><point name="center" x="{_fmt(c[0])}" y="{_fmt(c[1])}" z="{_fmt(c[2])}"/>'
                         f'<float name="radius" value="{_fmt(m["radius"])}"/><ref id="m{m["material"]}"/>')
                if m["radiance"] is not None:
                    x.append(f'  <emitter type="area"><rgb name="radiance" value="{_vec(m["radiance"])}"/></emitter>')
                x.append('</shape>')
                continue
            write_ply(os.path.join(directory, f"mesh{i}.ply"), m["positions"], m["indices"], m["normals"], m["uvs"])
            x.append(f'<shape type="ply"><string name="filename" value="mesh{i}.ply"/><ref id="m{m["material"]}"/>')
            if m["radiance"] is not None:
                x.append(f'  <emitter type="area"><rgb name="radiance" value="{_vec(m["radiance"])}"/></emitter>')
            x.append('</shape>')
        x.append('</scene>')
        path = os.path.join(directory, "scene.xml")
        with open(path, "w") as f:
            f.write("\n".join(x) + "\n")
        return path


def write_ply(path, positions, indices, normals, uvs=None):
    """binary_little_endian PLY in the form src/parse/parse_ply.cpp:16-31,83-120 reads."""
    nv, nf = len(positions), len(indices)
    props = ["x", "y", "z", "nx", "ny", "nz"] + (["u", "v"] if uvs is not None else [])
    hdr = ["ply", "format binary_little_endian 1.0", f"element vertex {nv}"]
    hdr += [f"property float {p}" for p in props]
    hdr += [f"element face {nf}", "property list uchar int vertex_indices", "end_header"]
    cols = [positions.astype("<f4"), normals.astype("<f4")] + ([uvs.astype("<f4")] if uvs is not None else [])
    verts = np.ascontiguousarray(np.concatenate(cols, axis=1))
    faces = np.zeros(nf, dtype=np.dtype([("n", "u1"), ("i", "<i4", 3)]))
    faces["n"] = 3
    faces["i"] = indices
    with open(path, "wb") as f:
        f.write(("\n".join(hdr) + "\n").encode("ascii"))
        verts.tofile(f)
        faces.tofile(f)


def write_hdr(path, rgbe):
    """Flat (non-RLE) Radiance .hdr; stb_image falls back to flat decoding when a scanline does not
    start with the RLE marker, so give every row's first pixel a mantissa that cannot be mistaken for it."""
    h, w, _ = rgbe.shape
    data = rgbe.copy()
    if 8 <= w < 32768:
        clash = (data[:, 0, 0] == 2) & (data[:, 0, 1] == 2) & (data[:, 0, 2] < 128)
        assert not clash.any(), "first texel of a row looks like an RLE marker"
    with open(path, "wb") as f:
        f.write(b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n" + f"-Y {h} +X {w}\n".encode("ascii"))
        data.tofile(f)


# =============================================================================================
# BASELINE.json configs
# =============================================================================================
def cornell_box(width=512, height=512, spp=64, materials="diffuse") -> SceneBuilder:
    """Config 1: Cornell box, 32 triangles (5 walls, ceiling light, two rotated boxes), 2 area lights
    (each light triangle is its own light, parse_scene.cpp:940-943).  SURVEY.md Appendix D geometry."""
    b = SceneBuilder(width, height, (0, 1, 3.8), (0, 1, 0), (0, 1, 0), 39.3, spp, (0, 0, 0))
    if materials == "diffuse":
        white = b.material(sio.MAT_DIFFUSE, (0.73, 0.73, 0.73))
        red = b.material(sio.MAT_DIFFUSE, (0.65, 0.05, 0.05))
        green = b.material(sio.MAT_DIFFUSE, (0.12, 0.45, 0.15))
        tall = short = white
    else:  # "mixed": one of every implemented BSDF, for shading parity
        white = b.material(sio.MAT_DIFFUSE, (0.73, 0.73, 0.73))
        red = b.material(sio.MAT_PHONG, (0.65, 0.05, 0.05), exponent=20)
        green = b.material(sio.MAT_BLINN_PHONG, (0.12, 0.45, 0.15), exponent=30)
        tall = b.material(sio.MAT_MIRROR, (0.9, 0.9, 0.9))
        short = b.material(sio.MAT_PLASTIC, (0.2, 0.3, 0.7), eta=1.5)
    black = b.material(sio.MAT_DIFFUSE, (0, 0, 0))
    if materials != "diffuse":
        floor = b.material(sio.MAT_BLINN_MICROFACET, (0.6, 0.6, 0.5), exponent=50)
        back = b.material(sio.MAT_DISNEY_DIFFUSE, (0.7, 0.6, 0.5), roughness=0.7, subsurface=0.4)
    else:
        floor = back = white
    b.quad((-1, 0, 1), (1, 0, 1), (1, 0, -1), (-1, 0, -1), floor)          # floor  y=0, normal +y
    b.quad((-1, 2, -1), (1, 2, -1), (1, 2, 1), (-1, 2, 1), white)          # ceiling y=2, normal -y
    b.quad((-1, 0, -1), (1, 0, -1), (1, 2, -1), (-1, 2, -1), back)         # back   z=-1, normal +z
    b.quad((-1, 0, 1), (-1, 0, -1), (-1, 2, -1), (-1, 2, 1), red)          # left   x=-1, normal +x
    b.quad((1, 0, -1), (1, 0, 1), (1, 2, 1), (1, 2, -1), green)            # right  x=+1, normal -x
    b.quad((-0.25, 1.98, -0.25), (0.25, 1.98, -0.25), (0.25, 1.98, 0.25), (-0.25, 1.98, 0.25), black,
           radiance=(17, 12, 4))                                           # light, normal -y
    b.box((-0.35, 0.6, -0.3), (0.3, 0.6, 0.3), 18, tall)
    b.box((0.35, 0.3, 0.35), (0.3, 0.3, 0.3), -17, short)
    return b


def cornell_stubs(width=64, height=64, spp=4, clearcoat=False) -> SceneBuilder:
    """Parity scene for the branches no BASELINE config reaches: one surface per "Disney" alternative the reference only
    stubs (metal, glass, sheen, bsdf evaluate as Lambertian: src/materials/disney_*.inl) and a point emitter next to the
    area light, which contributes nothing but takes a third of the uniform light picks (src/light.cpp:5-7,
    path_tracing.h:33).  clearcoat=True puts DisneyClearcoat on the tall box: the reference's eval for it returns an
    UNINITIALISED vector (`return {};` through `TVector3() {}`, disney_clearcoat.inl:26, vector.h:30), so that variant can
    only be compared between our own CPU restatement and the GPU (both return 0), never pinned to the reference."""
    b = SceneBuilder(width, height, (0, 1, 3.8), (0, 1, 0), (0, 1, 0), 39.3, spp, (0.05, 0.04, 0.03))
    white = b.material(sio.MAT_DIFFUSE, (0.73, 0.73, 0.73))
    metal = b.material(sio.MAT_DISNEY_METAL, (0.8, 0.6, 0.3))
    glass = b.material(sio.MAT_DISNEY_GLASS, (0.6, 0.8, 0.9))
    coat = b.material(sio.MAT_DISNEY_CLEARCOAT) if clearcoat else b.material(sio.MAT_DISNEY_METAL, (0.9, 0.9, 0.2))
    sheen = b.material(sio.MAT_DISNEY_SHEEN, (0.7, 0.2, 0.5))
    principled = b.material(sio.MAT_DISNEY_BSDF, (0.3, 0.7, 0.4))
    black = b.material(sio.MAT_DIFFUSE, (0, 0, 0))
    b.point_light((0.3, 1.2, 0.4), (5, 6, 7))
    _room(b, metal, white, principled, glass, sheen)
    b.quad((-0.25, 1.98, -0.25), (0.25, 1.98, -0.25), (0.25, 1.98, 0.25), (-0.25, 1.98, 0.25), black, radiance=(17, 12, 4))
    b.box((-0.35, 0.6, -0.3), (0.3, 0.6, 0.3), 18, coat)
    b.box((0.35, 0.3, 0.35), (0.3, 0.3, 0.3), -17, white)
    return b


def _grid_mesh(n, extent, height_fn):
    """(n+1)^2 vertices on [-extent,extent]^2 in xz, y = height_fn(x,z); 2 n^2 triangles; smooth normals
    from central differences; uv = grid parameter."""
    g = np.linspace(-extent, extent, n + 1)
    X, Z = np.meshgrid(g, g, indexing="xy")          # row = z index, col = x index
    Y = height_fn(X, Z)
    step = 2.0 * extent / n
    dYdx = np.gradient(Y, step, axis=1)
    dYdz = np.gradient(Y, step, axis=0)
    N = np.stack([-dYdx, np.ones_like(Y), -dYdz], axis=-1)
    N /= np.linalg.norm(N, axis=-1, keepdims=True)
    P = np.stack([X, Y, Z], axis=-1).reshape(-1, 3)
    UV = np.stack([(X + extent) / (2 * extent), (Z + extent) / (2 * extent)], axis=-1).reshape(-1, 2)
    i, j = np.meshgrid(np.arange(n), np.arange(n), indexing="xy")   # i: x cell, j: z cell
    v00 = (j * (n + 1) + i).ravel()
    v10, v01, v11 = v00 + 1, v00 + (n + 1), v00 + (n + 2)
    # counter-clockwise seen from +y: (x,z) -> (x,z+1) -> (x+1,z+1)
    tris = np.empty((2 * n * n, 3), np.int32)
    tris[0::2] = np.stack([v00, v01, v11], axis=1)
    tris[1::2] = np.stack([v00, v11, v10], axis=1)
    return P, tris, N.reshape(-1, 3), UV


def heightfield(n=708, width=1920, height=1080, spp=256, seed=1234, mtype=sio.MAT_BLINN_MICROFACET) -> SceneBuilder:
    """Config 2: displaced grid, n=708 -> 1 002 528 triangles on [-100,100]^2 (+ a 2-triangle quad light),
    blinn_microfacet (the reference's only microfacet BRDF: SURVEY.md section 0 gap 1), one-sample MIS."""
    rng = np.random.default_rng(seed)
    noise = rng.normal(0.0, 1.0, size=(n + 1, n + 1))
    P, T, N, UV = _grid_mesh(n, 100.0, lambda X, Z: 15.0 * np.sin(0.06 * X) * np.cos(0.05 * Z) + noise)
    b = SceneBuilder(width, height, (0, 160, 240), (0, 0, 0), (0, 1, 0), 45.0, spp, (0.05, 0.05, 0.08))
    if mtype == sio.MAT_BLINN_MICROFACET:
        m = b.material(mtype, (0.7, 0.6, 0.4), exponent=100)
    elif mtype == sio.MAT_GGX:
        m = b.material(mtype, (0.7, 0.6, 0.4), alpha=0.14)   # the GGX variant of config 2 (extension, unpinned)
    else:
        m = b.material(mtype, (0.7, 0.6, 0.4))
    black = b.material(sio.MAT_DIFFUSE, (0, 0, 0))
    b.mesh(P, T, N, UV, m)
    b.quad((-50, 200, -50), (50, 200, -50), (50, 200, 50), (-50, 200, 50), black, radiance=(20, 20, 20))  # normal -y
    return b


def multi_light(width=1920, height=1080, spp=1024, n_side=20, seed=3) -> SceneBuilder:
    """Config 4: Cornell-style room scaled x50 with n_side^2 small emissive triangles under the ceiling
    (each its own light), plus diffuse / glossy blockers; multi-sample MIS."""
    S = 50.0
    rng = np.random.default_rng(seed)
    b = SceneBuilder(width, height, (0, S, 3.8 * S), (0, S, 0), (0, 1, 0), 39.3, spp, (0, 0, 0))
    white = b.material(sio.MAT_DIFFUSE, (0.73, 0.73, 0.73))
    red = b.material(sio.MAT_DIFFUSE, (0.65, 0.05, 0.05))
    green = b.material(sio.MAT_DIFFUSE, (0.12, 0.45, 0.15))
    black = b.material(sio.MAT_DIFFUSE, (0, 0, 0))
    glossy = b.material(sio.MAT_BLINN_MICROFACET, (0.8, 0.7, 0.5), exponent=60)
    phong = b.material(sio.MAT_PHONG, (0.5, 0.5, 0.6), exponent=40)
    q = lambda *p, **k: b.quad(*[tuple(S * np.array(v)) for v in p], **k)
    q((-1, 0, 1), (1, 0, 1), (1, 0, -1), (-1, 0, -1), material=white)
    q((-1, 2, -1), (1, 2, -1), (1, 2, 1), (-1, 2, 1), material=white)
    q((-1, 0, -1), (1, 0, -1), (1, 2, -1), (-1, 2, -1), material=white)
    q((-1, 0, 1), (-1, 0, -1), (-1, 2, -1), (-1, 2, 1), material=red)
    q((1, 0, -1), (1, 0, 1), (1, 2, 1), (1, 2, -1), material=green)
    # emissive triangles: one mesh per light so that each can carry its own radiance
    cell = 1.6 * S / n_side
    for j in range(n_side):
        for i in range(n_side):
            cx = -0.8 * S + (i + 0.5) * cell + rng.uniform(-0.15, 0.15) * cell
            cz = -0.8 * S + (j + 0.5) * cell + rng.uniform(-0.15, 0.15) * cell
            y = 1.96 * S - rng.uniform(0, 0.02) * S
            r = 0.3 * cell
            a = rng.uniform(0, 2 * np.pi)
            pts = [(cx + r * np.cos(a + k * 2 * np.pi / 3), y, cz + r * np.sin(a + k * 2 * np.pi / 3)) for k in range(3)]
            rad = rng.uniform(5, 50)
            # wind so that the geometric normal is -y (facing the room); shading normals -y
            b.mesh(pts, [[0, 1, 2]], [(0, -1, 0)] * 3, None, black, radiance=(rad, rad, rad))
    b.box((-0.35 * S, 0.6 * S, -0.3 * S), (0.3 * S, 0.6 * S, 0.3 * S), 18, glossy)
    b.box((0.35 * S, 0.3 * S, 0.35 * S), (0.3 * S, 0.3 * S, 0.3 * S), -17, phong)
    b.box((0.0, 1.2 * S, -0.6 * S), (0.5 * S, 0.02 * S, 0.2 * S), 5, white, bottom=True)
    return b


def _room(b, floor, ceiling, back, left, right):
    b.quad((-1, 0, 1), (1, 0, 1), (1, 0, -1), (-1, 0, -1), floor)
    b.quad((-1, 2, -1), (1, 2, -1), (1, 2, 1), (-1, 2, 1), ceiling)
    b.quad((-1, 0, -1), (1, 0, -1), (1, 2, -1), (-1, 2, -1), back)
    b.quad((-1, 0, 1), (-1, 0, -1), (-1, 2, -1), (-1, 2, 1), left)
    b.quad((1, 0, -1), (1, 0, 1), (1, 2, 1), (1, 2, -1), right)


def procedural_rgbe(w, h, kind="checker", seed=0) -> np.ndarray:
    """Small procedural textures straight in RGBE bytes (exactly representable texels)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    out = np.zeros((h, w, 4), np.uint8)
    if kind == "checker":
        c = ((xx // max(1, w // 8) + yy // max(1, h // 8)) % 2).astype(bool)
        out[..., 0] = np.where(c, 230, 40)
        out[..., 1] = np.where(c, 200, 60)
        out[..., 2] = np.where(c, 90, 150)
    else:
        out[..., :3] = rng.integers(30, 250, size=(h, w, 3))
    out[..., 3] = 128  # mantissa * 2^-8  ->  values in (0, 1)
    out[:, 0, 0] |= 1  # never (2, 2, <128) at a row start: keeps stb_image on its flat-decoding path
    return out


def textured_room(width=256, height=256, spp=16) -> SceneBuilder:
    """Image textures on several BSDFs with non-trivial uv scale / offset (exercises the wrap-around column and
    row of src/texture.cpp:13-24) -- the pinned part of config 3 (the environment map itself has no reference)."""
    b = SceneBuilder(width, height, (0, 1, 3.8), (0, 1, 0), (0, 1, 0), 39.3, spp, (0.2, 0.25, 0.3))
    t0 = b.texture_rgbe(procedural_rgbe(16, 16, "checker"))
    t1 = b.texture_rgbe(procedural_rgbe(13, 7, "noise", seed=5))
    floor = b.material(sio.MAT_DIFFUSE, texture=t0, uvscale=(3.7, 2.3), uvoffset=(0.31, 0.05))
    back = b.material(sio.MAT_BLINN_MICROFACET, texture=t1, uvscale=(1.0, 1.0), uvoffset=(0.0, 0.0), exponent=40)
    left = b.material(sio.MAT_PHONG, texture=t1, uvscale=(-2.5, 4.0), uvoffset=(0.5, 0.5), exponent=12)
    right = b.material(sio.MAT_DISNEY_DIFFUSE, texture=t0, uvscale=(5.0, 5.0), uvoffset=(0.0, 0.0), roughness=0.4, subsurface=0.3)
    white = b.material(sio.MAT_DIFFUSE, (0.73, 0.73, 0.73))
    black = b.material(sio.MAT_DIFFUSE, (0, 0, 0))
    plastic = b.material(sio.MAT_PLASTIC, texture=t0, uvscale=(2.0, 2.0), uvoffset=(0.0, 0.0), eta=1.5)
    _room(b, floor, white, back, left, right)
    b.quad((-0.25, 1.98, -0.25), (0.25, 1.98, -0.25), (0.25, 1.98, 0.25), (-0.25, 1.98, 0.25), black, radiance=(17, 12, 4))
    b.box((-0.35, 0.6, -0.3), (0.3, 0.6, 0.3), 18, plastic)
    b.box((0.35, 0.3, 0.35), (0.3, 0.3, 0.3), -17, floor)
    return b


def sphere_room(width=256, height=256, spp=16) -> SceneBuilder:
    """Spheres (src/shape.cpp:13-42), including a spherical emitter (cone sampling, shape.cpp:125-144)."""
    b = SceneBuilder(width, height, (0, 1, 3.8), (0, 1, 0), (0, 1, 0), 39.3, spp, (0.1, 0.1, 0.1))
    white = b.material(sio.MAT_DIFFUSE, (0.73, 0.73, 0.73))
    red = b.material(sio.MAT_DIFFUSE, (0.65, 0.05, 0.05))
    green = b.material(sio.MAT_DIFFUSE, (0.12, 0.45, 0.15))
    black = b.material(sio.MAT_DIFFUSE, (0, 0, 0))
    mirror = b.material(sio.MAT_MIRROR, (0.9, 0.9, 0.9))
    glossy = b.material(sio.MAT_BLINN_PHONG, (0.3, 0.5, 0.8), exponent=25)
    _room(b, white, white, white, red, green)
    b.sphere((0.0, 1.6, 0.0), 0.2, black, radiance=(20, 18, 15))
    b.sphere((-0.45, 0.4, -0.2), 0.4, mirror)
    b.sphere((0.45, 0.3, 0.3), 0.3, glossy)
    b.quad((-0.2, 1.99, -0.7), (0.2, 1.99, -0.7), (0.2, 1.99, -0.4), (-0.2, 1.99, -0.4), black, radiance=(6, 8, 10))
    return b


def _icosphere(subdiv):
    """Unit icosphere: (vertices [n,3], faces [m,3]); 20 * 4^subdiv faces."""
    phi = (1 + 5 ** 0.5) / 2
    v = np.array([[-1, phi, 0], [1, phi, 0], [-1, -phi, 0], [1, -phi, 0], [0, -1, phi], [0, 1, phi], [0, -1, -phi], [0, 1, -phi],
                  [phi, 0, -1], [phi, 0, 1], [-phi, 0, -1], [-phi, 0, 1]], np.float64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    f = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2], [10, 7, 6],
                  [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5], [2, 4, 11], [6, 2, 10],
                  [8, 6, 7], [9, 8, 1]], np.int64)
    for _ in range(subdiv):
        e = np.sort(np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]]), axis=1)
        ue, inv = np.unique(e, axis=0, return_inverse=True)
        mid = v[ue[:, 0]] + v[ue[:, 1]]
        mid /= np.linalg.norm(mid, axis=1, keepdims=True)
        m = len(v) + inv.reshape(3, -1)          # midpoint ids of edges (01), (12), (20) per face
        v = np.concatenate([v, mid])
        a, b, c = f[:, 0], f[:, 1], f[:, 2]
        f = np.concatenate([np.stack([a, m[0], m[2]], 1), np.stack([b, m[1], m[0]], 1), np.stack([c, m[2], m[1]], 1),
                            np.stack([m[0], m[1], m[2]], 1)])
    return v, f.astype(np.int32)


def instanced_spheres(width=3840, height=2160, spp=4096, copies_side=11, subdiv=6, seed=5, n_lights=64) -> SceneBuilder:
    """Config 5: copies_side^2 copies (random yaw) of a displaced icosphere (20*4^subdiv faces: subdiv 6 = 81 920, so
    11x11 copies + the ground = 10.04 M triangles; the reference has no instancing, so the copies are flattened) on a ground grid,
    extent about +-1000, a mix of diffuse and blinn_microfacet materials, n_lights quad lights."""
    rng = np.random.default_rng(seed)
    sv, sf = _icosphere(subdiv)
    b = SceneBuilder(width, height, (0, 900, 2200), (0, 0, 0), (0, 1, 0), 40.0, spp, (0.02, 0.03, 0.05))
    mats = [b.material(sio.MAT_DIFFUSE, (0.7, 0.7, 0.7)), b.material(sio.MAT_DIFFUSE, (0.7, 0.3, 0.2)),
            b.material(sio.MAT_BLINN_MICROFACET, (0.8, 0.7, 0.4), exponent=80),
            b.material(sio.MAT_BLINN_MICROFACET, (0.5, 0.6, 0.8), exponent=30)]
    black = b.material(sio.MAT_DIFFUSE, (0, 0, 0))
    # bumpy unit sphere shared by all copies
    bump = 1.0 + 0.08 * np.sin(7 * sv[:, 0]) * np.sin(5 * sv[:, 1] + 1.0) * np.sin(6 * sv[:, 2] + 2.0)
    base = sv * bump[:, None]
    pitch = 2000.0 / copies_side
    for j in range(copies_side):
        for i in range(copies_side):
            yaw = rng.uniform(0, 2 * np.pi)
            c, s = np.cos(yaw), np.sin(yaw)
            R = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])
            r = 0.38 * pitch * rng.uniform(0.7, 1.0)
            centre = np.array([-1000 + (i + 0.5) * pitch, r * 1.05, -1000 + (j + 0.5) * pitch])
            P = base @ R.T * r + centre
            N = sv @ R.T
            b.mesh(P, sf, N, None, mats[(i + j * copies_side) % len(mats)])
    gp, gt, gn, guv = _grid_mesh(256, 1100.0, lambda X, Z: 4.0 * np.sin(0.01 * X) * np.cos(0.013 * Z))
    b.mesh(gp, gt, gn, guv, mats[0])
    side = int(np.ceil(np.sqrt(n_lights)))
    for k in range(n_lights):
        lx = -900 + 1800 * ((k % side) + 0.5) / side
        lz = -900 + 1800 * ((k // side) + 0.5) / side
        h = 40.0
        b.quad((lx - h, 700, lz - h), (lx + h, 700, lz - h), (lx + h, 700, lz + h), (lx - h, 700, lz + h), black,
               radiance=(30, 28, 25))
    return b


def sky_environment(width=2048, height=1024, seed=11, sun_peak=5e4, sun_sigma_deg=1.5) -> np.ndarray:
    """Synthetic HDR sky: vertical gradient + ground tint + one Gaussian sun (config 3)."""
    v = (np.arange(height) + 0.5) / height                     # 0 = zenith, 1 = nadir
    u = (np.arange(width) + 0.5) / width
    theta = v[:, None] * np.pi
    phi = u[None, :] * 2 * np.pi - np.pi
    d = np.stack([np.sin(theta) * np.cos(phi), np.cos(theta) * np.ones_like(phi), -np.sin(theta) * np.sin(phi)], axis=-1)
    up = np.clip(d[..., 1], 0, 1)[..., None]
    sky = (1 - up) * np.array([0.9, 0.95, 1.0]) + up * np.array([0.25, 0.45, 0.9])
    ground = np.array([0.18, 0.16, 0.14])
    img = np.where(d[..., 1:2] >= 0, sky, ground)
    rng = np.random.default_rng(seed)
    sun = np.array([np.cos(rng.uniform(0, 2 * np.pi)) * 0.6, 0.7, 0.0])
    sun[2] = np.sqrt(max(0.0, 1 - sun[0] ** 2 - sun[1] ** 2))
    ang = np.arccos(np.clip(d @ sun, -1, 1))
    img = img + sun_peak * np.exp(-0.5 * (ang / np.radians(sun_sigma_deg)) ** 2)[..., None] * np.array([1.0, 0.95, 0.85])
    return img


def ibl_scene(width=1024, height=1024, spp=512, n_objects=64, seed=7, env_size=(2048, 1024), sample_env=True) -> SceneBuilder:
    """Config 3: textured procedural objects (boxes and bumpy icospheres with image textures) on a textured ground, lit by
    an importance-sampled HDR environment map.  The environment part is an extension without a reference implementation."""
    rng = np.random.default_rng(seed)
    b = SceneBuilder(width, height, (0, 14, 34), (0, 2.5, 0), (0, 1, 0), 40.0, spp, (0, 0, 0))
    checker = b.texture_rgbe(procedural_rgbe(64, 64, "checker"))
    noise = b.texture_rgbe(procedural_rgbe(64, 64, "noise", seed=3))
    ground = b.material(sio.MAT_DIFFUSE, texture=checker, uvscale=(12.0, 12.0))
    mats = [b.material(sio.MAT_DIFFUSE, texture=noise, uvscale=(2.0, 2.0)),
            b.material(sio.MAT_BLINN_MICROFACET, texture=checker, uvscale=(3.0, 3.0), exponent=60),
            b.material(sio.MAT_BLINN_MICROFACET, texture=noise, uvscale=(1.0, 1.0), exponent=200),
            b.material(sio.MAT_DIFFUSE, (0.7, 0.7, 0.7))]
    gp, gt, gn, guv = _grid_mesh(64, 40.0, lambda X, Z: 0.15 * np.sin(0.4 * X) * np.cos(0.3 * Z))
    b.mesh(gp, gt, gn, guv, ground)
    sv, sf = _icosphere(4)                                           # 5120 faces
    suv = np.stack([np.arctan2(sv[:, 2], sv[:, 0]) / (2 * np.pi) + 0.5, np.arccos(np.clip(sv[:, 1], -1, 1)) / np.pi], axis=1)
    side = int(np.ceil(np.sqrt(n_objects)))
    for k in range(n_objects):
        cx = -28 + 56 * ((k % side) + 0.5) / side + rng.uniform(-1.5, 1.5)
        cz = -28 + 56 * ((k // side) + 0.5) / side + rng.uniform(-1.5, 1.5)
        r = rng.uniform(1.0, 2.4)
        m = mats[k % len(mats)]
        if k % 3 == 0:
            b.box((cx, r, cz), (r * 0.8, r, r * 0.8), rng.uniform(0, 90), m, bottom=True)
        else:
            bump = 1.0 + 0.1 * np.sin(6 * sv[:, 0] + k) * np.sin(5 * sv[:, 1]) * np.sin(7 * sv[:, 2])
            b.mesh(sv * bump[:, None] * r + np.array([cx, r * 1.05, cz]), sf, sv, suv, m)
    b.environment(sky_environment(*env_size), sample=sample_env)
    return b


def build(name: str, **kw) -> SceneBuilder:
    return {"cornell": cornell_box, "heightfield": heightfield, "multi_light": multi_light,
            "textured": textured_room, "spheres": sphere_room, "instanced": instanced_spheres, "ibl": ibl_scene, "stubs": cornell_stubs}[name](**kw)
