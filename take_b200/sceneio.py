"""Flat scene description (the host-side mirror of TaKe's `Scene` aggregate).

`FlatScene` holds exactly what the reference's hot path reads from `Scene`
(reference: src/scene.h:13-33) after the variants have been resolved to POD:
one global vertex pool, one primitive list in *shape order* (the reference
pushes one `Triangle` shape per mesh face, src/parse/parse_scene.cpp:937-945,
so primitive id == index into `scene.shapes`), materials, lights, textures,
camera and background.  It converts to the C-ABI `TakeSceneDesc`
(include/take_gpu.h) and (de)serialises as a TAKESCN1 file:

    char  magic[8] = "TAKESCN1"
    i64   hdr[8]   = nv, np, nmat, ntex, nlight, nsphere, spp, 0
    i64   width, height
    f64   lookfrom[3], lookat[3], up[3], vfov          (src/camera.h:5-11)
    f64   background[3]                                (src/scene.h:23)
    f64   positions[3*nv], normals[3*nv], uvs[2*nv]    (src/shape.h:13-18, concatenated over meshes)
    i32   indices[3*np], prim_material[np], prim_light[np]  (+ one i32 pad if 5*np is odd)
    u8    prim_flags[np] padded to a multiple of 8     (bit0 mesh has normals, bit1 mesh has uvs, bit2 sphere)
    f64   spheres[4*nsphere]                           (center xyz, radius; src/shape.h:20-23)
    MatRec   materials[nmat]   {i32 type, i32 tex_id, f64 color[3], f64 uscale,vscale,uoffset,voffset, f64 p[2]}
    LightRec lights[nlight]    {i32 kind, i32 prim_id, f64 intensity[3], f64 position[3]}
    per texture: i64 w, h ; f64 rgb[w*h*3]             (src/image.h:13-39 layout, row 0 first)
    if hdr[7] & 1: i64 w, h ; f64 rgb[w*h*3]           environment map (extension); hdr[7] & 2 = env_sample
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

# Material::index() order of the reference's std::variant (src/material.h:82-93).
MAT_DIFFUSE, MAT_MIRROR, MAT_PLASTIC, MAT_PHONG, MAT_BLINN_PHONG, MAT_BLINN_MICROFACET = range(6)
MAT_DISNEY_DIFFUSE, MAT_DISNEY_METAL, MAT_DISNEY_GLASS, MAT_DISNEY_CLEARCOAT, MAT_DISNEY_SHEEN, MAT_DISNEY_BSDF = range(6, 12)
MAT_GGX = 12   # EXTENSION: no counterpart in the reference (include/take_gpu.h)

LIGHT_POINT, LIGHT_AREA = 0, 1
PRIM_HAS_NORMALS, PRIM_HAS_UVS, PRIM_SPHERE = 1, 2, 4

MAT_DTYPE = np.dtype([("type", "<i4"), ("tex_id", "<i4"), ("color", "<f8", 3), ("uv", "<f8", 4), ("p", "<f8", 2)])
LIGHT_DTYPE = np.dtype([("kind", "<i4"), ("prim_id", "<i4"), ("intensity", "<f8", 3), ("position", "<f8", 3)])
assert MAT_DTYPE.itemsize == 80 and LIGHT_DTYPE.itemsize == 56


class TakeCamera(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("lookfrom", C.c_double * 3),
                ("lookat", C.c_double * 3), ("up", C.c_double * 3), ("vfov", C.c_double)]


class TakeTextureDesc(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("rgb", C.c_void_p)]


class TakeSceneDesc(C.Structure):
    _fields_ = [
        ("camera", TakeCamera),
        ("background", C.c_double * 3),
        ("num_vertices", C.c_int64),
        ("positions", C.c_void_p), ("normals", C.c_void_p), ("uvs", C.c_void_p),
        ("num_prims", C.c_int64),
        ("indices", C.c_void_p), ("prim_material", C.c_void_p), ("prim_light", C.c_void_p), ("prim_flags", C.c_void_p),
        ("num_spheres", C.c_int64), ("spheres", C.c_void_p),
        ("num_materials", C.c_int32), ("num_textures", C.c_int32), ("num_lights", C.c_int32), ("reserved", C.c_int32),
        ("materials", C.c_void_p), ("textures", C.c_void_p), ("lights", C.c_void_p),
        ("env_width", C.c_int32), ("env_height", C.c_int32), ("env_sample", C.c_int32), ("reserved2", C.c_int32),
        ("env_rgb", C.c_void_p),
    ]


@dataclass
class FlatScene:
    width: int
    height: int
    lookfrom: np.ndarray
    lookat: np.ndarray
    up: np.ndarray
    vfov: float
    background: np.ndarray
    positions: np.ndarray          # f64 [nv,3]
    normals: np.ndarray            # f64 [nv,3]
    uvs: np.ndarray                # f64 [nv,2]
    indices: np.ndarray            # i32 [np,3]
    prim_material: np.ndarray      # i32 [np]
    prim_light: np.ndarray         # i32 [np]
    prim_flags: np.ndarray         # u8  [np]
    spheres: np.ndarray            # f64 [ns,4]
    materials: np.ndarray          # MAT_DTYPE [nmat]
    lights: np.ndarray             # LIGHT_DTYPE [nlight]
    textures: list = field(default_factory=list)   # list of f64 [h,w,3]
    spp: int = 16
    env: object = None             # optional lat-long environment map f64 [h,w,3] (extension, see take_gpu.h)
    env_sample: bool = False

    @property
    def num_prims(self) -> int:
        return int(self.prim_material.shape[0])

    def _canon(self):
        c = np.ascontiguousarray
        self.lookfrom = c(self.lookfrom, dtype=np.float64)
        self.lookat = c(self.lookat, dtype=np.float64)
        self.up = c(self.up, dtype=np.float64)
        self.background = c(self.background, dtype=np.float64)
        self.positions = c(self.positions, dtype=np.float64).reshape(-1, 3)
        self.normals = c(self.normals, dtype=np.float64).reshape(-1, 3)
        self.uvs = c(self.uvs, dtype=np.float64).reshape(-1, 2)
        self.indices = c(self.indices, dtype=np.int32).reshape(-1, 3)
        self.prim_material = c(self.prim_material, dtype=np.int32)
        self.prim_light = c(self.prim_light, dtype=np.int32)
        self.prim_flags = c(self.prim_flags, dtype=np.uint8)
        self.spheres = c(self.spheres, dtype=np.float64).reshape(-1, 4)
        self.materials = c(self.materials, dtype=MAT_DTYPE)
        self.lights = c(self.lights, dtype=LIGHT_DTYPE)
        self.textures = [c(t, dtype=np.float64) for t in self.textures]
        if self.env is not None:
            self.env = c(self.env, dtype=np.float64)
        return self

    # ---- C-ABI view -------------------------------------------------------------------------
    def to_desc(self) -> TakeSceneDesc:
        """TakeSceneDesc whose pointers alias this object's arrays (keep `self` alive while in use)."""
        self._canon()
        d = TakeSceneDesc()
        d.camera.width, d.camera.height = self.width, self.height
        d.camera.lookfrom[:] = self.lookfrom.tolist()
        d.camera.lookat[:] = self.lookat.tolist()
        d.camera.up[:] = self.up.tolist()
        d.camera.vfov = float(self.vfov)
        d.background[:] = self.background.tolist()
        d.num_vertices = self.positions.shape[0]
        d.positions, d.normals, d.uvs = (a.ctypes.data for a in (self.positions, self.normals, self.uvs))
        d.num_prims = self.num_prims
        d.indices = self.indices.ctypes.data
        d.prim_material = self.prim_material.ctypes.data
        d.prim_light = self.prim_light.ctypes.data
        d.prim_flags = self.prim_flags.ctypes.data
        d.num_spheres = self.spheres.shape[0]
        d.spheres = self.spheres.ctypes.data
        d.num_materials, d.num_textures, d.num_lights = len(self.materials), len(self.textures), len(self.lights)
        d.materials = self.materials.ctypes.data
        d.lights = self.lights.ctypes.data
        self._tex_descs = (TakeTextureDesc * max(1, len(self.textures)))()
        for i, t in enumerate(self.textures):
            self._tex_descs[i].height, self._tex_descs[i].width = t.shape[0], t.shape[1]
            self._tex_descs[i].rgb = t.ctypes.data
        d.textures = C.addressof(self._tex_descs)
        if self.env is not None:
            d.env_height, d.env_width = self.env.shape[0], self.env.shape[1]
            d.env_sample = 1 if self.env_sample else 0
            d.env_rgb = self.env.ctypes.data
        return d

    # ---- TAKESCN1 ---------------------------------------------------------------------------
    def save(self, path):
        self._canon()
        n_p = self.num_prims
        with open(path, "wb") as f:
            f.write(b"TAKESCN1")
            np.array([self.positions.shape[0], n_p, len(self.materials), len(self.textures), len(self.lights),
                      self.spheres.shape[0], self.spp,
                      (1 if self.env is not None else 0) | (2 if self.env_sample else 0)], dtype="<i8").tofile(f)
            np.array([self.width, self.height], dtype="<i8").tofile(f)
            np.concatenate([self.lookfrom, self.lookat, self.up, [self.vfov], self.background]).astype("<f8").tofile(f)
            for a in (self.positions, self.normals, self.uvs, self.indices, self.prim_material, self.prim_light):
                a.tofile(f)
            if (5 * n_p) % 2:
                np.zeros(1, "<i4").tofile(f)
            flags = np.zeros((n_p + 7) & ~7, np.uint8)
            flags[:n_p] = self.prim_flags
            flags.tofile(f)
            self.spheres.tofile(f)
            self.materials.tofile(f)
            self.lights.tofile(f)
            for t in self.textures:
                np.array([t.shape[1], t.shape[0]], dtype="<i8").tofile(f)
                t.tofile(f)
            if self.env is not None:
                np.array([self.env.shape[1], self.env.shape[0]], dtype="<i8").tofile(f)
                self.env.tofile(f)

    @staticmethod
    def load(path) -> "FlatScene":
        buf = np.fromfile(path, dtype=np.uint8)
        if bytes(buf[:8]) != b"TAKESCN1":
            raise ValueError(f"{path}: not a TAKESCN1 file")
        off = [8]

        def take(dtype, count):
            dt = np.dtype(dtype)
            a = np.frombuffer(buf, dtype=dt, count=count, offset=off[0]).copy()
            off[0] += dt.itemsize * count
            return a

        nv, n_p, nmat, ntex, nlight, nsph, spp, envflags = (int(v) for v in take("<i8", 8))
        w, h = (int(v) for v in take("<i8", 2))
        cam = take("<f8", 13)
        pos, nrm, uv = take("<f8", 3 * nv), take("<f8", 3 * nv), take("<f8", 2 * nv)
        idx, pmat, plight = take("<i4", 3 * n_p), take("<i4", n_p), take("<i4", n_p)
        if (5 * n_p) % 2:
            take("<i4", 1)
        flags = take(np.uint8, (n_p + 7) & ~7)[:n_p]
        sph = take("<f8", 4 * nsph)
        mats, lights = take(MAT_DTYPE, nmat), take(LIGHT_DTYPE, nlight)
        textures = []
        for _ in range(ntex):
            tw, th = (int(v) for v in take("<i8", 2))
            textures.append(take("<f8", tw * th * 3).reshape(th, tw, 3))
        env = None
        if envflags & 1:
            ew, eh = (int(v) for v in take("<i8", 2))
            env = take("<f8", ew * eh * 3).reshape(eh, ew, 3)
        return FlatScene(w, h, cam[0:3], cam[3:6], cam[6:9], float(cam[9]), cam[10:13], pos, nrm, uv, idx, pmat,
                         plight, flags, sph, mats, lights, textures, spp, env, bool(envflags & 2))._canon()

    def same_as(self, other: "FlatScene") -> list:
        """Names of fields that differ (bit-for-bit on values; -0.0 == 0.0)."""
        self._canon(); other._canon()
        bad = []
        for name in ("width", "height", "vfov"):
            if getattr(self, name) != getattr(other, name):
                bad.append(name)
        for name in ("lookfrom", "lookat", "up", "background", "positions", "normals", "uvs", "indices",
                     "prim_material", "prim_light", "prim_flags", "spheres", "materials", "lights"):
            a, b = getattr(self, name), getattr(other, name)
            if a.shape != b.shape or not np.array_equal(a, b):
                bad.append(name)
        if len(self.textures) != len(other.textures) or any(
                a.shape != b.shape or not np.array_equal(a, b) for a, b in zip(self.textures, other.textures)):
            bad.append("textures")
        if (self.env is None) != (other.env is None) or (self.env is not None and not np.array_equal(self.env, other.env)) \
                or self.env_sample != other.env_sample:
            bad.append("env")
        return bad
