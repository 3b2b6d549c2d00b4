"""Host-side mirror of the reference's render interface, over the C-ABI library (include/take_gpu.h).

Reference call stack being replaced:
    main (src/main.cpp:8) -> render(params) (src/render.cpp:9) -> parse_scene -> build_bvh -> tile loop
    -> path_tracing (src/integrator/path_tracing.h) -> scene_intersect / scene_occluded (src/scene.cpp:25-64)

`GpuScene` = the reference's `Scene` after `build_bvh` (acceleration structures built, data resident on one
B200); its methods carry the reference's names: `intersect` = scene_intersect, `occluded` = scene_occluded,
`render` = the body of render() after parsing.  There is no CPU path: if libtake_gpu.so is missing or no CUDA
device is usable, everything here raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .sceneio import FlatScene, TakeSceneDesc

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TAKE_GPU_LIB") or os.path.join(_HERE, "libtake_gpu.so")   # TAKE_GPU_LIB: A/B builds while tuning

INTEGRATORS = {"mis": 0, "raw": 1, "one_sample_mis": 2, "one_sample_mis_power": 3}
ISECT_FAST, ISECT_EXACT = 0, 1
RENDER_NO_SORT, RENDER_STAGE_TIMES, RENDER_COUNT_TESTS, RENDER_RUSSIAN_ROULETTE = 1, 2, 4, 8

EXPORTS = [
    "take_gpu_device_count", "take_gpu_scene_create", "take_gpu_scene_destroy", "take_gpu_intersect",
    "take_gpu_occluded", "take_gpu_intersect_device", "take_gpu_render", "take_gpu_render_device",
    "take_gpu_radiance_samples", "take_gpu_render_multi", "take_gpu_scene_stream", "take_gpu_scene_info", "take_gpu_last_error",
    "take_gpu_version", "take_gpu_exr_packed_size", "take_gpu_exr_pack_device", "take_gpu_exr_pack", "take_gpu_exr_write_packed",
    "take_gpu_render_to_exr", "take_gpu_render_async", "take_gpu_render_wait",
    "take_gpu_multi_create", "take_gpu_multi_render", "take_gpu_multi_destroy",
    "take_gpu_builder_create", "take_gpu_builder_destroy", "take_gpu_builder_add_ply", "take_gpu_builder_add_mesh",
    "take_gpu_builder_add_sphere", "take_gpu_builder_add_point_light", "take_gpu_builder_finish", "take_gpu_builder_timings",
    "take_gpu_scene_desc_save", "take_gpu_scene_create_timings", "take_gpu_scene_debug_tree",
    "take_gpu_scene_provisional_stats", "take_gpu_release_cached_memory",
]

RAY_DTYPE = np.dtype([("origin", "<f8", 3), ("dir", "<f8", 3), ("tmin", "<f8"), ("tmax", "<f8")])
HIT_DTYPE = np.dtype([("prim_id", "<i4"), ("pad", "<i4"), ("t", "<f8"), ("u", "<f8"), ("v", "<f8")])
assert RAY_DTYPE.itemsize == 64 and HIT_DTYPE.itemsize == 32


class TakeRenderOpts(C.Structure):
    _fields_ = [("integrator", C.c_int32), ("max_depth", C.c_int32), ("spp_begin", C.c_int64), ("spp_end", C.c_int64),
                ("seed", C.c_uint64), ("flags", C.c_int32), ("reserved", C.c_int32)]


class TakeStats(C.Structure):
    _fields_ = [("samples", C.c_int64), ("extend_rays", C.c_int64), ("shadow_rays", C.c_int64), ("shaded", C.c_int64),
                ("box_tests", C.c_int64), ("tri_tests", C.c_int64), ("kernel_launches", C.c_int64),
                ("ms_total", C.c_double), ("ms_generate", C.c_double), ("ms_extend", C.c_double),
                ("ms_shade", C.c_double), ("ms_shadow", C.c_double), ("ms_sort", C.c_double), ("ms_other", C.c_double),
                ("shadow_box_tests", C.c_int64), ("shadow_tri_tests", C.c_int64), ("miss_after_light_sample", C.c_int64),
                ("waves", C.c_int64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class TakeGpuError(RuntimeError):
    pass


def kernel_source_hash() -> str:
    """Short hash of the sources the KERNELS of libtake_gpu.so are built from (*.cu, *.cuh, the headers they include and the
    Makefile with its flags; not the host-only *.cpp files: mesh loading, EXR writing, host tree builders): profile summaries
    under profiles/ carry it, so a number read from a capture can be tied to the kernels it was captured from (bench.py quotes
    ncu figures only when it matches)."""
    import glob
    import hashlib
    h = hashlib.sha1()
    d = os.path.join(_HERE, "csrc")
    for p in sorted(glob.glob(os.path.join(d, "*.cu")) + glob.glob(os.path.join(d, "*.cuh")) + glob.glob(os.path.join(d, "*.h")) +
                    [os.path.join(d, "Makefile")]):
        h.update(os.path.basename(p).encode())
        h.update(open(p, "rb").read())
    return h.hexdigest()[:12]


_lib = None


def load_library(path: str = LIB_PATH):
    """dlopen the C-ABI library.  Raises if it has not been built (python __graft_entry__.py / make -C take_b200/csrc)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(path):
        raise TakeGpuError(f"{path} not found: build the CUDA extension first (there is no CPU fallback)")
    L = C.CDLL(path)
    vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int
    L.take_gpu_last_error.restype = C.c_char_p
    L.take_gpu_version.restype = C.c_char_p
    L.take_gpu_device_count.argtypes = [C.POINTER(C.c_int)]
    L.take_gpu_scene_create.argtypes = [i32, C.POINTER(TakeSceneDesc), C.POINTER(vp)]
    L.take_gpu_scene_destroy.argtypes = [vp]
    L.take_gpu_intersect.argtypes = [vp, vp, i64, vp, i32]
    L.take_gpu_intersect_device.argtypes = [vp, vp, i64, vp, i32]
    L.take_gpu_occluded.argtypes = [vp, vp, i64, vp]
    L.take_gpu_render.argtypes = [vp, C.POINTER(TakeRenderOpts), vp, vp, C.POINTER(TakeStats)]
    L.take_gpu_render_device.argtypes = [vp, C.POINTER(TakeRenderOpts), vp, vp, C.POINTER(TakeStats)]
    L.take_gpu_radiance_samples.argtypes = [vp, C.POINTER(TakeRenderOpts), i64, vp, vp, vp, vp]
    L.take_gpu_scene_stream.restype = vp
    L.take_gpu_scene_stream.argtypes = [vp]
    L.take_gpu_scene_info.argtypes = [vp, vp]
    L.take_gpu_scene_create_timings.argtypes = [vp, vp]
    L.take_gpu_scene_provisional_stats.argtypes = [vp, vp]
    L.take_gpu_scene_debug_tree.restype = i64
    L.take_gpu_scene_debug_tree.argtypes = [vp, vp, vp]
    L.take_gpu_render_async.argtypes = [vp, C.POINTER(TakeRenderOpts), vp, vp, C.POINTER(i64)]
    L.take_gpu_render_wait.argtypes = [vp, i64, C.POINTER(TakeStats)]
    L.take_gpu_exr_packed_size.restype = i64
    L.take_gpu_exr_packed_size.argtypes = [C.c_int32, C.c_int32]
    L.take_gpu_exr_pack_device.argtypes = [vp, vp, i64, vp]
    L.take_gpu_exr_pack.argtypes = [vp, vp, i64, vp]
    L.take_gpu_exr_write_packed.argtypes = [C.c_char_p, C.c_int32, C.c_int32, vp, C.c_int32]
    L.take_gpu_render_to_exr.argtypes = [vp, C.POINTER(TakeRenderOpts), C.c_char_p, C.POINTER(TakeStats)]
    L.take_gpu_builder_create.argtypes = [C.POINTER(vp)]
    L.take_gpu_builder_destroy.argtypes = [vp]
    L.take_gpu_builder_add_ply.argtypes = [vp, C.c_char_p, vp, vp, C.c_int32, C.c_int32, vp]
    L.take_gpu_builder_add_mesh.argtypes = [vp, i64, vp, vp, vp, i64, vp, C.c_int32, C.c_int32, vp]
    L.take_gpu_builder_add_sphere.argtypes = [vp, vp, C.c_double, C.c_int32, vp]
    L.take_gpu_builder_add_point_light.argtypes = [vp, vp, vp]
    L.take_gpu_builder_finish.argtypes = [vp, C.POINTER(TakeSceneDesc)]
    L.take_gpu_builder_timings.argtypes = [vp, vp]
    L.take_gpu_scene_desc_save.argtypes = [C.POINTER(TakeSceneDesc), i64, C.c_char_p]
    L.take_gpu_multi_create.argtypes = [i32, vp, C.POINTER(TakeSceneDesc), C.POINTER(vp)]
    L.take_gpu_multi_render.argtypes = [vp, C.POINTER(TakeRenderOpts), vp, vp, C.POINTER(TakeStats)]
    L.take_gpu_multi_destroy.argtypes = [vp]
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        raise TakeGpuError(f"take_gpu error {rc}: {load_library().take_gpu_last_error().decode(errors='replace')}")


def device_count() -> int:
    n = C.c_int(0)
    _check(load_library().take_gpu_device_count(C.byref(n)))
    return n.value


def make_rays(origin, direction, tmin=1e-7, tmax=np.inf) -> np.ndarray:
    """Pack rays as the reference's `Ray` (src/ray.h:4-9): n x 8 doubles."""
    origin = np.asarray(origin, np.float64); direction = np.asarray(direction, np.float64)
    n = max(origin.reshape(-1, 3).shape[0], direction.reshape(-1, 3).shape[0])
    rays = np.empty((n, 8), np.float64)
    rays[:, 0:3], rays[:, 3:6], rays[:, 6], rays[:, 7] = origin, direction, tmin, tmax
    return rays


class HostBuild:
    """Handle on the host-built acceleration structures of a scene (take_gpu_host_build): build once, save, load in the
    other ranks of a one-process-per-GPU job, create every rank's GpuScene from it (`GpuScene(flat, prebuilt=...)`)."""

    def __init__(self, flat: FlatScene = None, path: str = None):
        self.lib = L = load_library()
        L.take_gpu_host_build.argtypes = [C.POINTER(TakeSceneDesc), C.POINTER(C.c_void_p)]
        L.take_gpu_host_build_free.argtypes = [C.c_void_p]
        L.take_gpu_host_build_save.argtypes = [C.c_void_p, C.c_char_p]
        L.take_gpu_host_build_load.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
        L.take_gpu_scene_create_prebuilt.argtypes = [C.c_int, C.POINTER(TakeSceneDesc), C.c_void_p, C.POINTER(C.c_void_p)]
        self.h = C.c_void_p()
        if path is not None:
            _check(L.take_gpu_host_build_load(os.fsencode(path), C.byref(self.h)))
        else:
            desc = flat.to_desc()
            _check(L.take_gpu_host_build(C.byref(desc), C.byref(self.h)))

    def save(self, path: str):
        _check(self.lib.take_gpu_host_build_save(self.h, os.fsencode(path)))

    def arrays(self) -> dict:
        """Copies of the structures for inspection (layouts: take_b200/csrc/bvh_build.h)."""
        return _host_build_arrays(self.lib, self.h)

    def close(self):
        if getattr(self, "h", None):
            self.lib.take_gpu_host_build_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class GpuScene:
    def __init__(self, flat: FlatScene, device: int = 0, prebuilt: "HostBuild" = None):
        self.lib = load_library()
        self.flat = flat
        self.width, self.height = flat.width, flat.height
        self._desc = flat.to_desc()
        h = C.c_void_p()
        if prebuilt is not None:
            _check(self.lib.take_gpu_scene_create_prebuilt(device, C.byref(self._desc), prebuilt.h, C.byref(h)))
        else:
            _check(self.lib.take_gpu_scene_create(device, C.byref(self._desc), C.byref(h)))
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.lib.take_gpu_scene_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self) -> dict:
        out = (C.c_double * 6)()
        _check(self.lib.take_gpu_scene_info(self.h, out))
        keys = ["build_ms_reference_tree", "build_ms_fast_tree", "fast_tree_depth", "sah_cost", "fast_nodes", "sm_count"]
        return dict(zip(keys, out))

    def create_timings(self) -> dict:
        """Milliseconds of take_gpu_scene_create by phase (does not wait for the background reference-order tree)."""
        out = (C.c_double * 8)()
        _check(self.lib.take_gpu_scene_create_timings(self.h, out))
        keys = ["validate_ms", "upload_ms", "host_boxes_ms", "device_build_ms", "records_ms", "total_ms", "device_built", "reference_tree_pending"]
        return dict(zip(keys, out))

    def provisional_stats(self) -> dict:
        """Renders that ran before the background reference-order tree had arrived, and how many of them were repeated."""
        out = (C.c_int64 * 2)()
        _check(self.lib.take_gpu_scene_provisional_stats(self.h, out))
        return {"provisional_renders": int(out[0]), "provisional_reruns": int(out[1])}

    def debug_tree(self, leaf_prims=True):
        """(wide nodes as a structured array, primitive id per leaf slot or None): the fast tree as it sits on the device."""
        nw = self.lib.take_gpu_scene_debug_tree(self.h, None, None)
        if nw < 0:
            _check(int(nw))
        wide = np.zeros(nw, WIDE_NODE_DTYPE)
        lp = np.zeros(self.flat.num_prims, np.int32) if leaf_prims else None
        rc = self.lib.take_gpu_scene_debug_tree(self.h, wide.ctypes.data, lp.ctypes.data if (lp is not None and lp.size) else None)
        if rc < 0:
            _check(int(rc))
        return wide, lp

    @property
    def stream(self) -> int:
        return int(self.lib.take_gpu_scene_stream(self.h) or 0)

    # scene_intersect (src/scene.cpp:25-47)
    def intersect(self, rays, exact: bool = False):
        rays = np.ascontiguousarray(rays, np.float64).reshape(-1, 8)
        hits = np.empty(len(rays), HIT_DTYPE)
        _check(self.lib.take_gpu_intersect(self.h, rays.ctypes.data, len(rays), hits.ctypes.data,
                                           ISECT_EXACT if exact else ISECT_FAST))
        return hits["prim_id"].copy(), hits["t"].copy(), np.stack([hits["u"], hits["v"]], axis=1)

    # scene_occluded (src/scene.cpp:49-64)
    def occluded(self, rays):
        rays = np.ascontiguousarray(rays, np.float64).reshape(-1, 8)
        occ = np.empty(len(rays), np.uint8)
        _check(self.lib.take_gpu_occluded(self.h, rays.ctypes.data, len(rays), occ.ctypes.data))
        return occ

    def _opts(self, integrator, max_depth, spp_begin, spp_end, seed, flags=0):
        # (rr_start: first loop iteration of the Russian-roulette extension, used with RENDER_RUSSIAN_ROULETTE; 0 = default 3)
        return TakeRenderOpts(INTEGRATORS[integrator], max_depth, spp_begin, spp_end, seed, flags, getattr(self, "rr_start", 0))

    # the tile loop of render() (src/render.cpp:59-82): sums of samples [spp_begin, spp_end) per pixel
    def render_sums(self, integrator="mis", max_depth=5, spp_begin=0, spp_end=1, seed=0, sumsq=True, flags=0, out=None):
        """The blocking take_gpu_render.  `out` = (sum, sumsq) host arrays to fill (page-locked ones make the copy 3x faster)."""
        if out is not None:
            s, s2 = out
            assert s.dtype == np.float64 and s.flags["C_CONTIGUOUS"] and s.size == self.height * self.width * 3
            sumsq = s2 is not None
        else:
            s = np.empty((self.height, self.width, 3), np.float64)
            s2 = np.empty_like(s) if sumsq else None
        st = TakeStats()
        o = self._opts(integrator, max_depth, spp_begin, spp_end, seed, flags)
        _check(self.lib.take_gpu_render(self.h, C.byref(o), s.ctypes.data, s2.ctypes.data if sumsq else None, C.byref(st)))
        return s, s2, st.as_dict()

    def render_async(self, sum_out: np.ndarray, sumsq_out, integrator="mis", max_depth=5, spp_begin=0, spp_end=1, seed=0, flags=0) -> int:
        """Queue a render into caller-owned (ideally page-locked) host arrays; returns the ticket for render_wait()."""
        assert sum_out.dtype == np.float64 and sum_out.flags["C_CONTIGUOUS"] and sum_out.size == self.height * self.width * 3
        t = C.c_int64(-1)
        o = self._opts(integrator, max_depth, spp_begin, spp_end, seed, flags)
        _check(self.lib.take_gpu_render_async(self.h, C.byref(o), sum_out.ctypes.data,
                                              sumsq_out.ctypes.data if sumsq_out is not None else None, C.byref(t)))
        return t.value

    def render_wait(self, ticket: int) -> dict:
        st = TakeStats()
        _check(self.lib.take_gpu_render_wait(self.h, ticket, C.byref(st)))
        return st.as_dict()

    def render_sums_device(self, d_sum_ptr: int, d_sumsq_ptr: int, integrator="mis", max_depth=5, spp_begin=0, spp_end=1,
                           seed=0, flags=0):
        """Accumulate into caller-owned DEVICE buffers (e.g. torch tensors' data_ptr()) -- no host copies."""
        st = TakeStats()
        o = self._opts(integrator, max_depth, spp_begin, spp_end, seed, flags)
        _check(self.lib.take_gpu_render_device(self.h, C.byref(o), d_sum_ptr, d_sumsq_ptr or None, C.byref(st)))
        return st.as_dict()

    # the output step (imwrite .exr, src/image.cpp:157-175): device half -> filtered 16-scanline blocks of half B,G,R
    def exr_pack(self, sum_rgb, spp: int) -> np.ndarray:
        sum_rgb = np.ascontiguousarray(sum_rgb, np.float64)
        assert sum_rgb.shape == (self.height, self.width, 3)
        out = np.empty(self.lib.take_gpu_exr_packed_size(self.width, self.height), np.uint8)
        _check(self.lib.take_gpu_exr_pack(self.h, sum_rgb.ctypes.data, spp, out.ctypes.data))
        return out

    def exr_pack_device(self, d_sum_ptr: int, spp: int) -> np.ndarray:
        out = np.empty(self.lib.take_gpu_exr_packed_size(self.width, self.height), np.uint8)
        _check(self.lib.take_gpu_exr_pack_device(self.h, d_sum_ptr, spp, out.ctypes.data))
        return out

    def render_to_exr(self, path, integrator="mis", max_depth=5, spp_begin=0, spp_end=1, seed=0, flags=0) -> dict:
        """render() + imwrite("image.exr") (src/main.cpp:21-24) without bringing the FP64 sums to the host."""
        st = TakeStats()
        o = self._opts(integrator, max_depth, spp_begin, spp_end, seed, flags)
        _check(self.lib.take_gpu_render_to_exr(self.h, C.byref(o), os.fsencode(path), C.byref(st)))
        return st.as_dict()

    def radiance_samples(self, px, py, s, integrator="mis", max_depth=5, seed=0):
        px = np.ascontiguousarray(px, np.int32); py = np.ascontiguousarray(py, np.int32); s = np.ascontiguousarray(s, np.int64)
        out = np.empty((len(px), 3), np.float64)
        o = self._opts(integrator, max_depth, 0, 0, seed)
        _check(self.lib.take_gpu_radiance_samples(self.h, C.byref(o), len(px), px.ctypes.data, py.ctypes.data,
                                                  s.ctypes.data, out.ctypes.data))
        return out


class DescBuilder:
    """take_gpu_builder_*: big meshes straight into the flat arrays of a TakeSceneDesc on all host threads (replaces
    parse_ply + compute_normals + the per-face Shape / light expansion of the reference's parser for the shapes routed
    through it).  Shapes must be added in scene-file order.  `arrays()` copies the geometry / light arrays out."""

    def __init__(self):
        self.lib = load_library()
        self.h = C.c_void_p()
        _check(self.lib.take_gpu_builder_create(C.byref(self.h)))

    @staticmethod
    def _vec(a, n):
        if a is None:
            return None, None
        a = np.ascontiguousarray(a, np.float64).reshape(-1)
        assert a.size == n
        return a, a.ctypes.data

    def add_ply(self, path, material_id, to_world=None, inv_to_world=None, face_normals=False, radiance=None):
        m, mp = self._vec(to_world, 16)
        mi, mip = self._vec(inv_to_world, 16)
        r, rp = self._vec(radiance, 3)
        _check(self.lib.take_gpu_builder_add_ply(self.h, os.fsencode(path), mp, mip, material_id, int(face_normals), rp))

    def add_mesh(self, positions, indices, material_id, normals=None, uvs=None, compute_missing_normals=True, radiance=None):
        pos = np.ascontiguousarray(positions, np.float64).reshape(-1, 3)
        idx = np.ascontiguousarray(indices, np.int32).reshape(-1, 3)
        n, npn = self._vec(normals, pos.size)
        u, upn = self._vec(uvs, 2 * len(pos))
        r, rp = self._vec(radiance, 3)
        _check(self.lib.take_gpu_builder_add_mesh(self.h, len(pos), pos.ctypes.data, npn, upn, len(idx), idx.ctypes.data, material_id,
                                                  int(compute_missing_normals), rp))

    def add_sphere(self, center, radius, material_id, radiance=None):
        c, cp = self._vec(center, 3)
        r, rp = self._vec(radiance, 3)
        _check(self.lib.take_gpu_builder_add_sphere(self.h, cp, float(radius), material_id, rp))

    def add_point_light(self, intensity, position):
        i, ip = self._vec(intensity, 3)
        p, pp = self._vec(position, 3)
        _check(self.lib.take_gpu_builder_add_point_light(self.h, ip, pp))

    def timings(self) -> dict:
        out = (C.c_double * 4)()
        _check(self.lib.take_gpu_builder_timings(self.h, out))
        return dict(zip(("ms_read", "ms_convert", "ms_normals", "ms_append"), out))

    def desc(self) -> TakeSceneDesc:
        """Geometry and light fields filled (pointers into the builder: keep it alive); the rest zero."""
        d = TakeSceneDesc()
        _check(self.lib.take_gpu_builder_finish(self.h, C.byref(d)))
        return d

    def arrays(self) -> dict:
        from .sceneio import LIGHT_DTYPE
        d = self.desc()

        def arr(ptr, n, dt):
            if n == 0:
                return np.zeros(0, dt)
            return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(n * np.dtype(dt).itemsize,)).view(dt).copy()

        nv, npr = d.num_vertices, d.num_prims
        return dict(positions=arr(d.positions, 3 * nv, np.float64).reshape(-1, 3), normals=arr(d.normals, 3 * nv, np.float64).reshape(-1, 3),
                    uvs=arr(d.uvs, 2 * nv, np.float64).reshape(-1, 2), indices=arr(d.indices, 3 * npr, np.int32).reshape(-1, 3),
                    prim_material=arr(d.prim_material, npr, np.int32), prim_light=arr(d.prim_light, npr, np.int32),
                    prim_flags=arr(d.prim_flags, npr, np.uint8), spheres=arr(d.spheres, 4 * d.num_spheres, np.float64).reshape(-1, 4),
                    lights=arr(d.lights, d.num_lights, LIGHT_DTYPE))

    def close(self):
        if self.h:
            self.lib.take_gpu_builder_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MultiGpuScene:
    """Persistent single-process multi-GPU handle (take_gpu_multi_create / _render / _destroy): the host-side trees are
    built once, one replica per device, one NCCL communicator; each `render_sums` call shards the sample range over the
    devices and returns the reduced sums from devices[0]."""

    def __init__(self, flat: FlatScene, devices):
        self.lib = L = load_library()
        self.flat = flat
        self.devices = list(devices)
        devs = (C.c_int * len(self.devices))(*self.devices)
        self._desc = flat.to_desc()
        self.h = C.c_void_p()
        _check(L.take_gpu_multi_create(len(self.devices), devs, C.byref(self._desc), C.byref(self.h)))

    def render_sums(self, integrator="mis", max_depth=5, spp_begin=0, spp_end=1, seed=0, sumsq=True, out=None):
        H, W = self.flat.height, self.flat.width
        s, s2 = out if out is not None else (np.empty((H, W, 3)), np.empty((H, W, 3)) if sumsq else None)
        o = TakeRenderOpts(INTEGRATORS[integrator], max_depth, spp_begin, spp_end, seed, 0, 0)
        st = TakeStats()
        _check(self.lib.take_gpu_multi_render(self.h, C.byref(o), s.ctypes.data, s2.ctypes.data if s2 is not None else None,
                                              C.byref(st)))
        return s, s2, st.as_dict()

    def close(self):
        if self.h:
            self.lib.take_gpu_multi_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def render_multi(flat: FlatScene, devices, integrator="mis", max_depth=5, spp_begin=0, spp_end=1, seed=0, sumsq=True):
    """take_gpu_render_multi: one process, several GPUs, one NCCL sum-reduce.  Returns (sum, sumsq, stats)."""
    L = load_library()
    L.take_gpu_render_multi.argtypes = [C.c_int, C.c_void_p, C.POINTER(TakeSceneDesc), C.POINTER(TakeRenderOpts), C.c_void_p,
                                        C.c_void_p, C.POINTER(TakeStats)]
    devs = (C.c_int * len(devices))(*devices)
    desc = flat.to_desc()
    s = np.empty((flat.height, flat.width, 3), np.float64)
    s2 = np.empty_like(s) if sumsq else None
    st = TakeStats()
    o = TakeRenderOpts(INTEGRATORS[integrator], max_depth, spp_begin, spp_end, seed, 0, 0)
    _check(L.take_gpu_render_multi(len(devices), devs, C.byref(desc), C.byref(o), s.ctypes.data,
                                   s2.ctypes.data if sumsq else None, C.byref(st)))
    return s, s2, st.as_dict()


def render(params, device: int = 0, integrator: str = "mis", seed: int = 0) -> np.ndarray:
    """Mirror of `Image3 render(const std::vector<std::string>& params)` (src/render.h:5, src/render.cpp:9-87):
    params[0] is the scene (a TAKESCN1 file written by the scene flattener), `-max_depth N` as in render.cpp:14-23
    (default 50).  Returns the H x W x 3 float64 image, row 0 = top (what render.cpp:78 stores)."""
    if len(params) < 1:
        return np.zeros((0, 0, 3))
    max_depth, filename = 50, None
    i = 0
    while i < len(params):
        if params[i] == "-max_depth":
            i += 1
            max_depth = int(params[i])
        elif filename is None:
            filename = params[i]
        i += 1
    flat = FlatScene.load(filename)
    scene = GpuScene(flat, device)
    try:
        s, _, _ = scene.render_sums(integrator, max_depth, 0, flat.spp, seed, sumsq=False)
    finally:
        scene.close()
    return s / float(flat.spp)


# ---- host-only diagnostics -------------------------------------------------------------------------------
REF_NODE_DTYPE = np.dtype([("lo", "<f8", 3), ("hi", "<f8", 3), ("left", "<i4"), ("right", "<i4"), ("prim", "<i4"), ("pad", "<i4")])
FAST_NODE_DTYPE = np.dtype([("c0lox", "<f4"), ("c0hix", "<f4"), ("c0loy", "<f4"), ("c0hiy", "<f4"),
                            ("c1lox", "<f4"), ("c1hix", "<f4"), ("c1loy", "<f4"), ("c1hiy", "<f4"),
                            ("c0loz", "<f4"), ("c0hiz", "<f4"), ("c1loz", "<f4"), ("c1hiz", "<f4"),
                            ("child0", "<i4"), ("child1", "<i4"), ("count0", "<i4"), ("count1", "<i4")])
WIDE_NODE_DTYPE = np.dtype([("lox", "<f4", 4), ("hix", "<f4", 4), ("loy", "<f4", 4), ("hiy", "<f4", 4), ("loz", "<f4", 4),
                            ("hiz", "<f4", 4), ("child", "<i4", 4), ("count", "<i4", 4)])
WIDE_EMPTY = 0x7fffffff
assert REF_NODE_DTYPE.itemsize == 64 and FAST_NODE_DTYPE.itemsize == 64 and WIDE_NODE_DTYPE.itemsize == 128


def write_exr_packed(path, width: int, height: int, packed: np.ndarray, threads: int = 0) -> None:
    """Host half of the output step: deflate the filtered blocks and write the .exr (no CUDA device needed)."""
    packed = np.ascontiguousarray(packed, np.uint8)
    L = load_library()
    assert packed.size == L.take_gpu_exr_packed_size(width, height)
    _check(L.take_gpu_exr_write_packed(os.fsencode(path), width, height, packed.ctypes.data, threads))


def _host_build_arrays(L, h) -> dict:
    L.take_gpu_host_build_info.argtypes = [C.c_void_p, C.c_void_p]
    L.take_gpu_host_build_copy.argtypes = [C.c_void_p] * 6
    info = (C.c_double * 8)()
    _check(L.take_gpu_host_build_info(h, info))
    n_ref, root, n_fast, n_prims, depth = (int(info[i]) for i in range(5))
    ref = np.zeros(n_ref, REF_NODE_DTYPE)
    rank = np.zeros(n_prims, np.int32)
    fast = np.zeros(n_fast, FAST_NODE_DTYPE)
    leaf = np.zeros(n_prims, np.int32)
    recs = np.zeros((n_prims, 12), np.float64)
    _check(L.take_gpu_host_build_copy(h, ref.ctypes.data, rank.ctypes.data, fast.ctypes.data, leaf.ctypes.data,
                                      recs.ctypes.data))
    L.take_gpu_host_build_wide.restype = C.c_int64
    L.take_gpu_host_build_wide.argtypes = [C.c_void_p, C.c_void_p]
    wide = np.zeros(int(L.take_gpu_host_build_wide(h, None)), WIDE_NODE_DTYPE)
    L.take_gpu_host_build_wide(h, wide.ctypes.data)
    return dict(ref_nodes=ref, ref_root=root, dfs_rank=rank, fast_nodes=fast, wide_nodes=wide, leaf_prims=leaf, leaf_records=recs,
                depth=depth, abs_max=float(info[5]), ms_ref=float(info[6]), ms_fast=float(info[7]))


def host_build(flat: FlatScene) -> dict:
    """The acceleration structures take_gpu_scene_create would upload, built on the host only (no CUDA needed)."""
    hb = HostBuild(flat)
    try:
        return hb.arrays()
    finally:
        hb.close()
