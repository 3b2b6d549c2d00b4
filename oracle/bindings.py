"""TEST INFRASTRUCTURE -- ctypes bindings for the two CPU checkers.

  RefLib     oracle/_ref/libtake_ref.so   the unmodified reference + C-ABI harness (oracle/ref_harness.cpp)
  OracleLib  oracle/libtake_oracle.so     our CPU restatement of the hot path     (oracle/take_oracle.cpp)

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module; the
product package (take_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libtake_ref.so")
REF_CLI = os.path.join(HERE, "_ref", "take_ref")
ORACLE_SO = os.path.join(HERE, "libtake_oracle.so")

INTEGRATORS = {"mis": 0, "raw": 1, "one_sample_mis": 2, "one_sample_mis_power": 3}


def build(ref: bool = True):
    """(Re)build the checkers.  The reference library can only be built where /root/reference exists."""
    targets = ["port"] + (["ref"] if ref and os.path.isdir(os.environ.get("TAKE_REF", "/root/reference")) else [])
    subprocess.run(["make", "-s", "-C", HERE, "-j8", *targets], check=True)


def have_ref() -> bool:
    return os.path.exists(REF_SO)


def _ptr(a, dtype):
    if a is None:
        return None
    assert a.dtype == dtype and a.flags["C_CONTIGUOUS"], (a.dtype, dtype)
    return a.ctypes.data_as(C.c_void_p)


def _rays(rays):
    rays = np.ascontiguousarray(rays, dtype=np.float64)
    assert rays.ndim == 2 and rays.shape[1] == 8
    return rays


class RefScene:
    """A scene parsed by the reference's own front end (parse_scene + build_bvh)."""

    def __init__(self, lib, xml_path):
        self.lib = lib
        self.h = lib.ref_scene_load(os.path.abspath(xml_path).encode())
        if not self.h:
            raise RuntimeError("reference parse_scene failed: " + lib.ref_last_error().decode())
        info = (C.c_int64 * 8)()
        lib.ref_scene_info(self.h, info)
        (self.num_prims, self.num_nodes, self.root, self.num_lights, self.width, self.height, self.spp,
         self.num_materials) = (int(v) for v in info)

    def close(self):
        if self.h:
            self.lib.ref_scene_free(self.h)
            self.h = None

    def dump(self, path):
        if self.lib.ref_scene_dump(self.h, os.fspath(path).encode()) != 0:
            raise RuntimeError(self.lib.ref_last_error().decode())

    def bvh(self):
        box = np.empty((self.num_nodes, 6), np.float64)
        links = np.empty((self.num_nodes, 3), np.int32)
        self.lib.ref_bvh_dump(self.h, _ptr(box, np.float64), _ptr(links, np.int32))
        return box, links

    def intersect(self, rays, records=False, threads=8):
        rays = _rays(rays)
        n = len(rays)
        prim = np.empty(n, np.int32)
        t = np.empty(n, np.float64)
        rec = np.empty((n, 16), np.float64) if records else None
        bad = self.lib.ref_intersect(self.h, _ptr(rays, np.float64), n, _ptr(prim, np.int32), _ptr(t, np.float64),
                                     _ptr(rec, np.float64), threads)
        if bad:
            raise AssertionError(f"primitive-id shim disagrees with scene_intersect on {bad} rays")
        return (prim, t, rec) if records else (prim, t)

    def occluded(self, rays, threads=8):
        rays = _rays(rays)
        occ = np.empty(len(rays), np.uint8)
        self.lib.ref_occluded(self.h, _ptr(rays, np.float64), len(rays), _ptr(occ, np.uint8), threads)
        return occ

    def render(self, integrator="mis", max_depth=5, spp_begin=0, spp_end=1, seed=0, threads=8, sumsq=True,
               row_begin=0, row_step=1):
        s = np.zeros((self.height, self.width, 3), np.float64)
        s2 = np.zeros_like(s) if sumsq else None
        rc = self.lib.ref_render(self.h, INTEGRATORS[integrator], max_depth, spp_begin, spp_end, seed, threads,
                                 _ptr(s, np.float64), _ptr(s2, np.float64), row_begin, row_step)
        if rc != 0:
            raise RuntimeError(self.lib.ref_last_error().decode())
        return s, s2

    def radiance_samples(self, px, py, s, integrator="mis", max_depth=5, seed=0, threads=8):
        px = np.ascontiguousarray(px, np.int32); py = np.ascontiguousarray(py, np.int32)
        s = np.ascontiguousarray(s, np.int64)
        out = np.empty((len(px), 3), np.float64)
        self.lib.ref_radiance_samples(self.h, INTEGRATORS[integrator], max_depth, seed, len(px), _ptr(px, np.int32),
                                      _ptr(py, np.int32), _ptr(s, np.int64), _ptr(out, np.float64), threads)
        return out


class RefLib:
    def __init__(self, path=REF_SO):
        L = self.lib = C.CDLL(path)
        L.ref_last_error.restype = C.c_char_p
        L.ref_scene_load.restype = C.c_void_p
        L.ref_scene_load.argtypes = [C.c_char_p]
        L.ref_scene_free.argtypes = [C.c_void_p]
        L.ref_scene_info.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_scene_dump.argtypes = [C.c_void_p, C.c_char_p]
        L.ref_bvh_dump.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_intersect.restype = C.c_int64
        L.ref_intersect.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.ref_occluded.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int]
        L.ref_render.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_uint64, C.c_int,
                                 C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.ref_radiance_samples.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint64, C.c_int64, C.c_void_p,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.ref_rng_selfcheck.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64, C.c_int]
        L.ref_imwrite.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_void_p]
        L.ref_imread3.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]

    def load(self, xml_path) -> RefScene:
        return RefScene(self.lib, xml_path)

    def imwrite(self, path, mean_rgb):
        """The reference's own imwrite (src/image.cpp:135-175) on an (H, W, 3) image of doubles."""
        a = np.ascontiguousarray(mean_rgb, np.float64)
        if self.lib.ref_imwrite(os.fsencode(path), a.shape[1], a.shape[0], _ptr(a, np.float64)) != 0:
            raise RuntimeError("reference imwrite failed: " + self.lib.ref_last_error().decode())

    def imread3(self, path, cap_pixels=1 << 24):
        """The reference's own imread3 (src/image.cpp:80-133)."""
        w, h = C.c_int(0), C.c_int(0)
        out = np.empty(cap_pixels * 3, np.float64)
        rc = self.lib.ref_imread3(os.fsencode(path), C.byref(w), C.byref(h), _ptr(out, np.float64), out.size)
        if rc != 0:
            raise RuntimeError("reference imread3 failed: " + self.lib.ref_last_error().decode())
        return out[: w.value * h.value * 3].reshape(h.value, w.value, 3).copy()

    def rng_selfcheck(self, seed, pixel, sample, n=300) -> int:
        return self.lib.ref_rng_selfcheck(seed, pixel, sample, n)


def run_stock_cli(xml_path, max_depth=5, threads=None, timeout=3600):
    """Time the reference's own renderer (`take <xml> -t N -max_depth D`, src/main.cpp + src/render.cpp) and
    return (render_seconds, bvh_seconds, parse_seconds, threads).  It writes ./image.exr, so run in the scene dir."""
    import re
    threads = threads or os.cpu_count()
    xml_path = os.path.abspath(xml_path)
    out = subprocess.run([REF_CLI, xml_path, "-t", str(threads), "-max_depth", str(max_depth)],
                         cwd=os.path.dirname(xml_path), capture_output=True, text=True, timeout=timeout, check=True).stdout
    g = lambda pat: float(re.search(pat, out).group(1))
    return (g(r"Finish building rendering\. Took ([0-9.eE+-]+) seconds"), g(r"Finish building BVH\. Took ([0-9.eE+-]+) seconds"),
            g(r"Scene parsing done\. Took ([0-9.eE+-]+) seconds"), threads)


class OracleScene:
    """A FlatScene loaded into the CPU restatement (oracle/take_oracle.cpp)."""

    def __init__(self, lib, flat):
        self.lib, self.flat = lib, flat
        self._desc = flat.to_desc()
        self.h = lib.oracle_scene_create(C.byref(self._desc))
        self.width, self.height = flat.width, flat.height
        self.num_prims = flat.num_prims

    def close(self):
        if self.h:
            self.lib.oracle_scene_free(self.h)
            self.h = None

    def set_russian_roulette(self, rr_start: int):
        """EXTENSION (no reference counterpart): Russian roulette from loop iteration rr_start on; 0 = off."""
        self.lib.oracle_set_russian_roulette.argtypes = [C.c_void_p, C.c_int]
        self.lib.oracle_set_russian_roulette(self.h, int(rr_start))

    def bvh(self):
        n = self.lib.oracle_bvh_size(self.h)
        box = np.empty((n, 6), np.float64)
        links = np.empty((n, 3), np.int32)
        self.lib.oracle_bvh_dump(self.h, _ptr(box, np.float64), _ptr(links, np.int32))
        return box, links, self.lib.oracle_bvh_root(self.h)

    def dfs_rank(self):
        rank = np.full(self.num_prims, -1, np.int32)
        self.lib.oracle_dfs_rank(self.h, _ptr(rank, np.int32))
        return rank

    def intersect(self, rays, records=False, counters=False, threads=8):
        rays = _rays(rays)
        n = len(rays)
        prim = np.empty(n, np.int32)
        t = np.empty(n, np.float64)
        uv = np.empty((n, 2), np.float64)
        rec = np.empty((n, 16), np.float64) if records else None
        cnt = np.zeros(2, np.int64)
        self.lib.oracle_intersect(self.h, _ptr(rays, np.float64), n, _ptr(prim, np.int32), _ptr(t, np.float64),
                                  _ptr(uv, np.float64), _ptr(rec, np.float64), _ptr(cnt, np.int64), threads)
        out = (prim, t, uv) + ((rec,) if records else ()) + ((cnt,) if counters else ())
        return out

    def occluded(self, rays, threads=8):
        rays = _rays(rays)
        occ = np.empty(len(rays), np.uint8)
        self.lib.oracle_occluded(self.h, _ptr(rays, np.float64), len(rays), _ptr(occ, np.uint8), threads)
        return occ

    def render(self, integrator="mis", max_depth=5, spp_begin=0, spp_end=1, seed=0, threads=8, sumsq=True,
               stats=False, row_begin=0, row_step=1):
        s = np.zeros((self.height, self.width, 3), np.float64)
        s2 = np.zeros_like(s) if sumsq else None
        st = np.zeros(6, np.int64)
        rc = self.lib.oracle_render(self.h, INTEGRATORS[integrator], max_depth, spp_begin, spp_end, seed, threads,
                                    _ptr(s, np.float64), _ptr(s2, np.float64), _ptr(st, np.int64), row_begin, row_step)
        assert rc == 0
        return (s, s2, st) if stats else (s, s2)

    def radiance_samples(self, px, py, s, integrator="mis", max_depth=5, seed=0, threads=8):
        px = np.ascontiguousarray(px, np.int32); py = np.ascontiguousarray(py, np.int32)
        s = np.ascontiguousarray(s, np.int64)
        out = np.empty((len(px), 3), np.float64)
        rc = self.lib.oracle_radiance_samples(self.h, INTEGRATORS[integrator], max_depth, seed, len(px),
                                              _ptr(px, np.int32), _ptr(py, np.int32), _ptr(s, np.int64),
                                              _ptr(out, np.float64), threads)
        assert rc == 0
        return out

    def env_sample(self, u1, u2):
        out = np.empty(4, np.float64)
        self.lib.oracle_env_sample(self.h, float(u1), float(u2), _ptr(out, np.float64))
        return out[:3].copy(), float(out[3])

    def env_eval(self, direction):
        d = np.ascontiguousarray(direction, np.float64)
        out = np.empty(4, np.float64)
        self.lib.oracle_env_eval(self.h, _ptr(d, np.float64), _ptr(out, np.float64))
        return out[:3].copy(), float(out[3])

    def primary_rays(self, px, py, s=None, seed=0, jitter=True):
        px = np.ascontiguousarray(px, np.int32); py = np.ascontiguousarray(py, np.int32)
        s = np.zeros(len(px), np.int64) if s is None else np.ascontiguousarray(s, np.int64)
        rays = np.empty((len(px), 8), np.float64)
        self.lib.oracle_primary_rays(self.h, seed, len(px), _ptr(px, np.int32), _ptr(py, np.int32),
                                     _ptr(s, np.int64), int(jitter), _ptr(rays, np.float64))
        return rays


class OracleLib:
    def __init__(self, path=ORACLE_SO):
        L = self.lib = C.CDLL(path)
        vp, i64, u64, i32, u32 = C.c_void_p, C.c_int64, C.c_uint64, C.c_int, C.c_uint32
        L.oracle_scene_create.restype = vp
        L.oracle_scene_create.argtypes = [vp]
        L.oracle_scene_free.argtypes = [vp]
        L.oracle_bvh_size.restype = i64
        L.oracle_bvh_size.argtypes = [vp]
        L.oracle_bvh_root.argtypes = [vp]
        L.oracle_bvh_dump.argtypes = [vp, vp, vp]
        L.oracle_dfs_rank.argtypes = [vp, vp]
        L.oracle_intersect.argtypes = [vp, vp, i64, vp, vp, vp, vp, vp, i32]
        L.oracle_occluded.argtypes = [vp, vp, i64, vp, i32]
        L.oracle_render.argtypes = [vp, i32, i32, i64, i64, u64, i32, vp, vp, vp, i32, i32]
        L.oracle_radiance_samples.argtypes = [vp, i32, i32, u64, i64, vp, vp, vp, vp, i32]
        L.oracle_primary_rays.argtypes = [vp, u64, i64, vp, vp, vp, i32, vp]
        L.oracle_stream_real.restype = C.c_double
        L.oracle_stream_real.argtypes = [u64, u32, u64, u32]
        L.oracle_philox.argtypes = [vp, vp, vp]
        L.oracle_env_sample.argtypes = [vp, C.c_double, C.c_double, vp]
        L.oracle_env_eval.argtypes = [vp, vp, vp]
        L.oracle_exr_pack.argtypes = [i32, i32, vp, i64, vp]

    def load(self, flat) -> OracleScene:
        return OracleScene(self.lib, flat)

    def exr_pack(self, sum_rgb, spp):
        """CPU restatement of the output step up to the deflate input (checker for take_gpu_exr_pack*)."""
        a = np.ascontiguousarray(sum_rgb, np.float64)
        out = np.empty(a.shape[0] * a.shape[1] * 6, np.uint8)
        self.lib.oracle_exr_pack(a.shape[1], a.shape[0], _ptr(a, np.float64), spp, _ptr(out, np.uint8))
        return out

    def philox(self, ctr, key):
        ctr = np.ascontiguousarray(ctr, np.uint32); key = np.ascontiguousarray(key, np.uint32)
        out = np.empty(4, np.uint32)
        self.lib.oracle_philox(_ptr(ctr, np.uint32), _ptr(key, np.uint32), _ptr(out, np.uint32))
        return out

    def stream_real(self, seed, pixel, sample, k) -> float:
        return self.lib.oracle_stream_real(seed, pixel, sample, k)


def secondary_rays(rays, t, prim, seed=0):
    """Random secondary rays leaving the hit points of `rays` (for hits only): uniform directions on the sphere,
    tmin = 1e-7, tmax = inf -- the shape of ray the integrators spawn (path_tracing.h:79)."""
    rng = np.random.default_rng(seed)
    hit = prim >= 0
    o = rays[hit, 0:3] + rays[hit, 3:6] * t[hit, None]
    d = rng.normal(size=o.shape)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    out = np.empty((len(o), 8))
    out[:, 0:3], out[:, 3:6], out[:, 6], out[:, 7] = o, d, 1e-7, np.inf
    return out
