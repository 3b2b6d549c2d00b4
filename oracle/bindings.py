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

INTEGRATORS = {"mis": 0, "raw": 1, "one_sample_mis": 2}


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

    def render(self, integrator="mis", max_depth=5, spp_begin=0, spp_end=1, seed=0, threads=8, sumsq=True):
        s = np.zeros((self.height, self.width, 3), np.float64)
        s2 = np.zeros_like(s) if sumsq else None
        rc = self.lib.ref_render(self.h, INTEGRATORS[integrator], max_depth, spp_begin, spp_end, seed, threads,
                                 _ptr(s, np.float64), _ptr(s2, np.float64))
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
                                 C.c_void_p, C.c_void_p]
        L.ref_radiance_samples.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint64, C.c_int64, C.c_void_p,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.ref_rng_selfcheck.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64, C.c_int]

    def load(self, xml_path) -> RefScene:
        return RefScene(self.lib, xml_path)

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
