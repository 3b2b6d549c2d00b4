"""The output step: per-pixel sums -> image.exr, replacing `color / spp` (src/render.cpp:78) and the
.exr branch of imwrite (src/image.cpp:157-175 -> vendored tinyexr: float_to_half_full, B/G/R planes, ZIP blocks).

Bar: the half bit patterns a reader gets out of our file are IDENTICAL to those out of the file the reference writes
for the same sums (the compressed bytes may differ: the reference deflates with miniz, we with zlib).  Checked three
ways: the CPU restatement against golden vectors produced by the unmodified reference, the host writer against the
reference's own reader, and (gpu) the device kernel against the restatement byte for byte."""
import os

import numpy as np
import pytest

from exr_reader import read_exr, unpack_blocks
from take_b200 import api, scenes

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "exr", "output_step.npz")


def golden_cases():
    g = np.load(GOLDEN)
    names = sorted({k.split("/")[0] for k in g.files})
    return g, names


G, NAMES = golden_cases()


def same_halves(a, b):
    """Equal bit patterns, except that any NaN matches any NaN with the same sign handling as tinyexr (qNaN 0x7e00)."""
    return np.array_equal(a, b)


@pytest.mark.parametrize("name", NAMES)
def test_restatement_matches_reference_golden(oracle_lib, name):
    s, spp = G[name + "/sum"], int(G[name + "/spp"])
    with np.errstate(all="ignore"):
        packed = oracle_lib.exr_pack(s, spp)
    got = unpack_blocks(packed, s.shape[1], s.shape[0])
    assert same_halves(got, G[name + "/half_rgb"])


@pytest.mark.parametrize("name", NAMES)
def test_restatement_matches_reference_live(oracle_lib, ref_lib, name, tmp_path):
    s, spp = G[name + "/sum"], int(G[name + "/spp"])
    p = str(tmp_path / "ref.exr")
    with np.errstate(all="ignore"):
        ref_lib.imwrite(p, s * (1.0 / spp))
        packed = oracle_lib.exr_pack(s, spp)
    e = read_exr(p)
    ref = np.stack([e["planes"]["R"], e["planes"]["G"], e["planes"]["B"]], axis=-1)
    assert same_halves(unpack_blocks(packed, s.shape[1], s.shape[0]), ref)


@pytest.mark.parametrize("name", NAMES)
def test_host_writer_file_decodes_to_the_same_halves(oracle_lib, gpu_lib, name, tmp_path):
    s, spp = G[name + "/sum"], int(G[name + "/spp"])
    H, W = s.shape[:2]
    with np.errstate(all="ignore"):
        packed = oracle_lib.exr_pack(s, spp)
    p = str(tmp_path / "ours.exr")
    api.write_exr_packed(p, W, H, packed, threads=3)
    e = read_exr(p)
    assert [c for c, _ in e["channels"]] == ["B", "G", "R"] and e["compression"] == 3 and (e["width"], e["height"]) == (W, H)
    got = np.stack([e["planes"]["R"], e["planes"]["G"], e["planes"]["B"]], axis=-1)
    assert same_halves(got, G[name + "/half_rgb"])
    # thread count must not change the file
    p1 = str(tmp_path / "ours1.exr")
    api.write_exr_packed(p1, W, H, packed, threads=1)
    assert open(p, "rb").read() == open(p1, "rb").read()


def test_incompressible_blocks_are_stored_raw(oracle_lib, gpu_lib, tmp_path):
    s = G["incompressible_29x19_spp1/sum"]
    p = str(tmp_path / "raw.exr")
    api.write_exr_packed(p, s.shape[1], s.shape[0], oracle_lib.exr_pack(s, 1))
    size = os.path.getsize(p)
    assert size <= s.shape[0] * s.shape[1] * 6 + 512          # never larger than raw + header/table
    e = read_exr(p)
    assert same_halves(np.stack([e["planes"][c] for c in "RGB"], axis=-1), G["incompressible_29x19_spp1/half_rgb"])


@pytest.mark.parametrize("name", [n for n in NAMES if "half_grid" not in n and "wide_range" not in n])
def test_reference_reader_accepts_our_file(oracle_lib, ref_lib, gpu_lib, name, tmp_path):
    s, spp = G[name + "/sum"], int(G[name + "/spp"])
    p = str(tmp_path / "ours.exr")
    api.write_exr_packed(p, s.shape[1], s.shape[0], oracle_lib.exr_pack(s, spp))
    img = ref_lib.imread3(p)                                    # the reference's own imread3 (tinyexr LoadEXR)
    want = G[name + "/half_rgb"].view(np.float16).astype(np.float64)
    assert img.shape == want.shape
    assert np.array_equal(img, want)


def test_writer_rejects_bad_arguments(gpu_lib, tmp_path):
    with pytest.raises(api.TakeGpuError):
        api._check(gpu_lib.take_gpu_exr_write_packed(None, 4, 4, None, 0))
    buf = np.zeros(4 * 4 * 6, np.uint8)
    with pytest.raises(api.TakeGpuError):
        api._check(gpu_lib.take_gpu_exr_write_packed(os.fsencode(str(tmp_path / "no_such_dir" / "x.exr")), 4, 4, buf.ctypes.data, 0))
    assert gpu_lib.take_gpu_exr_packed_size(0, 5) == 0


# ---- GPU -----------------------------------------------------------------------------------------------------------
def _scene_of_size(W, H):
    return api.GpuScene(scenes.cornell_box(W, H, 1).flat())


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_device_pack_equals_restatement(oracle_lib, name):
    s, spp = G[name + "/sum"], int(G[name + "/spp"])
    gs = _scene_of_size(s.shape[1], s.shape[0])
    with np.errstate(all="ignore"):
        want = oracle_lib.exr_pack(s, spp)
    got = gs.exr_pack(s, spp)
    gs.close()
    assert np.array_equal(got, want)                            # byte for byte: the deflate input is identical
    assert same_halves(unpack_blocks(got, s.shape[1], s.shape[0]), G[name + "/half_rgb"])


@pytest.mark.gpu
def test_render_to_exr_equals_render_then_reference_chain(oracle_lib, tmp_path):
    b = scenes.cornell_box(96, 70, 4)
    gs = api.GpuScene(b.flat())
    s, _, _ = gs.render_sums("mis", 5, 0, 4, seed=3)
    p = str(tmp_path / "image.exr")
    st = gs.render_to_exr(p, "mis", 5, 0, 4, seed=3)
    gs.close()
    assert st["samples"] == 96 * 70 * 4
    e = read_exr(p)
    got = np.stack([e["planes"][c] for c in "RGB"], axis=-1)
    want = unpack_blocks(oracle_lib.exr_pack(s, 4), 96, 70)     # restatement of the reference chain on the same sums
    assert same_halves(got, want)
    assert np.isfinite(got.view(np.float16).astype(np.float64)).all() and got.any()


@pytest.mark.gpu
def test_full_size_pack_round_trip(oracle_lib):
    """1920x1080 (BASELINE configs 2/4): size-independent property -- decoding the packed blocks gives back
    half(float(sum/spp)) for every pixel, computed independently with numpy where numpy's rounding agrees (non-ties)."""
    W, H, spp = 1920, 1080, 256
    rng = np.random.default_rng(5)
    s = rng.gamma(0.6, 3.0, (H, W, 3)) * spp
    gs = _scene_of_size(W, H)
    got = unpack_blocks(gs.exr_pack(s, spp), W, H)
    gs.close()
    want = unpack_blocks(oracle_lib.exr_pack(s, spp), W, H)
    assert same_halves(got, want)
    f = (s * (1.0 / spp)).astype(np.float32)
    near = f.astype(np.float16).view(np.uint16)                 # IEEE nearest-even: differs from tinyexr only on exact ties
    assert (np.abs(near.astype(np.int32) - got.astype(np.int32)) <= 1).all()
    assert (near != got).mean() < 1e-3
