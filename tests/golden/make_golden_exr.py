"""Generate tests/golden/exr/output_step.npz from the UNMODIFIED reference (oracle/_ref/libtake_ref.so):
the half bit patterns that the reference's own imwrite (src/image.cpp:135-175 -> vendored tinyexr, float_to_half_full)
stores for a set of per-pixel sums, read back from the files it wrote.  Run in the build container:
    python tests/golden/make_golden_exr.py
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from exr_reader import read_exr  # noqa: E402
from oracle import bindings as ob  # noqa: E402


def cases():
    """name -> (sum_rgb (H, W, 3) float64, spp).  Sizes exercise a ragged last block, odd widths and 1-pixel images;
    values exercise rounding ties, half denormals, overflow to infinity, negative values, zeros, inf and NaN."""
    rng = np.random.default_rng(20261018)
    out = {}
    out["radiance_37x21_spp7"] = (rng.gamma(0.7, 2.0, (21, 37, 3)) * 7, 7)
    out["wide_range_33x16_spp3"] = (np.exp(rng.uniform(-30, 14, (16, 33, 3))) * rng.choice([-1.0, 1.0], (16, 33, 3)), 3)
    # every half bit pattern's neighbourhood: exact halves, midpoints between halves (the tie rule), and just off them
    h = np.arange(0, 0x7c00, 7, dtype=np.uint16).view(np.float16).astype(np.float64)
    nxt = (np.arange(0, 0x7c00, 7, dtype=np.uint16) + 1).view(np.float16).astype(np.float64)
    mid = 0.5 * (h + np.where(np.isfinite(nxt), nxt, 65520.0))
    vals = np.concatenate([h, mid, np.nextafter(mid, 0), np.nextafter(mid, np.inf), -mid])
    vals = np.concatenate([vals, [0.0, -0.0, np.inf, -np.inf, np.nan, 65504.0, 65519.99, 65520.0, 1e-8, 2.9e-8, 3.0e-8, 5.96e-8, 1e300, -1e300]])
    n = (vals.size + 2) // 3 * 3
    vals = np.concatenate([vals, np.zeros(n - vals.size)])
    W = 51
    rows = (n // 3 + W - 1) // W
    img = np.zeros(rows * W * 3)
    img[:n] = vals
    out["half_grid_spp1"] = (img.reshape(rows, W, 3), 1)
    out["single_pixel_spp5"] = (np.array([[[1.0, 2.5, 1e-3]]]) * 5, 5)
    # uniformly random half patterns: deflate cannot shrink these blocks, so the writer stores them raw
    bits = rng.integers(0, 0x7c00, (19, 29, 3)).astype(np.uint16)
    out["incompressible_29x19_spp1"] = (bits.view(np.float16).astype(np.float64), 1)
    return out


def main():
    R = ob.RefLib()
    data = {}
    with tempfile.TemporaryDirectory() as td:
        for name, (s, spp) in cases().items():
            mean = s * (1.0 / spp)                      # src/render.cpp:78 via vector.h:194-197
            path = os.path.join(td, name + ".exr")
            with np.errstate(all="ignore"):
                R.imwrite(path, mean)
            e = read_exr(path)
            assert [c for c, _ in e["channels"]] == ["B", "G", "R"] and all(t == 1 for _, t in e["channels"])
            data[name + "/sum"] = s
            data[name + "/spp"] = np.int64(spp)
            data[name + "/half_rgb"] = np.stack([e["planes"]["R"], e["planes"]["G"], e["planes"]["B"]], axis=-1)
            data[name + "/compression"] = np.int64(e["compression"])
    np.savez_compressed(os.path.join(HERE, "exr", "output_step.npz"), **data)
    print("wrote", len(data) // 4, "cases")


if __name__ == "__main__":
    main()
