"""Minimal scanline OpenEXR reader for the tests (HALF/FLOAT channels, NONE/ZIPS/ZIP compression, single part).
Returns the raw half bit patterns, so two files can be compared bit for bit whatever deflate implementation wrote them."""
import struct
import zlib

import numpy as np


def _cstr(b, p):
    e = b.index(b"\0", p)
    return b[p:e].decode(), e + 1


def read_exr(path):
    b = open(path, "rb").read()
    magic, version = struct.unpack_from("<II", b, 0)
    assert magic == 20000630, "not an OpenEXR file"
    assert version & 0xff == 2 and not (version & 0x200), "single-part scanline files only"
    p, attrs = 8, {}
    while b[p] != 0:
        name, p = _cstr(b, p)
        typ, p = _cstr(b, p)
        (size,) = struct.unpack_from("<i", b, p)
        p += 4
        attrs[name] = (typ, b[p:p + size])
        p += size
    p += 1
    channels, q, ch = [], 0, attrs["channels"][1]
    while ch[q] != 0:
        name, q = _cstr(ch, q)
        ptype, _plin, xs, ys = struct.unpack_from("<iIii", ch, q)
        q += 16
        assert xs == 1 and ys == 1
        channels.append((name, ptype))
    comp = attrs["compression"][1][0]
    lines_per_block = {0: 1, 2: 1, 3: 16}[comp]
    x0, y0, x1, y1 = struct.unpack("<iiii", attrs["dataWindow"][1])
    W, H = x1 - x0 + 1, y1 - y0 + 1
    n_blocks = (H + lines_per_block - 1) // lines_per_block
    offsets = struct.unpack_from("<%dQ" % n_blocks, b, p)
    bpp = {0: 4, 1: 2, 2: 4}
    line_bytes = sum(bpp[t] for _, t in channels) * W
    out = {name: np.zeros((H, W), np.uint16 if t == 1 else np.uint32) for name, t in channels}
    for off in offsets:
        y, size = struct.unpack_from("<ii", b, off)
        data = b[off + 8: off + 8 + size]
        lines = min(lines_per_block, y1 + 1 - y)
        n = lines * line_bytes
        if comp != 0 and size != n:
            t = np.frombuffer(zlib.decompress(data), np.uint8).astype(np.int64)
            assert t.size == n
            t[1:] -= 128
            t = (np.cumsum(t) & 0xff).astype(np.uint8)          # undo the delta predictor
            raw = np.empty(n, np.uint8)
            half = (n + 1) // 2
            raw[0::2] = t[:half]                                  # undo the byte de-interleave
            raw[1::2] = t[half:]
        else:
            raw = np.frombuffer(data, np.uint8)
        q = 0
        for l in range(lines):
            for name, t in channels:
                nb = bpp[t] * W
                out[name][y - y0 + l] = raw[q:q + nb].view("<u2" if t == 1 else "<u4")
                q += nb
    return {"width": W, "height": H, "channels": channels, "compression": comp, "planes": out, "attrs": attrs}


def unpack_blocks(packed, width, height):
    """Invert the ZIP pre-filter of a take_gpu_exr_pack buffer: -> (H, W, 3) half bit patterns in R, G, B order."""
    packed = np.asarray(packed, np.uint8)
    line_bytes = width * 6
    img = np.zeros((height, width, 3), np.uint16)
    for y0 in range(0, height, 16):
        lines = min(16, height - y0)
        n = lines * line_bytes
        t = packed[y0 * line_bytes: y0 * line_bytes + n].astype(np.int64)
        t[1:] -= 128
        t = (np.cumsum(t) & 0xff).astype(np.uint8)
        raw = np.empty(n, np.uint8)
        half = (n + 1) // 2
        raw[0::2] = t[:half]
        raw[1::2] = t[half:]
        planes = raw.view("<u2").reshape(lines, 3, width)        # per line: B, G, R
        img[y0:y0 + lines, :, 2] = planes[:, 0]
        img[y0:y0 + lines, :, 1] = planes[:, 1]
        img[y0:y0 + lines, :, 0] = planes[:, 2]
    return img
