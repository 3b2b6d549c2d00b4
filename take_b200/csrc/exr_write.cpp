// Host half of the output step: deflate the pre-filtered 16-scanline blocks produced by k_exr_pack (exr_out.cuh) and
// write a scanline OpenEXR file with the attributes the reference's writer emits (src/image.cpp:157-175 -> tinyexr
// SaveEXR: channels B, G, R as HALF, ZIP compression, increasing-Y line order).  The reference deflates all blocks on
// the calling thread with the miniz it vendors; here every block is an independent task on `threads` host threads with
// the system zlib.  Compressed bytes may differ between the two deflate implementations; the decoded pixels do not.
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <zlib.h>

#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "../../include/take_gpu.h"

namespace {

void put_u32(std::vector<uint8_t> &v, uint32_t x) { for (int i = 0; i < 4; ++i) v.push_back((uint8_t)(x >> (8 * i))); }
void put_str(std::vector<uint8_t> &v, const char *s) { v.insert(v.end(), s, s + strlen(s) + 1); }
void put_f32(std::vector<uint8_t> &v, float f) { uint32_t u; memcpy(&u, &f, 4); put_u32(v, u); }

void attr(std::vector<uint8_t> &v, const char *name, const char *type, const std::vector<uint8_t> &data) {
    put_str(v, name);
    put_str(v, type);
    put_u32(v, (uint32_t)data.size());
    v.insert(v.end(), data.begin(), data.end());
}

// Inverse of the ZIP pre-filter (delta predictor, then byte re-interleave): needed only for blocks that deflate cannot
// shrink, which the format stores raw (tinyexr CompressZip, "Issue 40").
void unfilter(const uint8_t *src, size_t n, uint8_t *dst) {
    std::vector<uint8_t> t(src, src + n);
    for (size_t i = 1; i < n; ++i) t[i] = (uint8_t)(t[i - 1] + t[i] - 128);
    const size_t half = (n + 1) / 2;
    for (size_t i = 0; i < n; ++i) dst[i] = (i & 1) ? t[half + (i >> 1)] : t[i >> 1];
}

}  // namespace

extern "C" int64_t take_gpu_exr_packed_size(int32_t width, int32_t height) {
    return width > 0 && height > 0 ? (int64_t)width * height * 6 : 0;
}

extern "C" int take_exr_fail(int code, const std::string &msg);  // take_gpu.cu (sets take_gpu_last_error)

extern "C" int take_gpu_exr_write_packed(const char *path, int32_t width, int32_t height, const uint8_t *packed, int32_t threads) {
    if (!path || !packed || width <= 0 || height <= 0) return take_exr_fail(TAKE_E_INVALID, "take_gpu_exr_write_packed: bad arguments");
    const size_t line_bytes = (size_t)width * 6, block_bytes = 16 * line_bytes;
    const int n_blocks = (height + 15) / 16;
    std::vector<std::vector<uint8_t>> out(n_blocks);
    std::atomic<int> next{0}, bad{0};
    auto work = [&]() {
        for (;;) {
            const int b = next.fetch_add(1);
            if (b >= n_blocks) return;
            const size_t lines = (size_t)std::min(16, height - 16 * b), n = lines * line_bytes;
            const uint8_t *src = packed + (size_t)b * block_bytes;
            uLongf len = compressBound((uLong)n);
            std::vector<uint8_t> &dst = out[b];
            dst.resize(len);
            if (compress2(dst.data(), &len, src, (uLong)n, Z_DEFAULT_COMPRESSION) != Z_OK) { bad = 1; return; }
            if (len >= n) {  // stored raw, un-filtered
                dst.resize(n);
                unfilter(src, n, dst.data());
            } else {
                dst.resize(len);
            }
        }
    };
    const int nt = std::max(1, std::min(threads > 0 ? threads : (int)std::thread::hardware_concurrency(), n_blocks));
    std::vector<std::thread> pool;
    for (int i = 1; i < nt; ++i) pool.emplace_back(work);
    work();
    for (auto &t : pool) t.join();
    if (bad) return take_exr_fail(TAKE_E_INVALID, "take_gpu_exr_write_packed: deflate failed");

    std::vector<uint8_t> h;
    put_u32(h, 20000630u);  // magic 0x76 0x2f 0x31 0x01
    put_u32(h, 2u);         // version 2, single-part scanline
    {
        std::vector<uint8_t> ch;
        for (const char *name : {"B", "G", "R"}) {
            put_str(ch, name);
            put_u32(ch, 1u);  // HALF
            put_u32(ch, 0u);  // pLinear + 3 reserved bytes
            put_u32(ch, 1u);  // xSampling
            put_u32(ch, 1u);  // ySampling
        }
        ch.push_back(0);
        attr(h, "channels", "chlist", ch);
    }
    attr(h, "compression", "compression", {3});  // ZIP, 16 scanlines per block
    {
        std::vector<uint8_t> w;
        put_u32(w, 0); put_u32(w, 0); put_u32(w, (uint32_t)(width - 1)); put_u32(w, (uint32_t)(height - 1));
        attr(h, "dataWindow", "box2i", w);
        attr(h, "displayWindow", "box2i", w);
    }
    attr(h, "lineOrder", "lineOrder", {0});
    { std::vector<uint8_t> f; put_f32(f, 1.0f); attr(h, "pixelAspectRatio", "float", f); }
    { std::vector<uint8_t> f; put_f32(f, 0.0f); put_f32(f, 0.0f); attr(h, "screenWindowCenter", "v2f", f); }
    { std::vector<uint8_t> f; put_f32(f, 1.0f); attr(h, "screenWindowWidth", "float", f); }
    h.push_back(0);

    uint64_t off = h.size() + 8ull * n_blocks;
    std::vector<uint8_t> table;
    for (int b = 0; b < n_blocks; ++b) {
        for (int i = 0; i < 8; ++i) table.push_back((uint8_t)(off >> (8 * i)));
        off += 8 + out[b].size();
    }
    FILE *f = fopen(path, "wb");
    if (!f) return take_exr_fail(TAKE_E_INVALID, std::string("take_gpu_exr_write_packed: cannot open ") + path);
    bool ok = fwrite(h.data(), 1, h.size(), f) == h.size() && fwrite(table.data(), 1, table.size(), f) == table.size();
    for (int b = 0; b < n_blocks && ok; ++b) {
        std::vector<uint8_t> hd;
        put_u32(hd, (uint32_t)(16 * b));
        put_u32(hd, (uint32_t)out[b].size());
        ok = fwrite(hd.data(), 1, 8, f) == 8 && fwrite(out[b].data(), 1, out[b].size(), f) == out[b].size();
    }
    ok = (fclose(f) == 0) && ok;
    return ok ? TAKE_OK : take_exr_fail(TAKE_E_INVALID, std::string("take_gpu_exr_write_packed: short write to ") + path);
}
