// Output step on the device: what the reference does between the render loop and the deflate call of its EXR writer.
//
//   src/render.cpp:78            img(x, H-y-1) = color / Real(spp)      (vector.h:194-197: multiply by 1/spp)
//   src/image.cpp:157-161        double -> float (C cast, round to nearest even)
//   tinyexr.h float_to_half_full float -> half: truncate 13 mantissa bits, +1 if the first dropped bit is set
//                                (round half UP in magnitude -- not IEEE nearest-even), overflow -> inf, gradual underflow
//   tinyexr.h SaveEXR/EncodeChunk channels B, G, R (alphabetical), HALF, planar per scanline, 16-scanline ZIP blocks
//   tinyexr.h CompressZip        byte de-interleave (even bytes first, odd bytes after) + delta predictor (+128)
//
// The kernel writes every block in exactly the form CompressZip hands to deflate, so the host half of the step
// (exr_write.cpp) only runs deflate per block on its threads and writes the file.  One thread per output byte: each
// evaluates the (at most two) half values its byte and its predecessor come from; the work is a rounding error next
// to the render (6 bytes per pixel), so clarity wins over staging through shared memory.
#pragma once
#include <stdint.h>

namespace take {

#define TAKE_EXR_BLOCK_LINES 16

// tinyexr.h float_to_half_full, on the bit pattern.
__host__ __device__ inline uint16_t exr_float_to_half(uint32_t f) {
    const uint32_t sign = f >> 31, fexp = (f >> 23) & 0xffu, fman = f & 0x7fffffu;
    uint32_t o = 0;
    if (fexp == 0) {
        o = 0;  // signed zero / float denormal: underflows to zero
    } else if (fexp == 255) {
        o = (31u << 10) | (fman ? 0x200u : 0u);  // NaN -> qNaN, Inf -> Inf
    } else {
        const int newexp = (int)fexp - 127 + 15;
        if (newexp >= 31) {
            o = 31u << 10;  // overflow: infinity
        } else if (newexp <= 0) {
            if ((14 - newexp) <= 24) {  // half denormal
                const uint32_t mant = fman | 0x800000u;
                o = mant >> (14 - newexp);
                if ((mant >> (13 - newexp)) & 1u) o++;
            }
        } else {
            o = ((uint32_t)newexp << 10) | (fman >> 13);
            if (fman & 0x1000u) o++;  // may carry into the exponent (and up to infinity): intended
        }
    }
    return (uint16_t)((o & 0x7fffu) | (sign << 15));
}

struct ExrGeom {
    int32_t width, height;
    int64_t line_bytes;   // width * 3 channels * 2 bytes
    int64_t block_bytes;  // 16 * line_bytes (the last block may be shorter)
};

#ifdef __CUDACC__
// Raw (un-filtered) byte k of block `b`: scanline-major, per scanline the B, G, R planes of halves.
__device__ __forceinline__ uint32_t exr_raw_byte(const ExrGeom &g, const double *__restrict__ sum, double inv_spp, int64_t line0,
                                                 int64_t k) {
    const int64_t line = line0 + k / g.line_bytes;
    const int64_t in_line = k % g.line_bytes;
    const int32_t plane = (int32_t)(in_line / (2 * (int64_t)g.width));  // 0 = B, 1 = G, 2 = R
    const int64_t x = (in_line % (2 * (int64_t)g.width)) >> 1;
    const double mean = sum[3 * (line * g.width + x) + (2 - plane)] * inv_spp;
    const uint16_t h = exr_float_to_half(__float_as_uint(__double2float_rn(mean)));
    return (k & 1) ? (uint32_t)(h >> 8) : (uint32_t)(h & 0xffu);
}

__global__ void k_exr_pack(ExrGeom g, const double *__restrict__ sum, double inv_spp, uint8_t *__restrict__ packed) {
    const int64_t total = (int64_t)g.height * g.line_bytes;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / g.block_bytes;
        const int64_t p = i - b * g.block_bytes;
        const int64_t line0 = b * TAKE_EXR_BLOCK_LINES;
        const int64_t lines = min((int64_t)TAKE_EXR_BLOCK_LINES, (int64_t)g.height - line0);
        const int64_t n = lines * g.line_bytes, half = (n + 1) / 2;
        // de-interleaved position p holds raw byte 2p (first half) or 2(p-half)+1 (second half)
        const int64_t k = p < half ? 2 * p : 2 * (p - half) + 1;
        const uint32_t cur = exr_raw_byte(g, sum, inv_spp, line0, k);
        uint32_t out = cur;
        if (p > 0) {
            const int64_t q = p - 1;
            const int64_t kq = q < half ? 2 * q : 2 * (q - half) + 1;
            out = cur - exr_raw_byte(g, sum, inv_spp, line0, kq) + 128u;  // tinyexr: d = t[i] - p + (128 + 256)
        }
        packed[i] = (uint8_t)out;
    }
}
#endif

}  // namespace take
