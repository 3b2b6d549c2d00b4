// Device-side math, RNG and scene view shared by all kernels.
//
// Arithmetic contract: every double-precision expression below follows the reference's operation order and the
// whole library is compiled with -fmad=false, so no FMA contraction happens anywhere (the reference's x86-64
// build has none either); + - * / sqrt are IEEE-exact on both sides, which is what makes
// hit distances bit-identical.  FP32 appears only in the conservative box tests of the fast traversal, where
// FMAs are requested explicitly with fmaf().
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/take_gpu.h"
#include "bvh_build.h"

namespace take {

#define TAKE_EPS 1e-7  // c_EPSILON, src/take.h:30
#define TAKE_PI 3.14159265358979323846
#define TAKE_INVPI (1.0 / TAKE_PI)
#define TAKE_TWOPI (2.0 * TAKE_PI)
#define TAKE_INVTWOPI (1.0 / TAKE_TWOPI)

// ---- src/vector.h ---------------------------------------------------------------------------
struct D3 {
    double x, y, z;
};
struct D2 {
    double x, y;
};
__device__ __forceinline__ D3 mk3(double x, double y, double z) { D3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ D3 add(D3 a, D3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ D3 sub(D3 a, D3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ D3 neg(D3 a) { return mk3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ D3 mul(D3 a, double s) { return mk3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ D3 mulv(D3 a, D3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ D3 rsub(double s, D3 a) { return mk3(s - a.x, s - a.y, s - a.z); }            // vector.h:143-146
__device__ __forceinline__ D3 divs(D3 v, double s) { double inv = 1.0 / s; return mk3(v.x * inv, v.y * inv, v.z * inv); }  // :193-197
__device__ __forceinline__ double dot(D3 a, D3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }         // :222-225
__device__ __forceinline__ D3 cross(D3 a, D3 b) {
    return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ double length(D3 v) { return sqrt(dot(v, v)); }
__device__ __forceinline__ D3 normalize(D3 v) {  // :249-257
    double l = length(v);
    if (l <= 0) return mk3(0, 0, 0);
    return divs(v, l);
}
__device__ __forceinline__ D3 to_world(D3 n, D3 v) {  // :314-326 (Frisvad ONB)
    D3 x, y;
    if (n.z < -1 + 1e-6) {
        x = mk3(0, -1, 0);
        y = mk3(-1, 0, 0);
    } else {
        double a = 1 / (1 + n.z);
        double b = -n.x * n.y * a;
        x = mk3(1 - n.x * n.x * a, b, -n.x);
        y = mk3(b, 1 - n.y * n.y * a, -n.y);
    }
    return add(add(mul(x, v.x), mul(y, v.y)), mul(n, v.z));
}
__device__ __forceinline__ double clampd(double v, double lo, double hi) { return (v < lo) ? lo : (hi < v) ? hi : v; }
__device__ __forceinline__ double modulo1(double a) { double r = fmod(a, 1.0); return (r < 0.0) ? r + 1.0 : r; }  // take.h:66-69

// ---- counter-based sample streams -------------------------------------------------------------
// Word j of sample (pixel, sample) = Philox4x32-10(ctr = {j/4, pixel, sample_lo, sample_hi}, key = seed)[j%4];
// the k-th random_real consumes words 2k, 2k+1 the way libstdc++'s uniform_real_distribution<double> consumes
// two mt19937 outputs (src/take.h:89-91), so the reference can be driven with the very same numbers.
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

struct Rng {
    uint64_t seed, sample;
    uint32_t pixel, k;
    // (Keeping the unused half of a Philox block for the next draw was measured: +3 registers and a branch made the
    // shade kernels 6 % slower, and the two draws of a camera ray are folded by the compiler anyway.)
    __device__ __forceinline__ double next() {
        uint32_t w[4];
        philox4x32_10(k >> 1, pixel, (uint32_t)sample, (uint32_t)(sample >> 32), (uint32_t)seed, (uint32_t)(seed >> 32), w);
        const uint32_t w0 = (k & 1) ? w[2] : w[0], w1 = (k & 1) ? w[3] : w[1];
        ++k;
        double sum = (double)w0 + (double)w1 * 4294967296.0;
        double r = sum * 0x1p-64;                  // / 2^64 (exact scaling)
        if (r >= 1.0) r = 0x1.fffffffffffffp-1;    // nextafter(1, 0), <bits/random.tcc> generate_canonical
        return r;
    }
};

// ---- device view of the uploaded scene ----------------------------------------------------------
struct DevTexture {
    int32_t w, h;
    const double *rgb;
};

struct DevScene {
    // fast tree
    const float4 *nodes;    // 4 float4 per FastNode
    const float4 *wide_nodes;  // 8 float4 per WideNode (4-wide collapse of the same tree)
    const double2 *tris;    // 6 double2 per leaf slot: v0.xy | v0.z,idbits | e1.xy | e1.z,aux | e2.xy | e2.z,kind
    // reference-order tree
    const RefNode *ref_nodes;
    int32_t ref_root;
    int32_t fast_depth;
    // by-primitive / by-vertex shading data (FP64, reference layout)
    const double *positions, *normals, *uvs;
    const int32_t *indices, *prim_material, *prim_light, *dfs_rank;
    const uint8_t *prim_flags, *prim_mtype;
    const double *spheres;
    // per-primitive shading records (shading.cuh: ShadeRec layout), built on the device by k_build_shade_recs
    const double *shade_recs;
    int32_t shade_stride;  // doubles per record: 16 (no primitive carries uvs) or 20
    const double *light_recs;  // per-light records (shading.cuh: LightRec layout), built by k_build_light_recs
    const TakeMaterialDesc *materials;
    const TakeLightDesc *lights;
    const DevTexture *textures;
    int32_t num_lights, num_materials;
    const double *light_pmf, *light_cdf;   // power-proportional light pick (TAKE_INTEGRATOR_ONE_SAMPLE_MIS_POWER): pmf [N], cdf [N + 1]
    int64_t num_prims;
    // camera (src/render.cpp:37-44) and background
    int32_t width, height;
    D3 lookfrom, cam_u, cam_v, cam_w;
    double viewport_w, viewport_h;
    D3 background;
    float abs_max;  // max |coordinate| over the scene, for the per-ray conservative padding
    // EXTENSION (include/take_gpu.h): lat-long environment map and its sampling tables
    const double *env_rgb, *env_marg, *env_cond;  // texels [h][w][3]; marginal CDF [h+1]; conditional CDFs [h][w+1]
    double env_total;
    int32_t env_w, env_h;
    // Counts closest-hit leaf tests whose outcome depended on the tie-break rank (equal t with the current hit).  A render that
    // runs before the reference-order tree -- the source of the ranks -- has arrived is valid iff this stays 0 (take_gpu.cu).
    unsigned long long *tie_count;
    int32_t env_light;   // 1: the environment is entry number num_lights of the uniform light pick
    int32_t pick_count;  // num_lights + env_light: the N of sample_light (src/light.cpp:5-7) and of the 1/N in the light pdfs
};

// 256-bit read-only loads (sm_100: LDG.E.256).  The traversal kernels are bound by L1 wavefront throughput -- every
// lane of a divergent warp touches its own cache line, one wavefront per line per load instruction -- so moving a
// 64-byte node in two instructions instead of four halves the wavefronts per node.  Addresses must be 32-byte aligned.
#ifndef TAKE_LDG256
#define TAKE_LDG256 1
#endif
struct F8 { float4 a, b; };
struct D4 { double2 a, b; };
__device__ __forceinline__ F8 ldg_f8(const float4 *p) {
    F8 r;
#if TAKE_LDG256
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(r.a.x), "=f"(r.a.y), "=f"(r.a.z), "=f"(r.a.w), "=f"(r.b.x), "=f"(r.b.y), "=f"(r.b.z), "=f"(r.b.w) : "l"(p));
#else
    r.a = __ldg(p); r.b = __ldg(p + 1);
#endif
    return r;
}
__device__ __forceinline__ D4 ldg_d4(const double2 *p) {
    D4 r;
#if TAKE_LDG256
    asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.a.x), "=d"(r.a.y), "=d"(r.b.x), "=d"(r.b.y) : "l"(p));
#else
    r.a = __ldg(p); r.b = __ldg(p + 1);
#endif
    return r;
}

struct HitOut {
    int32_t prim;
    int32_t rank;
    double t, u, v;
};

}  // namespace take
