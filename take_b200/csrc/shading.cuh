// Hit reconstruction, textures, BSDFs and light sampling on the device, in FP64 with the reference's operation
// order (see device_common.cuh for the arithmetic contract).  Each function cites the reference lines it replaces.
#pragma once
#include "device_common.cuh"

namespace take {

struct Isect {  // src/intersection.h:4-12
    D3 pos, gn, sn;
    D2 uv;
    int32_t material, light;
};

__device__ __forceinline__ D3 ld3(const double *p) { return mk3(p[0], p[1], p[2]); }

__device__ __forceinline__ D2 sphere_uv(D3 p) {  // src/shape.cpp:3-11
    double theta = acos(-p.y);
    double phi = atan2(-p.z, p.x) + TAKE_PI;
    D2 r; r.x = phi / (2 * TAKE_PI); r.y = -theta / TAKE_PI;
    return r;
}

// Per-primitive shading record (ShadeRec): everything fill_isect needs about a primitive in ONE aligned block, so a
// shaded vertex costs four or five full 32-byte sectors issued together instead of a dependent chain
// prim -> {material, light, flags, indices} -> {3 positions, 3 normals, 3 uvs} of ~16 partially used sectors.
// Layout in doubles (stride DevScene::shade_stride = 16, or 20 when some primitive carries uvs):
//   [0..2]   triangle: normalize(cross(p1 - p0, p2 - p0)), the geometric normal before it is flipped towards the ray
//            (shape.cpp:84-88; the same operations on the same operands, done once per primitive instead of once per hit)
//            sphere:   centre, [3] = radius
//   [3..11]  triangle: the three vertex normals (when TAKE_PRIM_HAS_NORMALS)
//   [12]     bits: material id | light id << 32        [13] bits: primitive flags
//   [14..19] triangle: the three vertex uvs (when TAKE_PRIM_HAS_UVS; stride 20 only)
// The records are derived on the device from the uploaded reference-layout arrays (k_build_shade_recs), with the
// device's own -fmad=false FP64 arithmetic, so every value is the one fill_isect used to compute per hit.
#ifndef TAKE_SHADE_RECS
#define TAKE_SHADE_RECS 1
#endif

// src/shape.cpp:30-41 (sphere) and :80-108 (triangle), from the reference-layout arrays
__device__ __forceinline__ void fill_isect_arrays(const DevScene &sc, D3 o, D3 d, int prim, double t, double u, double v, Isect &out) {
    out.pos = add(o, mul(d, t));
    out.material = sc.prim_material[prim];
    out.light = sc.prim_light[prim];
    const uint8_t flags = sc.prim_flags[prim];
    const int32_t *id = sc.indices + 3 * (int64_t)prim;
    if (flags & TAKE_PRIM_SPHERE) {
        const double *s = sc.spheres + 4 * (int64_t)id[0];
        D3 gn = normalize(sub(out.pos, mk3(s[0], s[1], s[2])));
        out.gn = dot(d, gn) < 0 ? gn : neg(gn);
        out.sn = out.gn;
        out.uv = sphere_uv(out.gn);
        return;
    }
    const int64_t i0 = id[0], i1 = id[1], i2 = id[2];
    D3 v0 = ld3(sc.positions + 3 * i0);
    D3 e1 = sub(ld3(sc.positions + 3 * i1), v0), e2 = sub(ld3(sc.positions + 3 * i2), v0);
    D3 gn = normalize(cross(e1, e2));
    out.gn = dot(d, gn) < 0 ? gn : neg(gn);
    const double w = 1 - u - v;
    if (!(flags & TAKE_PRIM_HAS_UVS)) {
        out.uv.x = u; out.uv.y = v;
    } else {
        const double *a = sc.uvs + 2 * i0, *b = sc.uvs + 2 * i1, *c = sc.uvs + 2 * i2;
        out.uv.x = w * a[0] + u * b[0] + v * c[0];
        out.uv.y = w * a[1] + u * b[1] + v * c[1];
    }
    if (!(flags & TAKE_PRIM_HAS_NORMALS)) {
        out.sn = out.gn;
    } else {
        D3 n0 = ld3(sc.normals + 3 * i0), n1 = ld3(sc.normals + 3 * i1), n2 = ld3(sc.normals + 3 * i2);
        out.sn = normalize(add(add(mul(n0, w), mul(n1, u)), mul(n2, v)));
    }
}

// Writes the record of primitive `prim` (one thread per primitive, launched once by take_gpu_scene_create).
__global__ void k_build_shade_recs(DevScene sc, double *recs) {
    const int64_t prim = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (prim >= sc.num_prims) return;
    double *r = recs + prim * sc.shade_stride;
    const uint32_t flags = sc.prim_flags[prim];
    const int32_t *id = sc.indices + 3 * prim;
    for (int k = 0; k < sc.shade_stride; ++k) r[k] = 0.0;
    r[12] = __longlong_as_double((long long)(((unsigned long long)(uint32_t)sc.prim_light[prim] << 32) | (uint32_t)sc.prim_material[prim]));
    r[13] = __longlong_as_double((long long)flags);
    if (flags & TAKE_PRIM_SPHERE) {
        const double *s = sc.spheres + 4 * (int64_t)id[0];
        r[0] = s[0]; r[1] = s[1]; r[2] = s[2]; r[3] = s[3];
        return;
    }
    const int64_t i0 = id[0], i1 = id[1], i2 = id[2];
    const D3 v0 = ld3(sc.positions + 3 * i0);
    const D3 e1 = sub(ld3(sc.positions + 3 * i1), v0), e2 = sub(ld3(sc.positions + 3 * i2), v0);
    const D3 gn = normalize(cross(e1, e2));
    r[0] = gn.x; r[1] = gn.y; r[2] = gn.z;
    if (flags & TAKE_PRIM_HAS_NORMALS) {
        const int64_t iv[3] = {i0, i1, i2};
        for (int k = 0; k < 3; ++k)
            for (int c = 0; c < 3; ++c) r[3 + 3 * k + c] = sc.normals[3 * iv[k] + c];
    }
    if ((flags & TAKE_PRIM_HAS_UVS) && sc.shade_stride >= 20) {
        const int64_t iv[3] = {i0, i1, i2};
        for (int k = 0; k < 3; ++k)
            for (int c = 0; c < 2; ++c) r[14 + 2 * k + c] = sc.uvs[2 * iv[k] + c];
    }
}

// src/shape.cpp:30-41 (sphere) and :80-108 (triangle)
template <bool LD256>
__device__ __forceinline__ void fill_isect(const DevScene &sc, D3 o, D3 d, int prim, double t, double u, double v, Isect &out) {
#if !TAKE_SHADE_RECS
    fill_isect_arrays(sc, o, d, prim, t, u, v, out);
#else
    const double2 *R = reinterpret_cast<const double2 *>(sc.shade_recs + (int64_t)prim * sc.shade_stride);
    // LD256 = false: 128-bit loads, which the compiler is free to schedule next to their uses (measured: the MIS shade kernel
    // loses its 208 bytes of spills, shade -3 ... -6 % on configs 1, 2, 4).  The environment-map kernels (config 3, textured:
    // 160-byte records) measured 2 % faster with five 256-bit loads and keep them.
    D4 r0, r1, r2, r3, r4;
    r4.a.x = r4.a.y = r4.b.x = r4.b.y = 0.0;
    if (!LD256) {
        r0.a = __ldg(R); r0.b = __ldg(R + 1); r1.a = __ldg(R + 2); r1.b = __ldg(R + 3); r2.a = __ldg(R + 4); r2.b = __ldg(R + 5);
        r3.a = __ldg(R + 6); r3.b = __ldg(R + 7);
        if (sc.shade_stride >= 20) { r4.a = __ldg(R + 8); r4.b = __ldg(R + 9); }
    } else {
        r0 = ldg_d4(R); r1 = ldg_d4(R + 2); r2 = ldg_d4(R + 4); r3 = ldg_d4(R + 6);
        if (sc.shade_stride >= 20) r4 = ldg_d4(R + 8);
    }
    out.pos = add(o, mul(d, t));
    const unsigned long long ml = (unsigned long long)__double_as_longlong(r3.a.x);
    out.material = (int32_t)(uint32_t)ml;
    out.light = (int32_t)(uint32_t)(ml >> 32);
    const uint32_t flags = (uint32_t)__double_as_longlong(r3.a.y);
    if (flags & TAKE_PRIM_SPHERE) {
        D3 gn = normalize(sub(out.pos, mk3(r0.a.x, r0.a.y, r0.b.x)));
        out.gn = dot(d, gn) < 0 ? gn : neg(gn);
        out.sn = out.gn;
        out.uv = sphere_uv(out.gn);
        return;
    }
    const D3 gn = mk3(r0.a.x, r0.a.y, r0.b.x);
    out.gn = dot(d, gn) < 0 ? gn : neg(gn);
    const double w = 1 - u - v;
    if (!(flags & TAKE_PRIM_HAS_UVS)) {
        out.uv.x = u; out.uv.y = v;
    } else {
        out.uv.x = w * r3.b.x + u * r4.a.x + v * r4.b.x;
        out.uv.y = w * r3.b.y + u * r4.a.y + v * r4.b.y;
    }
    if (!(flags & TAKE_PRIM_HAS_NORMALS)) {
        out.sn = out.gn;
    } else {
        const D3 n0 = mk3(r0.b.y, r1.a.x, r1.a.y), n1 = mk3(r1.b.x, r1.b.y, r2.a.x), n2 = mk3(r2.a.y, r2.b.x, r2.b.y);
        out.sn = normalize(add(add(mul(n0, w), mul(n1, u)), mul(n2, v)));
    }
#endif
}

// src/texture.cpp:3-26, including the wrap-column quirk (weights use x2 = 0 there) taken literally.
__device__ __forceinline__ D3 eval_texture(const DevScene &sc, const TakeMaterialDesc &m, D2 uv) {
    if (m.tex_id < 0) return mk3(m.color[0], m.color[1], m.color[2]);
    const DevTexture img = sc.textures[m.tex_id];
    double x = img.w * modulo1(m.uscale * uv.x + m.uoffset);
    double y = img.h * modulo1(m.vscale * uv.y + m.voffset);
    int x1 = (int)floor(x);
    int x2 = (x1 + 1) == img.w ? 0 : (x1 + 1);
    int y1 = (int)floor(y);
    int y2 = (y1 + 1) == img.h ? 0 : (y1 + 1);
    auto px = [&](int xx, int yy) {
        xx = min(max(xx, 0), img.w - 1);  // the reference indexes unchecked; only a rounding corner case can reach w
        yy = min(max(yy, 0), img.h - 1);
        return ld3(img.rgb + 3 * ((int64_t)yy * img.w + xx));
    };
    D3 q11 = px(x1, y1), q12 = px(x1, y2), q21 = px(x2, y1), q22 = px(x2, y2);
    if (x1 == x2) x2 += 1;
    if (y1 == y2) y2 += 1;
    D3 acc = add(add(add(mul(mul(q11, x2 - x), y2 - y), mul(mul(q21, x - x1), y2 - y)), mul(mul(q12, x2 - x), y - y1)),
                 mul(mul(q22, x - x1), y - y1));
    return divs(acc, (double)((x2 - x1) * (y2 - y1)));
}

// ---- environment map (EXTENSION without a reference counterpart, see include/take_gpu.h) ----------------------
__device__ __forceinline__ double luminance(D3 c) { return c.x * 0.212671 + c.y * 0.715160 + c.z * 0.072169; }  // vector.h:309-311
__device__ __forceinline__ double clamp1(double v) { return v < -1.0 ? -1.0 : (v > 1.0 ? 1.0 : v); }

__device__ __forceinline__ void env_texel(const DevScene &sc, D3 d, int &i, int &j, double &theta) {
    theta = acos(clamp1(d.y));
    double u = (atan2(-d.z, d.x) + TAKE_PI) / (2 * TAKE_PI), v = theta / TAKE_PI;
    i = (int)floor(u * sc.env_w);
    j = (int)floor(v * sc.env_h);
    i = i < 0 ? 0 : (i >= sc.env_w ? sc.env_w - 1 : i);
    j = j < 0 ? 0 : (j >= sc.env_h ? sc.env_h - 1 : j);
}
__device__ __forceinline__ D3 env_rgb_at(const DevScene &sc, int i, int j) { return ld3(sc.env_rgb + 3 * ((int64_t)j * sc.env_w + i)); }
__device__ __forceinline__ double env_func(const DevScene &sc, int i, int j) {
    return luminance(env_rgb_at(sc, i, j)) * sin(TAKE_PI * (j + 0.5) / sc.env_h);
}
__device__ __forceinline__ D3 env_radiance(const DevScene &sc, D3 d) {
    int i, j;
    double theta;
    env_texel(sc, d, i, j, theta);
    return env_rgb_at(sc, i, j);
}
__device__ __forceinline__ D3 miss_radiance(const DevScene &sc, D3 d) { return sc.env_rgb ? env_radiance(sc, d) : sc.background; }
__device__ __forceinline__ double env_pdf(const DevScene &sc, D3 d) {
    int i, j;
    double theta;
    env_texel(sc, d, i, j, theta);
    double st = sin(theta);
    if (!(st > 0) || !(sc.env_total > 0)) return 0;
    return env_func(sc, i, j) / sc.env_total * ((double)sc.env_w * sc.env_h) / (2 * TAKE_PI * TAKE_PI * st);
}
__device__ __forceinline__ int upper_bound_idx(const double *a, int n, double x) {  // first index with a[idx] > x (light.cpp:14)
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = (lo + hi) / 2;
        if (x < a[mid]) hi = mid; else lo = mid + 1;
    }
    return lo;
}
__device__ __forceinline__ void env_sample_dir(const DevScene &sc, double u1, double u2, D3 &dir, double &pdf) {
    pdf = 0;
    dir = mk3(0, 1, 0);
    if (!(sc.env_total > 0)) return;
    const int W = sc.env_w, H = sc.env_h;
    double x = u1 * sc.env_total;
    int j = upper_bound_idx(sc.env_marg, H + 1, x) - 1;
    j = j < 0 ? 0 : (j > H - 1 ? H - 1 : j);
    double row = sc.env_marg[j + 1] - sc.env_marg[j];
    if (!(row > 0)) return;
    double dv = (x - sc.env_marg[j]) / row;
    const double *c = sc.env_cond + (int64_t)j * (W + 1);
    double y = u2 * c[W];
    int i = upper_bound_idx(c, W + 1, y) - 1;
    i = i < 0 ? 0 : (i > W - 1 ? W - 1 : i);
    double cell = c[i + 1] - c[i];
    if (!(cell > 0)) return;
    double du = (y - c[i]) / cell;
    double u = (i + du) / W, v = (j + dv) / H;
    double theta = v * TAKE_PI, phi = u * (2 * TAKE_PI) - TAKE_PI;
    double st = sin(theta);
    dir = mk3(st * cos(phi), cos(theta), -(st * sin(phi)));
    if (!(st > 0)) return;
    pdf = env_func(sc, i, j) / sc.env_total * ((double)W * H) / (2 * TAKE_PI * TAKE_PI * st);
}

// ---- src/material.h:121-140 -------------------------------------------------------------------
__device__ __forceinline__ D3 sample_hemisphere_cos(Rng &rng) {
    double u1 = rng.next();
    double u2 = rng.next();
    double phi = TAKE_TWOPI * u2;
    double sqrt_u1 = sqrt(clampd(u1, 0, 1));
    return mk3(cos(phi) * sqrt_u1, sin(phi) * sqrt_u1, sqrt(clampd(1 - u1, 0, 1)));
}
__device__ __forceinline__ double blinn_G_hat(D3 omega, D3 n, double alpha) {
    double odn = dot(omega, n);
    double a = sqrt(0.5 * alpha + 1) / sqrt(1 / (odn * odn) - 1);
    double a2 = a * a;
    return a < 1.6 ? (3.535 * a + 2.181 * a2) / (1 + 2.276 * a + 2.577 * a2) : 1;
}
__device__ __forceinline__ D3 shading_n(D3 dir_in, const Isect &v) { return dot(dir_in, v.sn) < 0 ? neg(v.sn) : v.sn; }
__device__ __forceinline__ D3 reflect(D3 dir_in, D3 n) { return add(neg(dir_in), mul(n, 2 * dot(dir_in, n))); }

__device__ __forceinline__ bool is_lambert_like(int t) {
    return t == TAKE_MAT_DIFFUSE || (t >= TAKE_MAT_DISNEY_DIFFUSE && t <= TAKE_MAT_DISNEY_BSDF);
}
// EXTENSION (TAKE_MAT_GGX): GGX distribution and Smith G1, no reference counterpart
__device__ __forceinline__ double ggx_D(double ndh, double alpha) {
    double a2 = alpha * alpha;
    double k = ndh * ndh * (a2 - 1) + 1;
    return a2 / (TAKE_PI * k * k);
}
__device__ __forceinline__ double ggx_G1(double ndw, double alpha) {
    double a2 = alpha * alpha;
    return 2 * ndw / (ndw + sqrt(a2 + (1 - a2) * ndw * ndw));
}
__device__ __forceinline__ bool is_specular(int t) { return t == TAKE_MAT_PLASTIC || t == TAKE_MAT_MIRROR; }

// phong.inl:9-19 / blinn_phong.inl:9-19 / blinn_phong_microfacet.inl:9-19: power-cosine lobe in a local frame
__device__ __forceinline__ D3 sample_power_cos_lobe(double exponent, Rng &rng) {
    double u1 = rng.next();
    double u2 = rng.next();
    double ra1 = 1 / (exponent + 1);
    double phi = TAKE_TWOPI * u2;
    double sqrt_u1 = sqrt(clampd(1 - pow(u1, 2 * ra1), 0, 1));
    return normalize(mk3(cos(phi) * sqrt_u1, sin(phi) * sqrt_u1, clampd(pow(u1, ra1), 0, 1)));
}

// sample_bsdf (src/material.cpp:76-82 + materials/*.inl); false == std::nullopt
__device__ __forceinline__ bool sample_bsdf(const TakeMaterialDesc &m, D3 dir_in, const Isect &v, Rng &rng, D3 &dir_out, double &pdf) {
    if (dot(v.gn, dir_in) < 0) return false;
    D3 n = shading_n(dir_in, v);
    const int t = m.type;
    if (is_lambert_like(t)) {  // diffuse.inl:1-14, disney_*.inl:1-14
        dir_out = to_world(n, sample_hemisphere_cos(rng));
        pdf = dot(v.gn, dir_out) < 0 ? 0.0 : fmax(dot(n, dir_out), 0.0) / TAKE_PI;
        return true;
    }
    if (t == TAKE_MAT_MIRROR) {  // mirror.inl:1-10
        dir_out = reflect(dir_in, n);
        pdf = 1;
        return true;
    }
    if (t == TAKE_MAT_PLASTIC) {  // plastic.inl:1-27
        D3 rd = reflect(dir_in, n);
        double eta = m.p[0];
        double F0 = pow((eta - 1) / (eta + 1), 2.0);
        double F = F0 + (1 - F0) * pow(1 - dot(n, rd), 5.0);
        double u = rng.next();
        if (u <= F) {
            dir_out = rd;
            pdf = 1;
        } else {
            dir_out = to_world(n, sample_hemisphere_cos(rng));
            pdf = dot(v.gn, dir_out) < 0 ? 0.0 : fmax(dot(n, dir_out), 0.0) / TAKE_PI;
        }
        return true;
    }
    if (t == TAKE_MAT_GGX) {  // EXTENSION: sample h ~ D(h) (n.h), reflect; structure of blinn_phong_microfacet.inl:1-29
        const double alpha = m.p[0];
        double u1 = rng.next();
        double u2 = rng.next();
        double phi = TAKE_TWOPI * u2;
        double cos_t = sqrt(clampd((1 - u1) / (1 + (alpha * alpha - 1) * u1), 0, 1));
        double sin_t = sqrt(clampd(1 - cos_t * cos_t, 0, 1));
        D3 h = normalize(to_world(n, mk3(cos(phi) * sin_t, sin(phi) * sin_t, cos_t)));
        dir_out = normalize(add(neg(dir_in), mul(h, 2 * dot(dir_in, h))));
        if (dot(v.gn, dir_out) <= 0 || dot(h, n) <= 0 || dot(dir_out, h) <= 0) pdf = 0;
        else pdf = ggx_D(clampd(dot(n, h), 0, 1), alpha) * dot(n, h) * 0.25 / dot(dir_out, h);
        return true;
    }
    const double ex = m.p[0];
    if (t == TAKE_MAT_PHONG) {  // phong.inl:1-28
        D3 local = sample_power_cos_lobe(ex, rng);
        D3 rd = normalize(reflect(dir_in, n));
        dir_out = normalize(to_world(rd, local));
        pdf = dot(v.gn, dir_out) < 0 ? 0.0 : fmax(0.0, (ex + 1) / TAKE_TWOPI * pow(dot(rd, dir_out), ex));
        return true;
    }
    // blinn_phong.inl:1-29 / blinn_phong_microfacet.inl:1-29
    D3 local_h = sample_power_cos_lobe(ex, rng);
    D3 h = normalize(to_world(n, local_h));
    dir_out = normalize(add(neg(dir_in), mul(h, 2 * dot(dir_in, h))));
    if (dot(v.gn, dir_out) <= 0 || dot(h, n) <= 0 || dot(dir_out, h) <= 0) {
        pdf = 0;
    } else if (t == TAKE_MAT_BLINN_PHONG) {
        pdf = (ex + 1) * 0.25 * TAKE_INVTWOPI * pow(dot(n, h), ex) / dot(dir_out, h);
    } else {
        pdf = (ex + 1) * 0.25 * TAKE_INVTWOPI * pow(clampd(dot(n, h), 0, 1), ex) / dot(dir_out, h);
    }
    return true;
}

// get_bsdf_pdf (src/material.cpp:84-90 + materials/*.inl)
__device__ __forceinline__ double bsdf_pdf(const TakeMaterialDesc &m, D3 dir_in, D3 dir_out, const Isect &v) {
    const int t = m.type;
    if (t == TAKE_MAT_MIRROR) return 0;  // mirror.inl:12-14
    if (dot(v.gn, dir_out) < 0) return 0;
    D3 n = shading_n(dir_in, v);
    if (is_lambert_like(t)) return fmax(dot(n, dir_out), 0.0) / TAKE_PI;  // diffuse.inl:16-21
    if (t == TAKE_MAT_PLASTIC) {  // plastic.inl:29-38
        double eta = m.p[0];
        double F0 = pow((eta - 1) / (eta + 1), 2.0);
        double F = F0 + (1 - F0) * pow(1 - dot(n, dir_out), 5.0);
        return (1 - F) * fmax(dot(n, dir_out), 0.0) / TAKE_PI;
    }
    if (t == TAKE_MAT_GGX) {  // EXTENSION
        D3 h = normalize(add(dir_out, dir_in));
        if (dot(v.gn, dir_out) <= 0 || dot(h, n) <= 0 || dot(dir_out, h) <= 0) return 0;
        return ggx_D(clampd(dot(n, h), 0, 1), m.p[0]) * dot(n, h) * 0.25 / dot(dir_out, h);
    }
    const double ex = m.p[0];
    if (t == TAKE_MAT_PHONG) {  // phong.inl:30-40
        D3 rd = normalize(reflect(dir_in, n));
        return fmax(0.0, (ex + 1) / TAKE_TWOPI * pow(dot(rd, dir_out), ex));
    }
    D3 h = normalize(add(dir_out, dir_in));  // blinn_phong.inl:31-41, blinn_phong_microfacet.inl:31-41
    if (dot(v.gn, dir_out) <= 0 || dot(h, n) <= 0 || dot(dir_out, h) <= 0) return 0;
    if (t == TAKE_MAT_BLINN_PHONG) return (ex + 1) * 0.25 * TAKE_INVTWOPI * pow(dot(n, h), ex) / dot(dir_out, h);
    return (ex + 1) * 0.25 * TAKE_INVTWOPI * pow(clampd(dot(n, h), 0, 1), ex) / dot(dir_out, h);
}

// eval (src/material.cpp:92-98 + materials/*.inl): BSDF * cos.  rec_pdf is SampleRecord::pdf, which Plastic::eval
// uses to tell its two lobes apart (plastic.inl:44).
__device__ __forceinline__ D3 bsdf_eval(const DevScene &sc, const TakeMaterialDesc &m, D3 dir_in, D3 dir_out, double rec_pdf,
                               const Isect &v) {
    const D3 zero = mk3(0, 0, 0);
    if (dot(v.gn, dir_in) < 0 || dot(v.gn, dir_out) < 0) return zero;
    D3 n = shading_n(dir_in, v);
    const int t = m.type;
    if (t == TAKE_MAT_DISNEY_CLEARCOAT) return zero;  // disney_clearcoat.inl:22-27
    if (t == TAKE_MAT_DIFFUSE || t == TAKE_MAT_DISNEY_METAL || t == TAKE_MAT_DISNEY_GLASS || t == TAKE_MAT_DISNEY_SHEEN ||
        t == TAKE_MAT_DISNEY_BSDF) {  // diffuse.inl:23-29 and the Lambertian stubs
        D3 Kd = eval_texture(sc, m, v.uv);
        return divs(mul(Kd, fmax(dot(n, dir_out), 0.0)), TAKE_PI);
    }
    if (t == TAKE_MAT_MIRROR) {  // mirror.inl:16-23
        D3 F0 = eval_texture(sc, m, v.uv);
        return add(F0, mul(rsub(1, F0), pow(1 - dot(n, dir_out), 5.0)));
    }
    if (t == TAKE_MAT_PLASTIC) {  // plastic.inl:40-52
        if (rec_pdf == 1.0) return mk3(1, 1, 1);
        D3 Kd = eval_texture(sc, m, v.uv);
        return divs(mul(Kd, fmax(dot(n, dir_out), 0.0)), TAKE_PI);
    }
    if (t == TAKE_MAT_GGX) {  // EXTENSION: F D G / (4 n.w_in), the form of blinn_phong_microfacet.inl:43-60
        D3 h = normalize(add(dir_out, dir_in));
        if (dot(n, dir_out) <= 0 || dot(dir_out, h) <= 0 || dot(dir_in, h) <= 0) return zero;
        D3 Ks = eval_texture(sc, m, v.uv);
        D3 Fh = add(Ks, mul(rsub(1, Ks), pow(1 - dot(h, dir_out), 5.0)));
        double Dh = ggx_D(clampd(dot(n, h), 0, 1), m.p[0]);
        double G = ggx_G1(dot(n, dir_out), m.p[0]) * ggx_G1(dot(n, dir_in), m.p[0]);
        return divs(mul(mul(mul(Fh, Dh), G), 0.25), dot(n, dir_in));
    }
    const double ex = m.p[0];
    if (t == TAKE_MAT_PHONG) {  // phong.inl:42-54
        D3 rd = normalize(reflect(dir_in, n));
        D3 Ks = eval_texture(sc, m, v.uv);
        if (dot(n, dir_out) <= 0) return zero;
        return mul(divs(mul(Ks, ex + 1), TAKE_TWOPI), pow(fmax(dot(dir_out, rd), 0.0), ex));
    }
    if (t == TAKE_MAT_BLINN_PHONG) {  // blinn_phong.inl:43-56
        if (dot(n, dir_out) <= 0) return zero;
        D3 h = normalize(add(dir_out, dir_in));
        D3 Ks = eval_texture(sc, m, v.uv);
        D3 Fh = add(Ks, mul(rsub(1, Ks), pow(1 - dot(h, dir_out), 5.0)));
        double s = (ex + 2) * 0.25 * TAKE_INVPI / (2 - pow(2.0, -ex / 2));
        return mul(mul(Fh, s), pow(fmax(0.0, dot(n, h)), ex));
    }
    if (t == TAKE_MAT_BLINN_MICROFACET) {  // blinn_phong_microfacet.inl:43-60
        D3 h = normalize(add(dir_out, dir_in));
        if (dot(n, dir_out) <= 0 || dot(dir_out, h) <= 0 || dot(dir_in, h) <= 0) return zero;
        D3 Ks = eval_texture(sc, m, v.uv);
        D3 Fh = add(Ks, mul(rsub(1, Ks), pow(1 - dot(h, dir_out), 5.0)));
        double Dh = (ex + 2) * TAKE_INVTWOPI * pow(clampd(dot(n, h), 0, 1), ex);
        double G = blinn_G_hat(dir_out, n, ex) * blinn_G_hat(dir_in, n, ex);
        return divs(mul(mul(mul(Fh, Dh), G), 0.25), dot(n, dir_in));
    }
    // TAKE_MAT_DISNEY_DIFFUSE: disney_diffuse.inl:22-47
    D3 h = normalize(add(dir_in, dir_out));
    double hdout = dot(h, dir_out), ndout = dot(n, dir_out), ndin = dot(n, dir_in);
    D3 Kd = eval_texture(sc, m, v.uv);
    double rough = m.p[0], subsurface = m.p[1];
    double p_in = pow(1 - dot(n, dir_in), 5.0), p_out = pow(1 - dot(n, dir_out), 5.0);
    double FD90 = 0.5 + 2 * rough * hdout * hdout;
    D3 f_base = mul(mul(mul(mul(Kd, TAKE_INVPI), 1 + (FD90 - 1) * p_in), 1 + (FD90 - 1) * p_out), ndout);
    double FSS90 = rough * hdout * hdout;
    double inner = (1 + (FSS90 - 1) * p_in) * (1 + (FSS90 - 1) * p_out) * (1 / (fabs(ndin) + fabs(ndout)) - 0.5) + 0.5;
    D3 f_ss = mul(mul(mul(mul(Kd, 1.25), TAKE_INVPI), inner), ndout);
    return add(mul(f_base, 1 - subsurface), mul(f_ss, subsurface));
}

// ---- lights: src/light.cpp:5-7,32-56, src/shape.cpp:125-184 ------------------------------------------
__device__ __forceinline__ double prim_area(const DevScene &sc, int prim) {  // get_area_op, shape.cpp:171-184
    const int32_t *id = sc.indices + 3 * (int64_t)prim;
    if (sc.prim_flags[prim] & TAKE_PRIM_SPHERE) {
        double r = sc.spheres[4 * (int64_t)id[0] + 3];
        return 4 * TAKE_PI * r * r;
    }
    D3 v0 = ld3(sc.positions + 3 * (int64_t)id[0]);
    return length(cross(sub(ld3(sc.positions + 3 * (int64_t)id[1]), v0), sub(ld3(sc.positions + 3 * (int64_t)id[2]), v0))) / 2;
}

// Per-light record (LightRec, TAKE_LIGHT_REC_STRIDE doubles, 256-byte aligned): what sample_on_light / get_light_pdf need
// about the emitter's primitive, gathered once on the device (k_build_light_recs) -- one dependent level
// light id -> record instead of light -> primitive -> {flags, indices} -> {positions, normals}, and the per-light
// constants (geometric normal, 1 / area) are computed once instead of once per connection, with the very operations
// the reference performs per call (shape.cpp:146-184), so the values are the same bits.
//   [0..8]   triangle: v0, v1, v2            sphere: centre [0..2], radius [3]
//   [9..11]  triangle: normalize(cross(v1 - v0, v2 - v0))
//   [12..20] triangle: the three vertex normals
//   [21]     triangle: 1 / get_area  (light.cpp:46-47)
//   [22]     bits: primitive flags
#define TAKE_LIGHT_REC_STRIDE 32
#ifndef TAKE_LIGHT_RECS
#define TAKE_LIGHT_RECS 1
#endif

__global__ void k_build_light_recs(DevScene sc, double *recs) {
    const int li = blockIdx.x * blockDim.x + threadIdx.x;
    if (li >= sc.num_lights) return;
    double *r = recs + (int64_t)li * TAKE_LIGHT_REC_STRIDE;
    for (int k = 0; k < TAKE_LIGHT_REC_STRIDE; ++k) r[k] = 0.0;
    const TakeLightDesc &l = sc.lights[li];
    if (l.kind != TAKE_LIGHT_AREA) return;
    const int prim = l.prim_id;
    const uint32_t flags = sc.prim_flags[prim];
    r[22] = __longlong_as_double((long long)flags);
    const int32_t *id = sc.indices + 3 * (int64_t)prim;
    if (flags & TAKE_PRIM_SPHERE) {
        const double *s = sc.spheres + 4 * (int64_t)id[0];
        r[0] = s[0]; r[1] = s[1]; r[2] = s[2]; r[3] = s[3];
        return;
    }
    const int64_t iv[3] = {id[0], id[1], id[2]};
    for (int k = 0; k < 3; ++k)
        for (int c = 0; c < 3; ++c) {
            r[3 * k + c] = sc.positions[3 * iv[k] + c];
            r[12 + 3 * k + c] = sc.normals[3 * iv[k] + c];
        }
    const D3 v0 = ld3(r), v1 = ld3(r + 3), v2 = ld3(r + 6);
    const D3 n = normalize(cross(sub(v1, v0), sub(v2, v0)));
    r[9] = n.x; r[10] = n.y; r[11] = n.z;
    r[21] = 1 / prim_area(sc, prim);
}

// sample_on_light -> sample_on_shape_op (light.cpp:50-56, shape.cpp:125-169)
__device__ __forceinline__ void sample_on_light(const DevScene &sc, int light_id, int prim, D3 ref_pos, Rng &rng, D3 &pos, D3 &nrm) {
#if TAKE_LIGHT_RECS
    // 128-bit loads, consumed in stages (position first, then the normals) so that the record never sits in registers whole
    const double2 *R = reinterpret_cast<const double2 *>(sc.light_recs + (int64_t)light_id * TAKE_LIGHT_REC_STRIDE);
    const uint32_t flags = (uint32_t)__double_as_longlong(__ldg(R + 11).x);
    const double2 q0 = __ldg(R), q1 = __ldg(R + 1);
    if (flags & TAKE_PRIM_SPHERE) {  // shape.cpp:125-144
        D3 c = mk3(q0.x, q0.y, q1.x);
        double u1 = rng.next();
        double u2 = rng.next();
        double r = q1.y;
        double d = length(sub(c, ref_pos));
        double z = 1 + u1 * (r / d - 1);
        double z2 = z * z;
        double sin_theta = sqrt(clampd(1 - z2, 0, 1));
        D3 local_p = normalize(mk3(cos(2 * TAKE_PI * u2) * sin_theta, sin(2 * TAKE_PI * u2) * sin_theta, z));
        nrm = normalize(to_world(normalize(sub(ref_pos, c)), local_p));
        pos = add(c, mul(nrm, r));
        return;
    }
    double u1 = rng.next();  // shape.cpp:146-169
    double u2 = rng.next();
    double b1 = 1 - sqrt(u1);
    double b2 = sqrt(u1) * u2;
    double b0 = 1 - b1 - b2;
    const double2 q2 = __ldg(R + 2), q3 = __ldg(R + 3), q4 = __ldg(R + 4);
    const D3 v0 = mk3(q0.x, q0.y, q1.x), v1 = mk3(q1.y, q2.x, q2.y), v2 = mk3(q3.x, q3.y, q4.x);
    pos = add(add(mul(v0, b0), mul(v1, b1)), mul(v2, b2));
    const double2 q5 = __ldg(R + 5), q6 = __ldg(R + 6), q7 = __ldg(R + 7), q8 = __ldg(R + 8), q9 = __ldg(R + 9), q10 = __ldg(R + 10);
    const D3 n = mk3(q4.y, q5.x, q5.y);
    const D3 n0 = mk3(q6.x, q6.y, q7.x), n1 = mk3(q7.y, q8.x, q8.y), n2 = mk3(q9.x, q9.y, q10.x);
    D3 sn = add(add(mul(n0, b0), mul(n1, b1)), mul(n2, b2));
    nrm = dot(sn, n) > 0 ? n : neg(n);
#else
    const int32_t *id = sc.indices + 3 * (int64_t)prim;
    if (sc.prim_flags[prim] & TAKE_PRIM_SPHERE) {  // shape.cpp:125-144
        const double *s = sc.spheres + 4 * (int64_t)id[0];
        D3 c = mk3(s[0], s[1], s[2]);
        double u1 = rng.next();
        double u2 = rng.next();
        double r = s[3];
        double d = length(sub(c, ref_pos));
        double z = 1 + u1 * (r / d - 1);
        double z2 = z * z;
        double sin_theta = sqrt(clampd(1 - z2, 0, 1));
        D3 local_p = normalize(mk3(cos(2 * TAKE_PI * u2) * sin_theta, sin(2 * TAKE_PI * u2) * sin_theta, z));
        nrm = normalize(to_world(normalize(sub(ref_pos, c)), local_p));
        pos = add(c, mul(nrm, r));
        return;
    }
    const int64_t i0 = id[0], i1 = id[1], i2 = id[2];  // shape.cpp:146-169
    D3 v0 = ld3(sc.positions + 3 * i0), v1 = ld3(sc.positions + 3 * i1), v2 = ld3(sc.positions + 3 * i2);
    double u1 = rng.next();
    double u2 = rng.next();
    double b1 = 1 - sqrt(u1);
    double b2 = sqrt(u1) * u2;
    double b0 = 1 - b1 - b2;
    pos = add(add(mul(v0, b0), mul(v1, b1)), mul(v2, b2));
    D3 n = normalize(cross(sub(v1, v0), sub(v2, v0)));
    D3 sn = add(add(mul(ld3(sc.normals + 3 * i0), b0), mul(ld3(sc.normals + 3 * i1), b1)), mul(ld3(sc.normals + 3 * i2), b2));
    nrm = dot(sn, n) > 0 ? n : neg(n);
#endif
}

__device__ __forceinline__ double light_pdf_area(const DevScene &sc, int light_id, D3 light_pos, D3 ref_pos) {  // light.cpp:32-48
    const TakeLightDesc &l = sc.lights[light_id];
    if (l.kind != TAKE_LIGHT_AREA) return 0;
#if TAKE_LIGHT_RECS
    const double *r = sc.light_recs + (int64_t)light_id * TAKE_LIGHT_REC_STRIDE;
    if ((uint32_t)__double_as_longlong(__ldg(r + 22)) & TAKE_PRIM_SPHERE) {
        double rad = __ldg(r + 3);
        double d = length(sub(light_pos, ref_pos));
        return 1 / (TAKE_TWOPI * rad * rad * (1 - rad / d));
    }
    return __ldg(r + 21);
#else
    const int prim = l.prim_id;
    if (sc.prim_flags[prim] & TAKE_PRIM_SPHERE) {
        double r = sc.spheres[4 * (int64_t)sc.indices[3 * (int64_t)prim] + 3];
        double d = length(sub(light_pos, ref_pos));
        return 1 / (TAKE_TWOPI * r * r * (1 - r / d));
    }
    return 1 / prim_area(sc, prim);
#endif
}

__device__ __forceinline__ D3 light_intensity(const TakeLightDesc &l) { return mk3(l.intensity[0], l.intensity[1], l.intensity[2]); }

}  // namespace take
