// TEST INFRASTRUCTURE -- not part of the shipped product.  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline leg may load the library built from
// this file; the product (take_b200/) must never call it.
//
// CPU restatement of TaKe's hot path (TaKeTube/TaKe, /root/reference) in plain C-style C++:
// BVH build, ray/box, ray/triangle, ray/sphere, closest-hit traversal, BSDFs, textures, light
// sampling and the three live integrators, all in IEEE double with the reference's operation order
// (compiled with -ffp-contract=off; the reference build has no FMAs either).
// C++ rather than C only because the tree topology -- and with it the tie-break order of equal-t
// hits -- is defined by libstdc++'s std::sort over equal centroids (src/bvh.cpp:25-30).
//
// Parity status: PINNED.  tests/test_oracle_vs_reference.py checks this file against the unmodified
// reference (oracle/_ref/libtake_ref.so) bit-for-bit: BVH node arrays, (primitive id, t) and full hit
// records on primary + secondary rays, and per-pixel radiance sums of all three integrators driven by
// identical random streams; tests/golden/ holds vectors generated from the reference by
// tests/golden/make_golden.py for boxes without /root/reference.
//
// Each function cites the reference lines it follows.
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <thread>
#include <vector>

#include "../include/take_gpu.h"
#include "take_rng.h"

namespace {

const double EPS = 1e-7;  // c_EPSILON, src/take.h:30
const double PI = 3.14159265358979323846;
const double INVPI = 1.0 / PI;
const double TWOPI = 2.0 * PI;
const double INVTWOPI = 1.0 / TWOPI;

// ---- src/vector.h ---------------------------------------------------------------------------
struct V3 {
    double x, y, z;
    double operator[](int i) const { return (&x)[i]; }
};
struct V2 {
    double x, y;
};
inline V3 add(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 neg(V3 a) { return {-a.x, -a.y, -a.z}; }
inline V3 mul(V3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 mulv(V3 a, V3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline V3 rsub(double s, V3 a) { return {s - a.x, s - a.y, s - a.z}; }       // Real - Vector3, vector.h:143-146
inline V3 divs(V3 v, double s) { double inv = 1.0 / s; return {v.x * inv, v.y * inv, v.z * inv}; }  // :193-197
inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }  // :222-225
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline double length(V3 v) { return sqrt(dot(v, v)); }
inline V3 normalize(V3 v) {  // :249-257
    double l = length(v);
    if (l <= 0) return {0, 0, 0};
    return divs(v, l);
}
inline V3 to_world(V3 n, V3 v) {  // :314-326 (Frisvad)
    V3 x, y;
    if (n.z < -1 + 1e-6) {
        x = {0, -1, 0};
        y = {-1, 0, 0};
    } else {
        double a = 1 / (1 + n.z);
        double b = -n.x * n.y * a;
        x = {1 - n.x * n.x * a, b, -n.x};
        y = {b, 1 - n.y * n.y * a, -n.y};
    }
    return add(add(mul(x, v.x), mul(y, v.y)), mul(n, v.z));
}
inline double clampd(double v, double lo, double hi) { return (v < lo) ? lo : (hi < v) ? hi : v; }  // std::clamp
inline double modulo(double a, double b) { double r = fmod(a, b); return (r < 0.0) ? r + b : r; }  // take.h:66-69

struct Ray {
    V3 o, d;
    double tmin, tmax;
};

struct Isect {  // src/intersection.h:4-12 (+ primitive id, barycentrics)
    V3 pos, gn, sn;
    V2 uv;
    double t;
    int material, light, prim;
    double bu, bv;
};

struct Node {  // src/bvh.h:5-10
    double lo[3], hi[3];
    int left, right, prim;
};

struct Counters {
    int64_t extend = 0, shadow = 0, shaded = 0, box = 0, tri = 0, miss_after_light_sample = 0;
    void operator+=(const Counters &c) {
        extend += c.extend; shadow += c.shadow; shaded += c.shaded; box += c.box; tri += c.tri;
        miss_after_light_sample += c.miss_after_light_sample;
    }
};

struct Scene {
    TakeCamera cam;
    V3 background;
    std::vector<V3> pos, nrm;
    std::vector<V2> uv;
    std::vector<int32_t> idx, pmat, plight;
    std::vector<uint8_t> pflags;
    std::vector<double> spheres;
    std::vector<TakeMaterialDesc> mats;
    std::vector<TakeLightDesc> lights;
    struct Tex { int w, h; std::vector<double> rgb; };
    std::vector<Tex> tex;
    std::vector<Node> nodes;
    int root = -1;
    // EXTENSION without a reference counterpart (see include/take_gpu.h): lat-long environment map
    int env_w = 0, env_h = 0, env_sample = 0;
    std::vector<double> env_rgb, env_marg, env_cond;  // marginal CDF [h+1], conditional CDFs [h][w+1]
    double env_total = 0;
    bool has_env() const { return env_w > 0; }
    // tables of the power-sampling integrator (path_tracing.h:274-380): pmf[i] = light_power(i) / total, cdf = N + 1 running
    // sums from 0 to exactly 1 -- the layout light.cpp:9-23 reads; nothing in the reference fills them (ref_harness.cpp does)
    std::vector<double> power_pmf, power_cdf;
    // EXTENSION (README.md:19-24 lists Russian roulette as a goal; the reference has none): from loop iteration rr_start on,
    // a path survives with probability q = min(max component of the throughput, 0.95) and is divided by q.  0 = off.
    int rr_start = 0;
    // number of entries of the uniform light pick: the scene's lights plus, when it is sampled, the environment
    size_t pick_count() const { return lights.size() + ((has_env() && env_sample) ? 1 : 0); }
};

// ---- src/bvh.cpp:8-45 + src/scene.cpp:4-23 ----------------------------------------------------
struct KeyId {
    double key;
    int id;
};

int construct(Scene &sc, const std::vector<Node> &leaf, std::vector<int> &ids, int lo, int hi) {
    if (hi - lo == 1) {
        sc.nodes.push_back(leaf[ids[lo]]);
        return (int)sc.nodes.size() - 1;
    }
    Node big;
    for (int a = 0; a < 3; ++a) { big.lo[a] = INFINITY; big.hi[a] = -INFINITY; }
    for (int i = lo; i < hi; ++i) {  // merge(), src/bbox.h:45-55
        const Node &b = leaf[ids[i]];
        for (int a = 0; a < 3; ++a) {
            big.lo[a] = std::min(big.lo[a], b.lo[a]);
            big.hi[a] = std::max(big.hi[a], b.hi[a]);
        }
    }
    double ex = big.hi[0] - big.lo[0], ey = big.hi[1] - big.lo[1], ez = big.hi[2] - big.lo[2];
    int axis = (ex > ey && ex > ez) ? 0 : (ey > ex && ey > ez) ? 1 : 2;  // largest_axis, bbox.h:34-43
    std::vector<KeyId> tmp(hi - lo);
    for (int i = lo; i < hi; ++i) {
        const Node &b = leaf[ids[i]];
        tmp[i - lo] = {(b.hi[axis] + b.lo[axis]) * 0.5, ids[i]};  // (p_max + p_min) / Real(2) -> * (1/2)
    }
    // Same algorithm (libstdc++ introsort) and same comparison outcomes as the reference's sort of
    // BBoxWithID values, hence the same permutation, including among equal centroids.
    std::sort(tmp.begin(), tmp.end(), [](const KeyId &a, const KeyId &b) { return a.key < b.key; });
    for (int i = lo; i < hi; ++i) ids[i] = tmp[i - lo].id;
    tmp = std::vector<KeyId>();
    int mid = lo + (hi - lo) / 2;
    big.left = construct(sc, leaf, ids, lo, mid);
    big.right = construct(sc, leaf, ids, mid, hi);
    big.prim = -1;
    sc.nodes.push_back(big);
    return (int)sc.nodes.size() - 1;
}

void build_bvh(Scene &sc) {
    int n = (int)sc.pmat.size();
    std::vector<Node> leaf(n);
    for (int i = 0; i < n; ++i) {
        Node &b = leaf[i];
        b.left = b.right = -1;
        b.prim = i;
        if (sc.pflags[i] & TAKE_PRIM_SPHERE) {
            const double *s = &sc.spheres[4 * sc.idx[3 * i]];
            for (int a = 0; a < 3; ++a) { b.lo[a] = s[a] - s[3]; b.hi[a] = s[a] + s[3]; }
        } else {
            V3 p0 = sc.pos[sc.idx[3 * i]], p1 = sc.pos[sc.idx[3 * i + 1]], p2 = sc.pos[sc.idx[3 * i + 2]];
            for (int a = 0; a < 3; ++a) {
                b.lo[a] = std::min(std::min(p0[a], p1[a]), p2[a]);
                b.hi[a] = std::max(std::max(p0[a], p1[a]), p2[a]);
            }
        }
    }
    sc.nodes.clear();
    sc.nodes.reserve(2 * (size_t)n);
    std::vector<int> ids(n);
    for (int i = 0; i < n; ++i) ids[i] = i;
    sc.root = n > 0 ? construct(sc, leaf, ids, 0, n) : -1;
}

// ---- src/bbox.h:18-32 ------------------------------------------------------------------------
inline bool hit_box(const Node &b, const Ray &r) {
    double t_min = r.tmin, t_max = r.tmax;
    for (int a = 0; a < 3; ++a) {
        double ta = (b.lo[a] - r.o[a]) / r.d[a];
        double tb = (b.hi[a] - r.o[a]) / r.d[a];
        double t0 = fmin(ta, tb), t1 = fmax(ta, tb);
        t_min = fmax(t0, t_min);
        t_max = fmin(t1, t_max);
        if (t_max < t_min) return false;
    }
    return true;
}

// ---- src/shape.cpp:44-78 (the accept / reject part) --------------------------------------------
inline bool hit_triangle(const Scene &sc, int prim, const Ray &r, double &t, double &bu, double &bv) {
    V3 v0 = sc.pos[sc.idx[3 * prim]], v1 = sc.pos[sc.idx[3 * prim + 1]], v2 = sc.pos[sc.idx[3 * prim + 2]];
    V3 e1 = sub(v1, v0), e2 = sub(v2, v0);
    V3 h = cross(r.d, e2);
    double a = dot(e1, h);
    if (a > -EPS && a < EPS) return false;
    double f = 1.0 / a;
    V3 s = sub(r.o, v0);
    double u = f * dot(s, h);
    if (u < 0.0 || u > 1.0) return false;
    V3 q = cross(s, e1);
    double v = f * dot(r.d, q);
    if (v < 0.0 || u + v > 1.0) return false;
    double tt = f * dot(e2, q);
    if (tt < r.tmin || r.tmax < tt) return false;
    t = tt; bu = u; bv = v;
    return true;
}

// ---- src/shape.cpp:13-29 -----------------------------------------------------------------------
inline bool hit_sphere(const Scene &sc, int prim, const Ray &r, double &t) {
    const double *s = &sc.spheres[4 * sc.idx[3 * prim]];
    V3 c = {s[0], s[1], s[2]};
    double radius = s[3];
    V3 oc = sub(r.o, c);
    double a = dot(r.d, r.d);
    double half_b = dot(oc, r.d);
    double cc = dot(oc, oc) - radius * radius;
    double disc = half_b * half_b - a * cc;
    if (disc < 0) return false;
    double sqrtd = sqrt(disc);
    double root = (-half_b - sqrtd) / a;
    if (root < r.tmin || r.tmax < root) {
        root = (-half_b + sqrtd) / a;
        if (root < r.tmin || r.tmax < root) return false;
    }
    t = root;
    return true;
}

struct LeafHit {
    int prim = -1;
    double t = 0, u = 0, v = 0;
};

inline LeafHit hit_prim(const Scene &sc, int prim, const Ray &r, Counters &cn) {
    LeafHit h;
    cn.tri++;
    double t, u = 0, v = 0;
    bool ok = (sc.pflags[prim] & TAKE_PRIM_SPHERE) ? hit_sphere(sc, prim, r, t) : hit_triangle(sc, prim, r, t, u, v);
    if (ok) { h.prim = prim; h.t = t; h.u = u; h.v = v; }
    return h;
}

// ---- src/bvh.cpp:86-109 ------------------------------------------------------------------------
LeafHit traverse(const Scene &sc, int node_id, Ray ray, Counters &cn) {
    const Node &node = sc.nodes[node_id];
    if (node.prim != -1) return hit_prim(sc, node.prim, ray, cn);
    LeafHit left;
    cn.box++;
    if (hit_box(sc.nodes[node.left], ray)) {
        left = traverse(sc, node.left, ray, cn);
        if (left.prim != -1) ray.tmax = left.t;
    }
    cn.box++;
    if (hit_box(sc.nodes[node.right], ray)) {
        LeafHit right = traverse(sc, node.right, ray, cn);
        if (right.prim != -1) return right;
    }
    return left;
}

inline V2 sphere_uv(V3 p) {  // src/shape.cpp:3-11
    double theta = acos(-p.y);
    double phi = atan2(-p.z, p.x) + PI;
    return {phi / (2 * PI), -theta / PI};
}

// Hit record: src/shape.cpp:30-41 (sphere) and :80-108 (triangle).
inline void fill_isect(const Scene &sc, const Ray &r, const LeafHit &h, Isect &o) {
    int prim = h.prim;
    o.prim = prim;
    o.t = h.t;
    o.bu = h.u; o.bv = h.v;
    o.pos = add(r.o, mul(r.d, h.t));
    o.material = sc.pmat[prim];
    o.light = sc.plight[prim];
    if (sc.pflags[prim] & TAKE_PRIM_SPHERE) {
        const double *s = &sc.spheres[4 * sc.idx[3 * prim]];
        V3 gn = normalize(sub(o.pos, V3{s[0], s[1], s[2]}));
        o.gn = dot(r.d, gn) < 0 ? gn : neg(gn);
        o.sn = o.gn;
        o.uv = sphere_uv(o.gn);
        return;
    }
    int i0 = sc.idx[3 * prim], i1 = sc.idx[3 * prim + 1], i2 = sc.idx[3 * prim + 2];
    V3 e1 = sub(sc.pos[i1], sc.pos[i0]), e2 = sub(sc.pos[i2], sc.pos[i0]);
    V3 gn = normalize(cross(e1, e2));
    o.gn = dot(r.d, gn) < 0 ? gn : neg(gn);
    double u = h.u, v = h.v;
    double w = 1 - u - v;
    if (!(sc.pflags[prim] & TAKE_PRIM_HAS_UVS)) {
        o.uv = {u, v};
    } else {
        V2 a = sc.uv[i0], b = sc.uv[i1], c = sc.uv[i2];
        o.uv = {w * a.x + u * b.x + v * c.x, w * a.y + u * b.y + v * c.y};
    }
    if (!(sc.pflags[prim] & TAKE_PRIM_HAS_NORMALS)) {
        o.sn = o.gn;
    } else {
        V3 n0 = sc.nrm[i0], n1 = sc.nrm[i1], n2 = sc.nrm[i2];
        o.sn = normalize(add(add(mul(n0, w), mul(n1, u)), mul(n2, v)));
    }
}

// scene_intersect, src/scene.cpp:25-47 (BVH branch)
inline bool scene_intersect(const Scene &sc, const Ray &r, Isect &out, Counters &cn) {
    cn.extend++;
    if (sc.root < 0) return false;
    LeafHit h = traverse(sc, sc.root, r, cn);
    if (h.prim == -1) return false;
    fill_isect(sc, r, h, out);
    return true;
}
// scene_occluded, src/scene.cpp:49-64: a full closest-hit query reduced to a boolean
inline bool scene_occluded(const Scene &sc, const Ray &r, Counters &cn) {
    cn.shadow++;
    if (sc.root < 0) return false;
    return traverse(sc, sc.root, r, cn).prim != -1;
}

// ---- src/texture.cpp:3-26 ----------------------------------------------------------------------
inline V3 eval_texture(const Scene &sc, const TakeMaterialDesc &m, V2 uv) {
    if (m.tex_id < 0) return {m.color[0], m.color[1], m.color[2]};
    const Scene::Tex &img = sc.tex[m.tex_id];
    double x = img.w * modulo(m.uscale * uv.x + m.uoffset, 1.0);
    double y = img.h * modulo(m.vscale * uv.y + m.voffset, 1.0);
    int x1 = (int)floor(x);
    int x2 = (x1 + 1) == img.w ? 0 : (x1 + 1);
    int y1 = (int)floor(y);
    int y2 = (y1 + 1) == img.h ? 0 : (y1 + 1);
    auto px = [&](int xx, int yy) {
        // the reference indexes unchecked; x == width can only arise from a rounding corner case
        xx = std::min(std::max(xx, 0), img.w - 1);
        yy = std::min(std::max(yy, 0), img.h - 1);
        const double *p = &img.rgb[3 * ((size_t)yy * img.w + xx)];
        return V3{p[0], p[1], p[2]};
    };
    V3 q11 = px(x1, y1), q12 = px(x1, y2), q21 = px(x2, y1), q22 = px(x2, y2);
    if (x1 == x2) x2 += 1;
    if (y1 == y2) y2 += 1;
    V3 acc = add(add(add(mul(mul(q11, x2 - x), y2 - y), mul(mul(q21, x - x1), y2 - y)), mul(mul(q12, x2 - x), y - y1)),
                 mul(mul(q22, x - x1), y - y1));
    return divs(acc, (double)((x2 - x1) * (y2 - y1)));
}

// ---- environment map (EXTENSION: our own design, parity unpinned except the constant-map case) ---------------
// Conventions follow the reference's idioms: phi of get_sphere_uv (src/shape.cpp:3-11), luminance() (src/vector.h:
// 309-311), CDF inversion by upper_bound + clamp (src/light.cpp:9-17).
inline double luminance(V3 c) { return c.x * 0.212671 + c.y * 0.715160 + c.z * 0.072169; }
inline double clamp1(double v) { return v < -1.0 ? -1.0 : (v > 1.0 ? 1.0 : v); }

inline void env_texel(const Scene &sc, V3 d, int &i, int &j, double &theta) {
    theta = acos(clamp1(d.y));
    double u = (atan2(-d.z, d.x) + PI) / (2 * PI), v = theta / PI;
    i = (int)floor(u * sc.env_w);
    j = (int)floor(v * sc.env_h);
    i = i < 0 ? 0 : (i >= sc.env_w ? sc.env_w - 1 : i);
    j = j < 0 ? 0 : (j >= sc.env_h ? sc.env_h - 1 : j);
}
inline V3 env_rgb_at(const Scene &sc, int i, int j) {
    const double *p = &sc.env_rgb[3 * ((size_t)j * sc.env_w + i)];
    return {p[0], p[1], p[2]};
}
inline double env_func(const Scene &sc, int i, int j) {  // sampling weight of a texel: luminance x sin(theta at the row centre)
    return luminance(env_rgb_at(sc, i, j)) * sin(PI * (j + 0.5) / sc.env_h);
}
inline V3 env_radiance(const Scene &sc, V3 d) {
    int i, j;
    double theta;
    env_texel(sc, d, i, j, theta);
    return env_rgb_at(sc, i, j);
}
inline V3 miss_radiance(const Scene &sc, V3 d) { return sc.has_env() ? env_radiance(sc, d) : sc.background; }
// solid-angle pdf of sampling direction d from the tables
inline double env_pdf(const Scene &sc, V3 d) {
    int i, j;
    double theta;
    env_texel(sc, d, i, j, theta);
    double st = sin(theta);
    if (!(st > 0) || !(sc.env_total > 0)) return 0;
    return env_func(sc, i, j) / sc.env_total * ((double)sc.env_w * sc.env_h) / (2 * PI * PI * st);
}
inline int upper_bound_idx(const double *a, int n, double x) {  // first index with a[idx] > x, in [0, n]
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = (lo + hi) / 2;
        if (x < a[mid]) hi = mid; else lo = mid + 1;
    }
    return lo;
}
// Two draws -> direction and its solid-angle pdf (0 if the map is black or the direction degenerate).
inline void env_sample_dir(const Scene &sc, double u1, double u2, V3 &dir, double &pdf) {
    pdf = 0;
    dir = {0, 1, 0};
    if (!(sc.env_total > 0)) return;
    const int W = sc.env_w, H = sc.env_h;
    double x = u1 * sc.env_total;
    int j = upper_bound_idx(sc.env_marg.data(), H + 1, x) - 1;
    j = j < 0 ? 0 : (j > H - 1 ? H - 1 : j);
    double row = sc.env_marg[j + 1] - sc.env_marg[j];
    if (!(row > 0)) return;
    double dv = (x - sc.env_marg[j]) / row;
    const double *c = &sc.env_cond[(size_t)j * (W + 1)];
    double y = u2 * c[W];
    int i = upper_bound_idx(c, W + 1, y) - 1;
    i = i < 0 ? 0 : (i > W - 1 ? W - 1 : i);
    double cell = c[i + 1] - c[i];
    if (!(cell > 0)) return;
    double du = (y - c[i]) / cell;
    double u = (i + du) / W, v = (j + dv) / H;
    double theta = v * PI, phi = u * (2 * PI) - PI;
    double st = sin(theta);
    dir = {st * cos(phi), cos(theta), -(st * sin(phi))};
    if (!(st > 0)) return;
    pdf = env_func(sc, i, j) / sc.env_total * ((double)W * H) / (2 * PI * PI * st);
}
void build_env_tables(Scene &sc) {
    const int W = sc.env_w, H = sc.env_h;
    sc.env_marg.assign(H + 1, 0.0);
    sc.env_cond.assign((size_t)H * (W + 1), 0.0);
    for (int j = 0; j < H; ++j) {
        double *c = &sc.env_cond[(size_t)j * (W + 1)];
        for (int i = 0; i < W; ++i) c[i + 1] = c[i] + env_func(sc, i, j);
        sc.env_marg[j + 1] = sc.env_marg[j] + c[W];
    }
    sc.env_total = sc.env_marg[H];
}

// ---- random stream -----------------------------------------------------------------------------
struct Rng {
    uint64_t seed, sample;
    uint32_t pixel, k;
    double next() { return take_stream_real(seed, pixel, sample, k++); }
};

// ---- src/material.h:121-140 ----------------------------------------------------------------------
inline V3 sample_hemisphere_cos(Rng &rng) {
    double u1 = rng.next();
    double u2 = rng.next();
    double phi = TWOPI * u2;
    double sqrt_u1 = sqrt(clampd(u1, 0, 1));
    return {cos(phi) * sqrt_u1, sin(phi) * sqrt_u1, sqrt(clampd(1 - u1, 0, 1))};
}
inline double blinn_G_hat(V3 omega, V3 n, double alpha) {
    double odn = dot(omega, n);
    double a = sqrt(0.5 * alpha + 1) / sqrt(1 / (odn * odn) - 1);
    double a2 = a * a;
    return a < 1.6 ? (3.535 * a + 2.181 * a2) / (1 + 2.276 * a + 2.577 * a2) : 1;
}

// EXTENSION (TAKE_MAT_GGX): GGX distribution and Smith G1, no reference counterpart
inline double ggx_D(double ndh, double alpha) {
    double a2 = alpha * alpha;
    double k = ndh * ndh * (a2 - 1) + 1;
    return a2 / (PI * k * k);
}
inline double ggx_G1(double ndw, double alpha) {
    double a2 = alpha * alpha;
    return 2 * ndw / (ndw + sqrt(a2 + (1 - a2) * ndw * ndw));
}

inline V3 shading_n(V3 dir_in, const Isect &v) { return dot(dir_in, v.sn) < 0 ? neg(v.sn) : v.sn; }
inline V3 reflect(V3 dir_in, V3 n) { return add(neg(dir_in), mul(n, 2 * dot(dir_in, n))); }

inline bool is_lambert_like(int t) {
    return t == TAKE_MAT_DIFFUSE || t == TAKE_MAT_DISNEY_DIFFUSE || t == TAKE_MAT_DISNEY_METAL ||
           t == TAKE_MAT_DISNEY_GLASS || t == TAKE_MAT_DISNEY_CLEARCOAT || t == TAKE_MAT_DISNEY_SHEEN ||
           t == TAKE_MAT_DISNEY_BSDF;  // (TAKE_MAT_GGX is not)
}

// Blinn-Phong half-vector sampling shared by blinn_phong.inl:1-29 and blinn_phong_microfacet.inl:1-29;
// Phong lobe sampling phong.inl:1-28 uses the same local frame around the mirror direction.
inline V3 sample_power_cos_lobe(double exponent, Rng &rng) {
    double u1 = rng.next();
    double u2 = rng.next();
    double ra1 = 1 / (exponent + 1);
    double phi = TWOPI * u2;
    double sqrt_u1 = sqrt(clampd(1 - pow(u1, 2 * ra1), 0, 1));
    return normalize(V3{cos(phi) * sqrt_u1, sin(phi) * sqrt_u1, clampd(pow(u1, ra1), 0, 1)});
}

// sample_bsdf: src/material.cpp:76-82 + materials/*.inl.  Returns false for nullopt.
bool sample_bsdf(const TakeMaterialDesc &m, V3 dir_in, const Isect &v, Rng &rng, V3 &dir_out, double &pdf) {
    if (dot(v.gn, dir_in) < 0) return false;
    V3 n = shading_n(dir_in, v);
    int t = m.type;
    if (is_lambert_like(t)) {  // diffuse.inl:1-14, disney_*.inl:1-14
        dir_out = to_world(n, sample_hemisphere_cos(rng));
        pdf = dot(v.gn, dir_out) < 0 ? 0.0 : fmax(dot(n, dir_out), 0.0) / PI;
        return true;
    }
    if (t == TAKE_MAT_MIRROR) {  // mirror.inl:1-10
        dir_out = reflect(dir_in, n);
        pdf = 1;
        return true;
    }
    if (t == TAKE_MAT_PLASTIC) {  // plastic.inl:1-27
        V3 rd = reflect(dir_in, n);
        double eta = m.p[0];
        double F0 = pow((eta - 1) / (eta + 1), 2.0);
        double F = F0 + (1 - F0) * pow(1 - dot(n, rd), 5.0);
        double u = rng.next();
        if (u <= F) {
            dir_out = rd;
            pdf = 1;
        } else {
            dir_out = to_world(n, sample_hemisphere_cos(rng));
            pdf = dot(v.gn, dir_out) < 0 ? 0.0 : fmax(dot(n, dir_out), 0.0) / PI;
        }
        return true;
    }
    if (t == TAKE_MAT_GGX) {  // EXTENSION: sample h ~ D(h) (n.h), reflect; structure of blinn_phong_microfacet.inl:1-29
        double alpha = m.p[0];
        double u1 = rng.next();
        double u2 = rng.next();
        double phi = TWOPI * u2;
        double cos_t = sqrt(clampd((1 - u1) / (1 + (alpha * alpha - 1) * u1), 0, 1));
        double sin_t = sqrt(clampd(1 - cos_t * cos_t, 0, 1));
        V3 h = normalize(to_world(n, V3{cos(phi) * sin_t, sin(phi) * sin_t, cos_t}));
        dir_out = normalize(add(neg(dir_in), mul(h, 2 * dot(dir_in, h))));
        if (dot(v.gn, dir_out) <= 0 || dot(h, n) <= 0 || dot(dir_out, h) <= 0) pdf = 0;
        else pdf = ggx_D(clampd(dot(n, h), 0, 1), alpha) * dot(n, h) * 0.25 / dot(dir_out, h);
        return true;
    }
    double ex = m.p[0];
    if (t == TAKE_MAT_PHONG) {  // phong.inl:1-28
        V3 local = sample_power_cos_lobe(ex, rng);
        V3 rd = normalize(reflect(dir_in, n));
        dir_out = normalize(to_world(rd, local));
        pdf = dot(v.gn, dir_out) < 0 ? 0.0 : fmax(0.0, (ex + 1) / TWOPI * pow(dot(rd, dir_out), ex));
        return true;
    }
    // blinn_phong.inl:1-29 / blinn_phong_microfacet.inl:1-29
    V3 local_h = sample_power_cos_lobe(ex, rng);
    V3 h = normalize(to_world(n, local_h));
    dir_out = normalize(add(neg(dir_in), mul(h, 2 * dot(dir_in, h))));
    if (dot(v.gn, dir_out) <= 0 || dot(h, n) <= 0 || dot(dir_out, h) <= 0) {
        pdf = 0;
    } else if (t == TAKE_MAT_BLINN_PHONG) {
        pdf = (ex + 1) * 0.25 * INVTWOPI * pow(dot(n, h), ex) / dot(dir_out, h);
    } else {
        pdf = (ex + 1) * 0.25 * INVTWOPI * pow(clampd(dot(n, h), 0, 1), ex) / dot(dir_out, h);
    }
    return true;
}

// get_bsdf_pdf: src/material.cpp:84-90 + materials/*.inl
double bsdf_pdf(const TakeMaterialDesc &m, V3 dir_in, V3 dir_out, const Isect &v) {
    int t = m.type;
    if (t == TAKE_MAT_MIRROR) return 0;  // mirror.inl:12-14
    if (dot(v.gn, dir_out) < 0) return 0;
    V3 n = shading_n(dir_in, v);
    if (is_lambert_like(t)) return fmax(dot(n, dir_out), 0.0) / PI;  // diffuse.inl:16-21
    if (t == TAKE_MAT_PLASTIC) {  // plastic.inl:29-38
        double eta = m.p[0];
        double F0 = pow((eta - 1) / (eta + 1), 2.0);
        double F = F0 + (1 - F0) * pow(1 - dot(n, dir_out), 5.0);
        return (1 - F) * fmax(dot(n, dir_out), 0.0) / PI;
    }
    if (t == TAKE_MAT_GGX) {  // EXTENSION
        V3 h = normalize(add(dir_out, dir_in));
        if (dot(v.gn, dir_out) <= 0 || dot(h, n) <= 0 || dot(dir_out, h) <= 0) return 0;
        return ggx_D(clampd(dot(n, h), 0, 1), m.p[0]) * dot(n, h) * 0.25 / dot(dir_out, h);
    }
    double ex = m.p[0];
    if (t == TAKE_MAT_PHONG) {  // phong.inl:30-40
        V3 rd = normalize(reflect(dir_in, n));
        return fmax(0.0, (ex + 1) / TWOPI * pow(dot(rd, dir_out), ex));
    }
    V3 h = normalize(add(dir_out, dir_in));  // blinn_phong.inl:31-41, blinn_phong_microfacet.inl:31-41
    if (dot(v.gn, dir_out) <= 0 || dot(h, n) <= 0 || dot(dir_out, h) <= 0) return 0;
    if (t == TAKE_MAT_BLINN_PHONG) return (ex + 1) * 0.25 * INVTWOPI * pow(dot(n, h), ex) / dot(dir_out, h);
    return (ex + 1) * 0.25 * INVTWOPI * pow(clampd(dot(n, h), 0, 1), ex) / dot(dir_out, h);
}

// eval: src/material.cpp:92-98 + materials/*.inl.  Returns BSDF * cos ("FG").
V3 bsdf_eval(const Scene &sc, const TakeMaterialDesc &m, V3 dir_in, V3 dir_out, double rec_pdf, const Isect &v) {
    const V3 zero = {0, 0, 0};
    if (dot(v.gn, dir_in) < 0 || dot(v.gn, dir_out) < 0) return zero;
    V3 n = shading_n(dir_in, v);
    int t = m.type;
    if (t == TAKE_MAT_DISNEY_CLEARCOAT) return zero;  // disney_clearcoat.inl:22-27 (`return {}`)
    if (t == TAKE_MAT_DIFFUSE || t == TAKE_MAT_DISNEY_METAL || t == TAKE_MAT_DISNEY_GLASS ||
        t == TAKE_MAT_DISNEY_SHEEN || t == TAKE_MAT_DISNEY_BSDF) {  // diffuse.inl:23-29
        V3 Kd = eval_texture(sc, m, v.uv);
        return divs(mul(Kd, fmax(dot(n, dir_out), 0.0)), PI);
    }
    if (t == TAKE_MAT_MIRROR) {  // mirror.inl:16-23
        V3 F0 = eval_texture(sc, m, v.uv);
        return add(F0, mul(rsub(1, F0), pow(1 - dot(n, dir_out), 5.0)));
    }
    if (t == TAKE_MAT_PLASTIC) {  // plastic.inl:40-52
        if (rec_pdf == 1.0) return {1, 1, 1};
        V3 Kd = eval_texture(sc, m, v.uv);
        return divs(mul(Kd, fmax(dot(n, dir_out), 0.0)), PI);
    }
    if (t == TAKE_MAT_GGX) {  // EXTENSION: F D G / (4 n.w_in), the form of blinn_phong_microfacet.inl:43-60
        V3 h = normalize(add(dir_out, dir_in));
        if (dot(n, dir_out) <= 0 || dot(dir_out, h) <= 0 || dot(dir_in, h) <= 0) return zero;
        V3 Ks = eval_texture(sc, m, v.uv);
        V3 Fh = add(Ks, mul(rsub(1, Ks), pow(1 - dot(h, dir_out), 5.0)));
        double Dh = ggx_D(clampd(dot(n, h), 0, 1), m.p[0]);
        double G = ggx_G1(dot(n, dir_out), m.p[0]) * ggx_G1(dot(n, dir_in), m.p[0]);
        return divs(mul(mul(mul(Fh, Dh), G), 0.25), dot(n, dir_in));
    }
    double ex = m.p[0];
    if (t == TAKE_MAT_PHONG) {  // phong.inl:42-54
        V3 rd = normalize(reflect(dir_in, n));
        V3 Ks = eval_texture(sc, m, v.uv);
        if (dot(n, dir_out) <= 0) return zero;
        return mul(divs(mul(Ks, ex + 1), TWOPI), pow(fmax(dot(dir_out, rd), 0.0), ex));
    }
    if (t == TAKE_MAT_BLINN_PHONG) {  // blinn_phong.inl:43-56
        if (dot(n, dir_out) <= 0) return zero;
        V3 h = normalize(add(dir_out, dir_in));
        V3 Ks = eval_texture(sc, m, v.uv);
        V3 Fh = add(Ks, mul(rsub(1, Ks), pow(1 - dot(h, dir_out), 5.0)));
        double s = (ex + 2) * 0.25 * INVPI / (2 - pow(2.0, -ex / 2));
        return mul(mul(Fh, s), pow(fmax(0.0, dot(n, h)), ex));
    }
    if (t == TAKE_MAT_BLINN_MICROFACET) {  // blinn_phong_microfacet.inl:43-60
        V3 h = normalize(add(dir_out, dir_in));
        if (dot(n, dir_out) <= 0 || dot(dir_out, h) <= 0 || dot(dir_in, h) <= 0) return zero;
        V3 Ks = eval_texture(sc, m, v.uv);
        V3 Fh = add(Ks, mul(rsub(1, Ks), pow(1 - dot(h, dir_out), 5.0)));
        double Dh = (ex + 2) * INVTWOPI * pow(clampd(dot(n, h), 0, 1), ex);
        double G = blinn_G_hat(dir_out, n, ex) * blinn_G_hat(dir_in, n, ex);
        return divs(mul(mul(mul(Fh, Dh), G), 0.25), dot(n, dir_in));
    }
    // TAKE_MAT_DISNEY_DIFFUSE: disney_diffuse.inl:22-47
    V3 h = normalize(add(dir_in, dir_out));
    double hdout = dot(h, dir_out), ndout = dot(n, dir_out), ndin = dot(n, dir_in);
    V3 Kd = eval_texture(sc, m, v.uv);
    double rough = m.p[0], subsurface = m.p[1];
    auto F = [&](V3 w, double FF) { return 1 + (FF - 1) * pow(1 - dot(n, w), 5.0); };
    double FD90 = 0.5 + 2 * rough * hdout * hdout;
    V3 f_base = mul(mul(mul(mul(Kd, INVPI), F(dir_in, FD90)), F(dir_out, FD90)), ndout);
    double FSS90 = rough * hdout * hdout;
    double inner = F(dir_in, FSS90) * F(dir_out, FSS90) * (1 / (fabs(ndin) + fabs(ndout)) - 0.5) + 0.5;
    V3 f_ss = mul(mul(mul(mul(Kd, 1.25), INVPI), inner), ndout);
    return add(mul(f_base, 1 - subsurface), mul(f_ss, subsurface));
}

// ---- lights: src/light.cpp:5-7,32-56, src/shape.cpp:125-184 -------------------------------------
struct LightSample {
    V3 pos, n;
};

inline double prim_area(const Scene &sc, int prim) {  // get_area_op, shape.cpp:171-184
    if (sc.pflags[prim] & TAKE_PRIM_SPHERE) {
        double r = sc.spheres[4 * sc.idx[3 * prim] + 3];
        return 4 * PI * r * r;
    }
    V3 v0 = sc.pos[sc.idx[3 * prim]], v1 = sc.pos[sc.idx[3 * prim + 1]], v2 = sc.pos[sc.idx[3 * prim + 2]];
    return length(cross(sub(v1, v0), sub(v2, v0))) / 2;
}

inline LightSample sample_on_prim(const Scene &sc, int prim, V3 ref_pos, Rng &rng) {
    if (sc.pflags[prim] & TAKE_PRIM_SPHERE) {  // shape.cpp:125-144 (cone sampling)
        const double *s = &sc.spheres[4 * sc.idx[3 * prim]];
        V3 c = {s[0], s[1], s[2]};
        double u1 = rng.next();
        double u2 = rng.next();
        double r = s[3];
        double d = length(sub(c, ref_pos));
        double z = 1 + u1 * (r / d - 1);
        double z2 = z * z;
        double sin_theta = sqrt(clampd(1 - z2, 0, 1));
        V3 local_p = normalize(V3{cos(2 * PI * u2) * sin_theta, sin(2 * PI * u2) * sin_theta, z});
        V3 n = normalize(to_world(normalize(sub(ref_pos, c)), local_p));
        return {add(c, mul(n, r)), n};
    }
    int i0 = sc.idx[3 * prim], i1 = sc.idx[3 * prim + 1], i2 = sc.idx[3 * prim + 2];  // shape.cpp:146-169
    V3 v0 = sc.pos[i0], v1 = sc.pos[i1], v2 = sc.pos[i2];
    double u1 = rng.next();
    double u2 = rng.next();
    double b1 = 1 - sqrt(u1);
    double b2 = sqrt(u1) * u2;
    double b0 = 1 - b1 - b2;
    V3 p = add(add(mul(v0, b0), mul(v1, b1)), mul(v2, b2));
    V3 n = normalize(cross(sub(v1, v0), sub(v2, v0)));
    V3 sn = add(add(mul(sc.nrm[i0], b0), mul(sc.nrm[i1], b1)), mul(sc.nrm[i2], b2));
    return {p, dot(sn, n) > 0 ? n : neg(n)};
}

inline double light_pdf_area(const Scene &sc, int light_id, V3 light_pos, V3 ref_pos) {  // light.cpp:32-48
    const TakeLightDesc &l = sc.lights[light_id];
    if (l.kind != TAKE_LIGHT_AREA) return 0;
    int prim = l.prim_id;
    if (sc.pflags[prim] & TAKE_PRIM_SPHERE) {
        double r = sc.spheres[4 * sc.idx[3 * prim] + 3];
        double d = length(sub(light_pos, ref_pos));
        return 1 / (TWOPI * r * r * (1 - r / d));
    }
    return 1 / prim_area(sc, prim);
}

inline bool is_specular(const TakeMaterialDesc &m) { return m.type == TAKE_MAT_PLASTIC || m.type == TAKE_MAT_MIRROR; }
inline V3 intensity(const TakeLightDesc &l) { return {l.intensity[0], l.intensity[1], l.intensity[2]}; }

// EXTENSION: Russian roulette at the top of loop iteration i (one extra draw per iteration from rr_start on); false = the
// path ends here.  Unbiased: the survivors are divided by their survival probability.
inline bool rr_survives(const Scene &sc, int i, V3 &throughput, Rng &rng) {
    if (sc.rr_start <= 0 || i < sc.rr_start) return true;
    const double q = fmin(fmax(fmax(throughput.x, throughput.y), throughput.z), 0.95);
    if (!(q > 0)) return false;
    if (rng.next() >= q) return false;
    throughput = divs(throughput, q);
    return true;
}

// ---- src/integrator/path_tracing.h:5-111 ---------------------------------------------------------
V3 path_tracing(const Scene &sc, Ray r, Rng &rng, int max_depth, Counters &cn) {
    Isect v;
    if (!scene_intersect(sc, r, v, cn)) return miss_radiance(sc, r.d);
    V3 radiance = {0, 0, 0}, throughput = {1, 1, 1};
    const size_t nl = sc.pick_count();            // == lights.size() unless a sampled environment map is present
    const bool env_light = nl > sc.lights.size();
    if (v.light != -1 && sc.lights[v.light].kind == TAKE_LIGHT_AREA)
        radiance = add(radiance, mulv(throughput, intensity(sc.lights[v.light])));
    for (int i = 0; i <= max_depth; ++i) {
        if (!rr_survives(sc, i, throughput, rng)) break;
        cn.shaded++;
        V3 dir_in = neg(r.d);
        const TakeMaterialDesc &m = sc.mats[v.material];
        bool spec = is_specular(m);
        V3 C1 = {0, 0, 0};
        if (nl > 0 && !spec) {
            int light_id = (int)floor(rng.next() * nl);
            if (env_light && light_id == (int)sc.lights.size()) {  // EXTENSION: the environment as a light
                double u1 = rng.next();
                double u2 = rng.next();
                V3 light_dir;
                double pdf_w;
                env_sample_dir(sc, u1, u2, light_dir, pdf_w);
                double lpdf = pdf_w / nl;
                if (lpdf <= 0) break;
                double bpdf = bsdf_pdf(m, dir_in, light_dir, v);
                if (bpdf > 0 && !isinf(lpdf)) {
                    V3 FG = bsdf_eval(sc, m, dir_in, light_dir, 0.0, v);
                    Ray sh = {v.pos, light_dir, EPS, INFINITY};
                    if (!scene_occluded(sc, sh, cn))
                        C1 = divs(mul(mulv(FG, env_radiance(sc, light_dir)), lpdf), lpdf * lpdf + bpdf * bpdf);
                }
            } else {
            const TakeLightDesc &l = sc.lights[light_id];
            if (l.kind == TAKE_LIGHT_AREA) {
                LightSample lp = sample_on_prim(sc, l.prim_id, v.pos, rng);
                double d = length(sub(lp.pos, v.pos));
                V3 light_dir = normalize(sub(lp.pos, v.pos));
                double lpdf = light_pdf_area(sc, light_id, lp.pos, v.pos) * (d * d) /
                              (fmax(dot(neg(lp.n), light_dir), 0.0) * nl);
                if (lpdf <= 0) break;
                double bpdf = bsdf_pdf(m, dir_in, light_dir, v);
                if (bpdf > 0 && !isinf(lpdf)) {
                    V3 FG = bsdf_eval(sc, m, dir_in, light_dir, 0.0, v);
                    Ray sh = {v.pos, light_dir, EPS, (1 - EPS) * d};
                    if (!scene_occluded(sc, sh, cn))
                        C1 = divs(mul(mulv(FG, intensity(l)), lpdf), lpdf * lpdf + bpdf * bpdf);
                }
            }
            }
        }
        radiance = add(radiance, mulv(throughput, C1));

        V3 C2 = {0, 0, 0};
        V3 rec_dir;
        double rec_pdf;
        if (!sample_bsdf(m, dir_in, v, rng, rec_dir, rec_pdf)) break;
        V3 FG = bsdf_eval(sc, m, dir_in, rec_dir, rec_pdf, v);
        V3 dir_out = normalize(rec_dir);
        double bpdf = rec_pdf, lpdf = 0;
        if (bpdf <= 0) break;
        r = {v.pos, dir_out, EPS, INFINITY};
        Isect nv;
        if (!scene_intersect(sc, r, nv, cn)) {
            if (env_light) {  // EXTENSION: the miss found the sampled environment -> MIS weight as for an emitter hit (:99)
                lpdf = env_pdf(sc, dir_out) / nl;
                V3 Ce = mul(mulv(FG, env_radiance(sc, dir_out)), spec ? (1 / bpdf) : (bpdf / (lpdf * lpdf + bpdf * bpdf)));
                radiance = add(radiance, mulv(throughput, Ce));
                break;
            }
            throughput = mulv(throughput, divs(FG, bpdf));
            radiance = add(radiance, mulv(throughput, miss_radiance(sc, dir_out)));
            break;
        }
        if (nv.light != -1) {
            double d = length(sub(nv.pos, v.pos));
            V3 light_dir = normalize(sub(nv.pos, v.pos));
            lpdf = light_pdf_area(sc, nv.light, nv.pos, v.pos) * (d * d) / (fmax(dot(neg(nv.gn), light_dir), 0.0) * nl);
            if (lpdf <= 0) break;
            const TakeLightDesc &l = sc.lights[nv.light];
            if (l.kind == TAKE_LIGHT_AREA)
                C2 = mul(mulv(FG, intensity(l)), spec ? (1 / bpdf) : (bpdf / (lpdf * lpdf + bpdf * bpdf)));
        }
        radiance = add(radiance, mulv(throughput, C2));
        throughput = mulv(throughput, divs(FG, bpdf));
        v = nv;
    }
    return radiance;
}

// ---- src/integrator/path_tracing.h:114-157 -------------------------------------------------------
V3 path_tracing_raw(const Scene &sc, Ray r, Rng &rng, int max_depth, Counters &cn) {
    Isect v;
    if (!scene_intersect(sc, r, v, cn)) return miss_radiance(sc, r.d);
    V3 radiance = {0, 0, 0}, throughput = {1, 1, 1};
    for (int i = 0; i <= max_depth; ++i) {
        if (v.light != -1 && sc.lights[v.light].kind == TAKE_LIGHT_AREA) {
            radiance = add(radiance, mulv(throughput, intensity(sc.lights[v.light])));
            break;
        } else if (v.light != -1) {
            // `if (v.area_light_id != -1) {... get_if fails ...} else {...}`: a hit on a non-area light id does nothing
            continue;
        }
        if (!rr_survives(sc, i, throughput, rng)) break;
        cn.shaded++;
        V3 dir_in = neg(r.d);
        const TakeMaterialDesc &m = sc.mats[v.material];
        V3 rec_dir;
        double pdf;
        if (!sample_bsdf(m, dir_in, v, rng, rec_dir, pdf)) break;
        V3 FG = bsdf_eval(sc, m, dir_in, rec_dir, pdf, v);
        V3 dir_out = normalize(rec_dir);
        if (pdf <= 0) break;
        throughput = mulv(throughput, divs(FG, pdf));
        r = {v.pos, dir_out, EPS, INFINITY};
        Isect nv;
        if (!scene_intersect(sc, r, nv, cn)) {
            radiance = add(radiance, mulv(throughput, miss_radiance(sc, dir_out)));
            break;
        }
        v = nv;
    }
    return radiance;
}

// ---- src/integrator/path_tracing.h:161-271 -------------------------------------------------------
V3 path_tracing_one_sample_mis(const Scene &sc, Ray r, Rng &rng, int max_depth, Counters &cn) {
    Isect v;
    if (!scene_intersect(sc, r, v, cn)) return miss_radiance(sc, r.d);
    V3 radiance = {0, 0, 0}, throughput = {1, 1, 1};
    const size_t nl = sc.pick_count();            // == lights.size() unless a sampled environment map is present
    const bool env_light = nl > sc.lights.size();
    for (int i = 0; i <= max_depth; ++i) {
        if (v.light != -1 && sc.lights[v.light].kind == TAKE_LIGHT_AREA) {
            radiance = add(radiance, mulv(throughput, intensity(sc.lights[v.light])));
            break;
        }
        if (!rr_survives(sc, i, throughput, rng)) break;
        cn.shaded++;
        V3 dir_in = neg(r.d);
        const TakeMaterialDesc &m = sc.mats[v.material];
        bool spec = is_specular(m);
        if (nl > 0 && !spec && rng.next() <= 0.5) {
            int light_id = (int)floor(rng.next() * nl);
            if (env_light && light_id == (int)sc.lights.size()) {  // EXTENSION: the environment as a light
                double u1 = rng.next();
                double u2 = rng.next();
                V3 light_dir;
                double pdf_w;
                env_sample_dir(sc, u1, u2, light_dir, pdf_w);
                double lpdf = pdf_w / nl;
                if (lpdf <= 0) break;
                double bpdf = bsdf_pdf(m, dir_in, light_dir, v);
                if (bpdf <= 0) break;
                V3 FG = bsdf_eval(sc, m, dir_in, light_dir, 0.0, v);
                r = {v.pos, light_dir, EPS, INFINITY};
                Isect nv;
                throughput = mulv(throughput, divs(FG, 0.5 * lpdf + 0.5 * bpdf));
                if (!scene_intersect(sc, r, nv, cn)) {  // reached the environment: that is the light we aimed at
                    radiance = add(radiance, mulv(throughput, env_radiance(sc, light_dir)));
                    break;
                }
                v = nv;  // hit an obstacle: continue from it, as the reference's area-light branch does (:220-225)
                continue;
            }
            const TakeLightDesc &l = sc.lights[light_id];
            if (l.kind == TAKE_LIGHT_AREA) {
                LightSample lp = sample_on_prim(sc, l.prim_id, v.pos, rng);
                double d = length(sub(lp.pos, v.pos));
                V3 light_dir = normalize(sub(lp.pos, v.pos));
                double lpdf = light_pdf_area(sc, light_id, lp.pos, v.pos) * (d * d) /
                              (fmax(dot(neg(lp.n), light_dir), 0.0) * nl);
                if (lpdf <= 0) break;
                double bpdf = bsdf_pdf(m, dir_in, light_dir, v);
                if (bpdf <= 0) break;
                V3 FG = bsdf_eval(sc, m, dir_in, light_dir, 0.0, v);
                r = {v.pos, light_dir, EPS, INFINITY};
                Isect nv;
                if (!scene_intersect(sc, r, nv, cn)) {
                    // The reference dereferences the empty optional here (path_tracing.h:220, undefined
                    // behaviour); a ray aimed at a sampled light point can only miss through rounding at a
                    // triangle edge.  We terminate the path and count the event (expected: 0).
                    cn.miss_after_light_sample++;
                    break;
                }
                v = nv;
                throughput = mulv(throughput, divs(FG, 0.5 * lpdf + 0.5 * bpdf));
            }
        } else {
            V3 rec_dir;
            double rec_pdf;
            if (!sample_bsdf(m, dir_in, v, rng, rec_dir, rec_pdf)) break;
            V3 FG = bsdf_eval(sc, m, dir_in, rec_dir, rec_pdf, v);
            V3 dir_out = normalize(rec_dir);
            double bpdf = rec_pdf;
            if (bpdf <= 0) break;
            r = {v.pos, dir_out, EPS, INFINITY};
            Isect nv;
            bool hit = scene_intersect(sc, r, nv, cn);
            double pdf = (nl == 0 || spec) ? bpdf : 0.5 * bpdf;
            if (!hit) {
                if (env_light && !spec) pdf += 0.5 * (env_pdf(sc, dir_out) / nl);  // EXTENSION: mixture pdf, as :255-265 for emitters
                throughput = mulv(throughput, divs(FG, pdf));
                radiance = add(radiance, mulv(throughput, miss_radiance(sc, dir_out)));
                break;
            }
            if (!spec && nv.light != -1) {
                double d = length(sub(nv.pos, v.pos));
                V3 light_dir = normalize(sub(nv.pos, v.pos));
                double lpdf = light_pdf_area(sc, nv.light, nv.pos, v.pos) * (d * d) /
                              (fmax(dot(neg(nv.gn), light_dir), 0.0) * nl);
                if (lpdf <= 0) break;
                pdf += 0.5 * lpdf;
            }
            throughput = mulv(throughput, divs(FG, pdf));
            v = nv;
        }
    }
    return radiance;
}

// ---- src/integrator/path_tracing.h:274-380: one-sample MIS with lights picked by power ------------------------------
// Dead code in the reference as shipped (its tables are never filled); pinned through oracle/ref_harness.cpp, which builds
// the tables the reference's readers expect -- the same way Scene::build_power_tables does here.
inline int sample_light_power(const Scene &sc, Rng &rng) {  // light.cpp:9-17
    const double u = rng.next();
    const int size = (int)sc.power_cdf.size() - 1;
    const double *ptr = std::upper_bound(sc.power_cdf.data(), sc.power_cdf.data() + size + 1, u);
    return std::min(std::max((int)(ptr - sc.power_cdf.data() - 1), 0), size - 1);
}

V3 path_tracing_one_sample_mis_power(const Scene &sc, Ray r, Rng &rng, int max_depth, Counters &cn) {
    Isect v;
    if (!scene_intersect(sc, r, v, cn)) return miss_radiance(sc, r.d);                                   // :277
    V3 radiance = {0, 0, 0}, throughput = {1, 1, 1};
    for (int i = 0; i <= max_depth; ++i) {
        if (v.light != -1 && sc.lights[v.light].kind == TAKE_LIGHT_AREA) {                       // :284-291
            radiance = add(radiance, mulv(throughput, intensity(sc.lights[v.light])));
            break;
        }
        if (!rr_survives(sc, i, throughput, rng)) break;
        cn.shaded++;
        V3 dir_in = neg(r.d);
        const TakeMaterialDesc &m = sc.mats[v.material];
        bool spec = is_specular(m);
        if (sc.lights.size() > 0 && !spec && rng.next() <= 0.5) {                                 // :299
            int light_id = sample_light_power(sc, rng);
            const TakeLightDesc &l = sc.lights[light_id];
            if (l.kind == TAKE_LIGHT_AREA) {
                LightSample lp = sample_on_prim(sc, l.prim_id, v.pos, rng);
                double d = length(sub(lp.pos, v.pos));
                V3 light_dir = normalize(sub(lp.pos, v.pos));
                double lpdf = light_pdf_area(sc, light_id, lp.pos, v.pos) * (d * d) * sc.power_pmf[light_id] /
                              (fmax(dot(neg(lp.n), light_dir), 0.0));                             // :309
                if (lpdf <= 0) break;
                double bpdf = bsdf_pdf(m, dir_in, light_dir, v);
                if (bpdf <= 0) break;
                V3 FG = bsdf_eval(sc, m, dir_in, light_dir, 0.0, v);
                r = {v.pos, light_dir, EPS, INFINITY};
                Isect nv;
                if (!scene_intersect(sc, r, nv, cn)) {                                            // :326-330 (checked here, unlike :220)
                    radiance = add(radiance, mulv(throughput, miss_radiance(sc, light_dir)));
                    break;
                }
                v = nv;
                if (v.light == -1) break;                                                         // :332-334
                throughput = mulv(throughput, divs(FG, 0.5 * lpdf + 0.5 * bpdf));
            }
        } else {
            V3 rec_dir;
            double rec_pdf;
            if (!sample_bsdf(m, dir_in, v, rng, rec_dir, rec_pdf)) break;
            V3 FG = bsdf_eval(sc, m, dir_in, rec_dir, rec_pdf, v);
            V3 dir_out = normalize(rec_dir);
            double bpdf = rec_pdf;
            if (bpdf <= 0) break;
            r = {v.pos, dir_out, EPS, INFINITY};
            Isect nv;
            bool hit = scene_intersect(sc, r, nv, cn);
            double pdf = (sc.lights.empty() || spec) ? bpdf : 0.5 * bpdf;
            if (!hit) {
                throughput = mulv(throughput, divs(FG, pdf));
                radiance = add(radiance, mulv(throughput, miss_radiance(sc, dir_out)));
                break;
            }
            if (!spec && nv.light != -1) {                                                        // :363-373
                double d = length(sub(nv.pos, v.pos));
                V3 light_dir = normalize(sub(nv.pos, v.pos));
                double lpdf = light_pdf_area(sc, nv.light, nv.pos, v.pos) * (d * d) * sc.power_pmf[nv.light] /
                              fmax(dot(neg(nv.gn), light_dir), 0.0);
                if (lpdf <= 0) break;
                pdf += 0.5 * lpdf;
            }
            throughput = mulv(throughput, divs(FG, pdf));
            v = nv;
        }
    }
    return radiance;
}

typedef V3 (*Integrator)(const Scene &, Ray, Rng &, int, Counters &);
Integrator pick(int id) {
    switch (id) {
        case TAKE_INTEGRATOR_MIS: return path_tracing;
        case TAKE_INTEGRATOR_RAW: return path_tracing_raw;
        case TAKE_INTEGRATOR_ONE_SAMPLE_MIS: return path_tracing_one_sample_mis;
        case TAKE_INTEGRATOR_ONE_SAMPLE_MIS_POWER: return path_tracing_one_sample_mis_power;
    }
    return nullptr;
}

struct CameraBasis {
    V3 u, v, w;
    double vw, vh;
};
CameraBasis camera_basis(const TakeCamera &c) {  // src/render.cpp:37-44
    CameraBasis b;
    double theta = c.vfov / 180 * PI;
    double h = tan(theta / 2);
    b.vh = 2 * h;
    b.vw = b.vh / c.height * c.width;
    V3 from = {c.lookfrom[0], c.lookfrom[1], c.lookfrom[2]}, at = {c.lookat[0], c.lookat[1], c.lookat[2]};
    V3 up = {c.up[0], c.up[1], c.up[2]};
    b.w = normalize(sub(from, at));
    b.u = normalize(cross(up, b.w));
    b.v = cross(b.w, b.u);
    return b;
}

// One path sample; (x, y) in the reference's y-up pixel loop (src/render.cpp:65-77).
inline V3 one_sample(const Scene &sc, const CameraBasis &b, Integrator f, int x, int y, int64_t s, uint64_t seed,
                     int max_depth, Counters &cn) {
    const TakeCamera &c = sc.cam;
    Rng rng = {seed, (uint64_t)s, (uint32_t)((c.height - y - 1) * c.width + x), 0};
    double jx = rng.next();
    double jy = rng.next();
    V3 d = sub(add(mul(mul(b.u, (x + jx) / c.width - 0.5), b.vw), mul(mul(b.v, (y + jy) / c.height - 0.5), b.vh)), b.w);
    Ray r = {{c.lookfrom[0], c.lookfrom[1], c.lookfrom[2]}, normalize(d), EPS, INFINITY};
    return f(sc, r, rng, max_depth, cn);
}

template <typename F>
void run_threads(int nthreads, int64_t count, F body) {
    if (nthreads < 1) nthreads = 1;
    std::atomic<int64_t> next{0};
    auto worker = [&]() {
        for (;;) {
            int64_t i = next.fetch_add(1);
            if (i >= count) return;
            body(i);
        }
    };
    std::vector<std::thread> pool;
    for (int i = 1; i < nthreads; ++i) pool.emplace_back(worker);
    worker();
    for (auto &t : pool) t.join();
}

inline Ray make_ray(const double *r) { return {{r[0], r[1], r[2]}, {r[3], r[4], r[5]}, r[6], r[7]}; }

}  // namespace

// light_power (light.cpp:25-30) per light, summed in light order
void build_power_tables(Scene &sc) {
    const size_t n = sc.lights.size();
    if (n == 0) return;
    std::vector<double> power(n);
    double total = 0;
    for (size_t i = 0; i < n; ++i) {
        const TakeLightDesc &l = sc.lights[i];
        power[i] = l.kind == TAKE_LIGHT_AREA ? (l.intensity[0] * 0.212671 + l.intensity[1] * 0.715160 + l.intensity[2] * 0.072169) *
                                                   prim_area(sc, l.prim_id) * PI
                                             : 0.0;
        total += power[i];
    }
    sc.power_pmf.assign(n, 0.0);
    sc.power_cdf.assign(n + 1, 0.0);
    for (size_t i = 0; i < n; ++i) {
        sc.power_pmf[i] = power[i] / total;
        sc.power_cdf[i + 1] = sc.power_cdf[i] + sc.power_pmf[i];
    }
    sc.power_cdf[n] = 1;
}

extern "C" {

void oracle_set_russian_roulette(void *h, int rr_start) { ((Scene *)h)->rr_start = rr_start; }

void *oracle_scene_create(const TakeSceneDesc *d) {
    Scene *sc = new Scene;
    sc->cam = d->camera;
    sc->background = {d->background[0], d->background[1], d->background[2]};
    sc->pos.resize(d->num_vertices); sc->nrm.resize(d->num_vertices); sc->uv.resize(d->num_vertices);
    for (int64_t i = 0; i < d->num_vertices; ++i) {
        sc->pos[i] = {d->positions[3 * i], d->positions[3 * i + 1], d->positions[3 * i + 2]};
        sc->nrm[i] = {d->normals[3 * i], d->normals[3 * i + 1], d->normals[3 * i + 2]};
        sc->uv[i] = {d->uvs[2 * i], d->uvs[2 * i + 1]};
    }
    sc->idx.assign(d->indices, d->indices + 3 * d->num_prims);
    sc->pmat.assign(d->prim_material, d->prim_material + d->num_prims);
    sc->plight.assign(d->prim_light, d->prim_light + d->num_prims);
    sc->pflags.assign(d->prim_flags, d->prim_flags + d->num_prims);
    sc->spheres.assign(d->spheres, d->spheres + 4 * d->num_spheres);
    sc->mats.assign(d->materials, d->materials + d->num_materials);
    sc->lights.assign(d->lights, d->lights + d->num_lights);
    for (int i = 0; i < d->num_textures; ++i) {
        const TakeTextureDesc &t = d->textures[i];
        sc->tex.push_back({t.width, t.height, std::vector<double>(t.rgb, t.rgb + (size_t)t.width * t.height * 3)});
    }
    if (d->env_rgb && d->env_width > 0 && d->env_height > 0) {
        sc->env_w = d->env_width;
        sc->env_h = d->env_height;
        sc->env_sample = d->env_sample;
        sc->env_rgb.assign(d->env_rgb, d->env_rgb + (size_t)d->env_width * d->env_height * 3);
        build_env_tables(*sc);
    }
    build_bvh(*sc);
    build_power_tables(*sc);
    return sc;
}

void oracle_scene_free(void *h) { delete (Scene *)h; }

int64_t oracle_bvh_size(void *h) { return (int64_t)((Scene *)h)->nodes.size(); }
int oracle_bvh_root(void *h) { return ((Scene *)h)->root; }
void oracle_bvh_dump(void *h, double *box, int32_t *links) {
    const Scene &sc = *(Scene *)h;
    for (size_t i = 0; i < sc.nodes.size(); ++i) {
        const Node &n = sc.nodes[i];
        for (int a = 0; a < 3; ++a) { box[6 * i + a] = n.lo[a]; box[6 * i + 3 + a] = n.hi[a]; }
        links[3 * i] = n.left; links[3 * i + 1] = n.right; links[3 * i + 2] = n.prim;
    }
}

// rays: n x 8 doubles.  prim -1 on miss.  uv (optional): n x 2 barycentrics.  rec (optional): n x 16 doubles
// {pos3, geo_normal3, shading_normal3, uv2, t, material_id, area_light_id, 0, 0}.  counters (optional): int64[2] = box, tri tests.
void oracle_intersect(void *h, const double *rays, int64_t n, int32_t *prim, double *t, double *uv, double *rec,
                      int64_t *counters, int nthreads) {
    const Scene &sc = *(Scene *)h;
    std::atomic<int64_t> box{0}, tri{0};
    const int64_t chunk = 4096;
    run_threads(nthreads, (n + chunk - 1) / chunk, [&](int64_t c) {
        Counters cn;
        for (int64_t i = c * chunk; i < std::min(n, (c + 1) * chunk); ++i) {
            Ray r = make_ray(rays + 8 * i);
            Isect v;
            bool hit = scene_intersect(sc, r, v, cn);
            prim[i] = hit ? v.prim : -1;
            t[i] = hit ? v.t : 0.0;
            if (uv) { uv[2 * i] = hit ? v.bu : 0.0; uv[2 * i + 1] = hit ? v.bv : 0.0; }
            if (rec) {
                double *o = rec + 16 * i;
                memset(o, 0, 16 * sizeof(double));
                if (hit) {
                    o[0] = v.pos.x; o[1] = v.pos.y; o[2] = v.pos.z;
                    o[3] = v.gn.x; o[4] = v.gn.y; o[5] = v.gn.z;
                    o[6] = v.sn.x; o[7] = v.sn.y; o[8] = v.sn.z;
                    o[9] = v.uv.x; o[10] = v.uv.y; o[11] = v.t;
                    o[12] = v.material; o[13] = v.light;
                }
            }
        }
        box += cn.box; tri += cn.tri;
    });
    if (counters) { counters[0] = box; counters[1] = tri; }
}

void oracle_occluded(void *h, const double *rays, int64_t n, uint8_t *occ, int nthreads) {
    const Scene &sc = *(Scene *)h;
    const int64_t chunk = 4096;
    run_threads(nthreads, (n + chunk - 1) / chunk, [&](int64_t c) {
        Counters cn;
        for (int64_t i = c * chunk; i < std::min(n, (c + 1) * chunk); ++i)
            occ[i] = scene_occluded(sc, make_ray(rays + 8 * i), cn) ? 1 : 0;
    });
}

// Brute force over ALL primitives with the same leaf test: min t, exact ties -> the primitive visited last
// in the reference's DFS order (rank[] = position of the primitive in that order).  SURVEY.md 7.2 item 1.
void oracle_dfs_rank(void *h, int32_t *rank) {
    const Scene &sc = *(Scene *)h;
    if (sc.root < 0) return;
    std::vector<int> stack = {sc.root};
    int next = 0;
    while (!stack.empty()) {
        int id = stack.back();
        stack.pop_back();
        const Node &n = sc.nodes[id];
        if (n.prim != -1) { rank[n.prim] = next++; continue; }
        stack.push_back(n.right);
        stack.push_back(n.left);
    }
}

// stats (optional): int64[6] = extend rays, shadow rays, shaded vertices, box tests, leaf tests, misses after a light sample
// Only image rows r (from the top) with r % row_step == row_begin are rendered (row_step = 1: all rows).
int oracle_render(void *h, int integrator, int max_depth, int64_t spp_begin, int64_t spp_end, uint64_t seed, int nthreads,
                  double *sum, double *sumsq, int64_t *stats, int row_begin, int row_step) {
    const Scene &sc = *(Scene *)h;
    Integrator f = pick(integrator);
    if (!f) return -1;
    CameraBasis b = camera_basis(sc.cam);
    int W = sc.cam.width, H = sc.cam.height;
    std::vector<Counters> per_row(H);
    run_threads(nthreads, H, [&](int64_t y) {
        Counters cn;
        if (row_step > 1 && (H - y - 1) % row_step != row_begin) return;
        for (int x = 0; x < W; ++x) {
            V3 acc = {0, 0, 0}, acc2 = {0, 0, 0};
            for (int64_t s = spp_begin; s < spp_end; ++s) {
                V3 c = one_sample(sc, b, f, x, (int)y, s, seed, max_depth, cn);
                acc = add(acc, c);
                acc2 = add(acc2, mulv(c, c));
            }
            size_t o = 3 * ((size_t)(H - y - 1) * W + x);
            sum[o] = acc.x; sum[o + 1] = acc.y; sum[o + 2] = acc.z;
            if (sumsq) { sumsq[o] = acc2.x; sumsq[o + 1] = acc2.y; sumsq[o + 2] = acc2.z; }
        }
        per_row[y] = cn;
    });
    if (stats) {
        Counters t;
        for (auto &c : per_row) t += c;
        stats[0] = t.extend; stats[1] = t.shadow; stats[2] = t.shaded; stats[3] = t.box; stats[4] = t.tri;
        stats[5] = t.miss_after_light_sample;
    }
    return 0;
}

int oracle_radiance_samples(void *h, int integrator, int max_depth, uint64_t seed, int64_t n, const int32_t *px,
                            const int32_t *py, const int64_t *s, double *out, int nthreads) {
    const Scene &sc = *(Scene *)h;
    Integrator f = pick(integrator);
    if (!f) return -1;
    CameraBasis b = camera_basis(sc.cam);
    const int64_t chunk = 256;
    run_threads(nthreads, (n + chunk - 1) / chunk, [&](int64_t c) {
        Counters cn;
        for (int64_t i = c * chunk; i < std::min(n, (c + 1) * chunk); ++i) {
            V3 v = one_sample(sc, b, f, px[i], sc.cam.height - 1 - py[i], s[i], seed, max_depth, cn);
            out[3 * i] = v.x; out[3 * i + 1] = v.y; out[3 * i + 2] = v.z;
        }
    });
    return 0;
}

// Primary camera rays exactly as one_sample() shoots them (n x 8 doubles), for intersection parity tests.
void oracle_primary_rays(void *h, uint64_t seed, int64_t n, const int32_t *px, const int32_t *py, const int64_t *s,
                         int jitter, double *rays) {
    const Scene &sc = *(Scene *)h;
    const TakeCamera &c = sc.cam;
    CameraBasis b = camera_basis(c);
    for (int64_t i = 0; i < n; ++i) {
        int x = px[i], y = c.height - 1 - py[i];
        Rng rng = {seed, (uint64_t)s[i], (uint32_t)(py[i] * c.width + x), 0};
        double jx = jitter ? rng.next() : 0.5;
        double jy = jitter ? rng.next() : 0.5;
        V3 d = normalize(sub(add(mul(mul(b.u, (x + jx) / c.width - 0.5), b.vw), mul(mul(b.v, (y + jy) / c.height - 0.5), b.vh)), b.w));
        double *o = rays + 8 * i;
        o[0] = c.lookfrom[0]; o[1] = c.lookfrom[1]; o[2] = c.lookfrom[2];
        o[3] = d.x; o[4] = d.y; o[5] = d.z; o[6] = EPS; o[7] = INFINITY;
    }
}

// EXTENSION test hooks: environment sampling.  out = {dir3, pdf}; pdf_of = pdf of an arbitrary direction; rad = radiance.
void oracle_env_sample(void *h, double u1, double u2, double *out) {
    const Scene &sc = *(Scene *)h;
    V3 d;
    double pdf;
    env_sample_dir(sc, u1, u2, d, pdf);
    out[0] = d.x; out[1] = d.y; out[2] = d.z; out[3] = pdf;
}
void oracle_env_eval(void *h, const double *dir, double *out) {
    const Scene &sc = *(Scene *)h;
    V3 d = {dir[0], dir[1], dir[2]};
    V3 r = env_radiance(sc, d);
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = env_pdf(sc, d);
}

double oracle_stream_real(uint64_t seed, uint32_t pixel, uint64_t sample, uint32_t k) {
    return take_stream_real(seed, pixel, sample, k);
}
void oracle_philox(const uint32_t *ctr, const uint32_t *key, uint32_t *out) { take_philox4x32_10(ctr, key, out); }

// ---- output step (TEST CHECKER for take_gpu_exr_pack*): src/render.cpp:78, src/image.cpp:157-161 and, from the
// reference's vendored tinyexr.h, float_to_half_full (:889-924), the B,G,R planar scanline layout of SaveEXR / EncodeChunk
// and the pre-filter of CompressZip (:1212-1258).  out = width*height*6 bytes, block after block.
static uint16_t exr_half(float f32) {
    uint32_t f;
    memcpy(&f, &f32, 4);
    const uint32_t sign = f >> 31, e = (f >> 23) & 0xffu, m = f & 0x7fffffu;
    uint32_t o = 0;
    if (e == 0) {
        o = 0;
    } else if (e == 255) {
        o = (31u << 10) | (m ? 0x200u : 0u);
    } else {
        const int ne = (int)e - 127 + 15;
        if (ne >= 31) {
            o = 31u << 10;
        } else if (ne <= 0) {
            if ((14 - ne) <= 24) {
                const uint32_t mant = m | 0x800000u;
                o = mant >> (14 - ne);
                if ((mant >> (13 - ne)) & 1u) o++;
            }
        } else {
            o = ((uint32_t)ne << 10) | (m >> 13);
            if (m & 0x1000u) o++;
        }
    }
    return (uint16_t)((o & 0x7fffu) | (sign << 15));
}

void oracle_exr_pack(int width, int height, const double *sum_rgb, int64_t spp, uint8_t *out) {
    const double inv = 1.0 / (double)spp;
    const size_t line_bytes = (size_t)width * 6;
    std::vector<uint8_t> raw, tmp;
    for (int y0 = 0; y0 < height; y0 += 16) {
        const int lines = std::min(16, height - y0);
        const size_t n = lines * line_bytes;
        raw.assign(n, 0);
        for (int l = 0; l < lines; ++l)
            for (int c = 0; c < 3; ++c)  // planes B, G, R
                for (int x = 0; x < width; ++x) {
                    const double mean = sum_rgb[3 * ((size_t)(y0 + l) * width + x) + (2 - c)] * inv;
                    const uint16_t h = exr_half((float)mean);
                    uint8_t *p = raw.data() + l * line_bytes + (size_t)c * width * 2 + (size_t)x * 2;
                    p[0] = (uint8_t)(h & 0xff); p[1] = (uint8_t)(h >> 8);
                }
        tmp.assign(n, 0);
        {   // reorder
            size_t t1 = 0, t2 = (n + 1) / 2, i = 0;
            for (;;) {
                if (i < n) tmp[t1++] = raw[i++]; else break;
                if (i < n) tmp[t2++] = raw[i++]; else break;
            }
        }
        {   // predictor
            int p = tmp[0];
            for (size_t i = 1; i < n; ++i) {
                int d = (int)tmp[i] - p + (128 + 256);
                p = tmp[i];
                tmp[i] = (uint8_t)d;
            }
        }
        memcpy(out + (size_t)y0 * line_bytes, tmp.data(), n);
    }
}

}  // extern "C"
