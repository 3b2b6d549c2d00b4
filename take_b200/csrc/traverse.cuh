// Ray-scene intersection on the device.
//
//  trace_exact  -- the reference's traversal verbatim: its own tree topology, FP64 slab test with divisions and
//                  fmin/fmax (src/bbox.h:18-32), left-then-right DFS with tmax shrinking after each accepted hit
//                  and later hits replacing earlier ones (src/bvh.cpp:86-109), unrolled into an explicit stack.
//                  Bit-identical (id and t) to the reference on every ray, including exact ties.
//  trace_fast   -- production path: SAH BVH2 with 64-byte nodes (both children's boxes, four float4 fetches),
//                  conservative FP32 slab tests, near-child-first ordering, per-thread stack in shared memory
//                  (conflict-free column layout) spilling to local memory, and the SAME FP64 leaf test.  Any
//                  traversal that never culls a primitive the FP64 leaf test would accept returns the same closest
//                  hit; equal-t candidates are resolved by the reference's DFS rank.
//
// Only trace_exact and the 4-wide trace_fast4 are compiled into the shipped library.  The variants that were measured and
// lost (the binary-node trace_fast, the speculative trace_spec4, the warp-persistent loop with lane
// refill) are kept for A/B runs behind -DTAKE_EXPERIMENTAL=1 (make EXTRA=-DTAKE_EXPERIMENTAL=1).
#pragma once
#include "device_common.cuh"

namespace take {

#ifndef TAKE_EXPERIMENTAL
#define TAKE_EXPERIMENTAL 0
#endif
#ifndef TAKE_STACK_SMEM
#define TAKE_STACK_SMEM 24   // entries per thread kept in shared memory
#endif
#define TAKE_STACK_LOCAL (96 - TAKE_STACK_SMEM)  // overflow entries in local memory (tree depth limit = 96, checked on the host)
#ifndef TAKE_PIN_SBASE
#define TAKE_PIN_SBASE 1
#endif
#ifndef TAKE_PUSH_ALWAYS
#define TAKE_PUSH_ALWAYS 0
#endif
#define TAKE_STACK_SMEM_ALLOC (TAKE_STACK_SMEM > 0 ? TAKE_STACK_SMEM : 1)
#ifndef TAKE_ANYHIT_UNSORTED
#define TAKE_ANYHIT_UNSORTED 1
#endif

// (256-bit read-only loads ldg_f8 / ldg_d4: device_common.cuh)

// ---- leaf tests: src/shape.cpp:44-78 (triangle) and :13-29 (sphere), accept/reject part ---------------
__device__ __forceinline__ bool hit_triangle(D3 v0, D3 e1, D3 e2, D3 o, D3 d, double tmin, double tmax, double &t,
                                             double &bu, double &bv) {
    D3 h = cross(d, e2);
    double a = dot(e1, h);
    if (a > -TAKE_EPS && a < TAKE_EPS) return false;
    double f = 1.0 / a;
    D3 s = sub(o, v0);
    double u = f * dot(s, h);
    if (u < 0.0 || u > 1.0) return false;
    D3 q = cross(s, e1);
    double v = f * dot(d, q);
    if (v < 0.0 || u + v > 1.0) return false;
    double tt = f * dot(e2, q);
    if (tt < tmin || tmax < tt) return false;
    t = tt; bu = u; bv = v;
    return true;
}

__device__ __forceinline__ bool hit_sphere(D3 c, double radius, D3 o, D3 d, double tmin, double tmax, double &t) {
    D3 oc = sub(o, c);
    double a = dot(d, d);
    double half_b = dot(oc, d);
    double cc = dot(oc, oc) - radius * radius;
    double disc = half_b * half_b - a * cc;
    if (disc < 0) return false;
    double sqrtd = sqrt(disc);
    double root = (-half_b - sqrtd) / a;
    if (root < tmin || tmax < root) {
        root = (-half_b + sqrtd) / a;
        if (root < tmin || tmax < root) return false;
    }
    t = root;
    return true;
}

// Leaf test by primitive id, from the reference-layout arrays (used by the exact path).
__device__ __forceinline__ bool hit_prim(const DevScene &sc, int prim, D3 o, D3 d, double tmin, double tmax, double &t,
                                         double &bu, double &bv) {
    bu = bv = 0;
    if (sc.prim_flags[prim] & TAKE_PRIM_SPHERE) {
        const double *s = sc.spheres + 4 * (int64_t)sc.indices[3 * (int64_t)prim];
        return hit_sphere(mk3(s[0], s[1], s[2]), s[3], o, d, tmin, tmax, t);
    }
    const int32_t *id = sc.indices + 3 * (int64_t)prim;
    const double *p0 = sc.positions + 3 * (int64_t)id[0], *p1 = sc.positions + 3 * (int64_t)id[1],
                 *p2 = sc.positions + 3 * (int64_t)id[2];
    D3 v0 = mk3(p0[0], p0[1], p0[2]);
    D3 e1 = sub(mk3(p1[0], p1[1], p1[2]), v0), e2 = sub(mk3(p2[0], p2[1], p2[2]), v0);
    return hit_triangle(v0, e1, e2, o, d, tmin, tmax, t, bu, bv);
}

// ---- exact mode ---------------------------------------------------------------------------------------
__device__ __forceinline__ bool slab_exact(const RefNode &b, D3 o, D3 d, double tmin, double tmax) {  // bbox.h:18-32
    double ta = (b.lo[0] - o.x) / d.x, tb = (b.hi[0] - o.x) / d.x;
    double t_min = fmax(fmin(ta, tb), tmin), t_max = fmin(fmax(ta, tb), tmax);
    if (t_max < t_min) return false;
    ta = (b.lo[1] - o.y) / d.y; tb = (b.hi[1] - o.y) / d.y;
    t_min = fmax(fmin(ta, tb), t_min); t_max = fmin(fmax(ta, tb), t_max);
    if (t_max < t_min) return false;
    ta = (b.lo[2] - o.z) / d.z; tb = (b.hi[2] - o.z) / d.z;
    t_min = fmax(fmin(ta, tb), t_min); t_max = fmin(fmax(ta, tb), t_max);
    return !(t_max < t_min);
}

// The recursion of bvh.cpp:86-109 is equivalent to: visit nodes in left-to-right DFS order; test a node's box when
// it is reached (with the tmax current at that moment); at a leaf, accept a hit with tmin <= t <= tmax, which then
// becomes both the result and the new tmax.  The root's own box is never tested.
__device__ inline void trace_exact(const DevScene &sc, D3 o, D3 d, double tmin, double tmax, HitOut &out) {
    out.prim = -1; out.rank = -1; out.t = 0; out.u = 0; out.v = 0;
    if (sc.ref_root < 0) return;
    int32_t stack[128];
    int sp = 0;
    int32_t node = sc.ref_root;
    bool first = true;
    for (;;) {
        const RefNode &n = sc.ref_nodes[node];
        if (first || slab_exact(n, o, d, tmin, tmax)) {
            first = false;
            if (n.prim != -1) {
                double t, bu, bv;
                if (hit_prim(sc, n.prim, o, d, tmin, tmax, t, bu, bv)) {
                    out.prim = n.prim; out.t = t; out.u = bu; out.v = bv;
                    tmax = t;
                }
            } else {
                stack[sp++] = n.right;
                node = n.left;
                continue;
            }
        }
        if (sp == 0) break;
        node = stack[--sp];
    }
}

// ---- fast mode ------------------------------------------------------------------------------------------
// Per-thread traversal stack of (node, entry distance) pairs.  The first TAKE_STACK_SMEM levels live in shared memory
// (column layout entry[level][thread]: one 64-bit access per lane, conflict-free), deeper levels in a local-memory
// array owned by the kernel.  The struct itself holds only scalars -- the shared-window address of the column, the
// overflow pointer and the depth -- so that after inlining all of it stays in registers (an earlier version embedded
// the overflow array, which pinned every field, `sp` included, in local memory: ~25 instructions and three dependent
// local loads per push).
#define TAKE_TRACE_BLOCK 128  // threads per block of every traversal kernel
struct TravStack {
    uint32_t sbase;  // shared-window byte address of entry[0][this thread]
    uint2 *l;        // overflow levels (local memory)
    int sp;
    __device__ __forceinline__ void init(uint2 *smem_block, uint2 *local_levels) {
        sbase = (uint32_t)__cvta_generic_to_shared(smem_block + threadIdx.x);
        l = local_levels;
        sp = 0;
#if TAKE_PIN_SBASE
        asm volatile("" : "+r"(sbase));  // opaque: keeps the address in a register instead of re-deriving it (S2R) per push
#endif
    }
    // Store at the top without moving it (callers bump `sp` by a predicate: a branch-free conditional push; the slot
    // above the top may hold garbage, which is never read).
    __device__ __forceinline__ void put_bits(int32_t node, uint32_t tn_bits) {
        if (sp < TAKE_STACK_SMEM)
            asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(sbase + (uint32_t)sp * (TAKE_TRACE_BLOCK * 8u)), "r"(node), "r"(tn_bits) : "memory");
        else
            l[sp - TAKE_STACK_SMEM] = make_uint2((uint32_t)node, tn_bits);
    }
    __device__ __forceinline__ void push_bits(int32_t node, uint32_t tn_bits) {
        if (sp < TAKE_STACK_SMEM)
            asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(sbase + (uint32_t)sp * (TAKE_TRACE_BLOCK * 8u)), "r"(node), "r"(tn_bits) : "memory");
        else
            l[sp - TAKE_STACK_SMEM] = make_uint2((uint32_t)node, tn_bits);
        ++sp;
    }
    __device__ __forceinline__ void push(int32_t node, float tn) { push_bits(node, __float_as_uint(tn)); }
    __device__ __forceinline__ void pop(int32_t &node, float &tn) {
        --sp;
        uint32_t a, b;
        if (sp < TAKE_STACK_SMEM) {
            asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "r"(sbase + (uint32_t)sp * (TAKE_TRACE_BLOCK * 8u)) : "memory");
        } else {
            const uint2 e = l[sp - TAKE_STACK_SMEM];
            a = e.x; b = e.y;
        }
        node = (int32_t)a;
        tn = __uint_as_float(b);
    }
};
// Declares the storage of a traversal stack inside a kernel of TAKE_TRACE_BLOCK threads and initialises `st`.
#define TAKE_DECLARE_STACK(st)                                                  \
    __shared__ uint2 st##_smem[TAKE_STACK_SMEM_ALLOC * TAKE_TRACE_BLOCK];       \
    uint2 st##_local[TAKE_STACK_LOCAL + 1];                                      \
    TravStack st;                                                               \
    st.init(st##_smem, st##_local)

// Reciprocal direction for the FMA slab form  t = plane * idir - (o -+ delta) * idir.  An infinite idir (zero direction
// component) would turn that into inf - inf = NaN for exactly the slabs that contain the origin; clamping |idir| to
// 1e18 keeps every product finite (|coordinate| * 1e18 << FLT_MAX), gives the correct signs (the delta padding makes
// "origin inside the slab" robustly negative/positive), and still acts as infinity: the smallest exit distance it can
// produce, delta * 1e18 >= 1e12 scene extents, is beyond any hit inside the scene bounds.
__device__ __forceinline__ float safe_rcp(float d) { return fminf(fmaxf(1.0f / d, -1e18f), 1e18f); }

#define TAKE_SLACK 1.00000191f  // 1 + 2^-19: relative slack on the exit distance (error analysis in DESIGN.md)

#define TAKE_NODE_DONE ((int32_t)0x80000000)  // "no node left" (never a valid leaf link: ~0x7fffffff is the empty-child code)

struct TravCounters {
    unsigned long long box, tri;
};

#if TAKE_EXPERIMENTAL
template <bool ANY_HIT, bool COUNT>
__device__ __forceinline__ void trace_fast(const DevScene &sc, D3 o, D3 d, double tmin, double tmax, TravStack &st,
                                           HitOut &out, TravCounters *cnt) {
    out.prim = -1; out.rank = -1; out.t = 0; out.u = 0; out.v = 0;
    if (sc.num_prims <= 0) return;
    // FP32 image of the ray.  Every box plane is tested as  t = plane * idir - (o -+ delta) * idir  with one FMA;
    // delta (absolute) covers the rounding of the origin, of the product (o*idir) and of the FMA itself.
    const float ox = (float)o.x, oy = (float)o.y, oz = (float)o.z;
    const float idx = safe_rcp((float)d.x), idy = safe_rcp((float)d.y), idz = safe_rcp((float)d.z);
    const float delta = 1.9073486e-6f * fmaxf(fmaxf(fabsf(ox), fabsf(oy)), fmaxf(fabsf(oz), sc.abs_max));  // 2^-19
    const float olx = -(ox + delta) * idx, ohx = -(ox - delta) * idx;
    const float oly = -(oy + delta) * idy, ohy = -(oy - delta) * idy;
    const float olz = -(oz + delta) * idz, ohz = -(oz - delta) * idz;
    const float tmin_f = __double2float_rd(tmin);
    float tbest_f = __double2float_ru(tmax);
    double best_t = tmax;

    st.sp = 0;
    int32_t node = 0;  // root
    for (;;) {
        while (node >= 0) {
            const F8 n01 = ldg_f8(sc.nodes + 4 * (int64_t)node), n23 = ldg_f8(sc.nodes + 4 * (int64_t)node + 2);
            const float4 q0 = n01.a, q1 = n01.b, q2 = n23.a, q3 = n23.b;
            if (COUNT) cnt->box += 2;
            float a, b;
            a = fmaf(q0.x, idx, olx); b = fmaf(q0.y, idx, ohx);
            float tn0 = fminf(a, b), tf0 = fmaxf(a, b);
            a = fmaf(q0.z, idy, oly); b = fmaf(q0.w, idy, ohy);
            tn0 = fmaxf(tn0, fminf(a, b)); tf0 = fminf(tf0, fmaxf(a, b));
            a = fmaf(q2.x, idz, olz); b = fmaf(q2.y, idz, ohz);
            tn0 = fmaxf(tn0, fminf(a, b)); tf0 = fminf(tf0, fmaxf(a, b));
            tn0 = fmaxf(tn0, tmin_f); tf0 = fminf(tf0, tbest_f);
            a = fmaf(q1.x, idx, olx); b = fmaf(q1.y, idx, ohx);
            float tn1 = fminf(a, b), tf1 = fmaxf(a, b);
            a = fmaf(q1.z, idy, oly); b = fmaf(q1.w, idy, ohy);
            tn1 = fmaxf(tn1, fminf(a, b)); tf1 = fminf(tf1, fmaxf(a, b));
            a = fmaf(q2.z, idz, olz); b = fmaf(q2.w, idz, ohz);
            tn1 = fmaxf(tn1, fminf(a, b)); tf1 = fminf(tf1, fmaxf(a, b));
            tn1 = fmaxf(tn1, tmin_f); tf1 = fminf(tf1, tbest_f);
            const bool h0 = tn0 <= tf0 * TAKE_SLACK, h1 = tn1 <= tf1 * TAKE_SLACK;
            const int32_t c0 = __float_as_int(q3.x), c1 = __float_as_int(q3.y);
            if (h0 && h1) {
                if (tn1 < tn0) { st.push(c0, tn0); node = c1; }
                else { st.push(c1, tn1); node = c0; }
            } else if (h0) {
                node = c0;
            } else if (h1) {
                node = c1;
            } else {
                // pop, skipping entries that the current best hit has made unreachable
                for (;;) {
                    if (st.sp == 0) return;
                    float tn;
                    st.pop(node, tn);
                    if (tn <= tbest_f * TAKE_SLACK) break;
                }
            }
        }
        // leaf: ~node = (first slot << 3) | (count - 1)
        {
            const int32_t code = ~node;
            const int64_t first = code >> 3;
            const int count = (code & 7) + 1;
            for (int k = 0; k < count; ++k) {
                const double2 *T = sc.tris + 6 * (first + k);
                const D4 t01 = ldg_d4(T), t23 = ldg_d4(T + 2), t45 = ldg_d4(T + 4);
                const double2 a0 = t01.a, a1 = t01.b, a2 = t23.a, a3 = t23.b, a4 = t45.a, a5 = t45.b;
                if (COUNT) cnt->tri += 1;
                double t, bu = 0, bv = 0;
                bool ok;
                if (a5.y == 0.0) {
                    ok = hit_triangle(mk3(a0.x, a0.y, a1.x), mk3(a2.x, a2.y, a3.x), mk3(a4.x, a4.y, a5.x), o, d, tmin, best_t,
                                      t, bu, bv);
                } else {
                    ok = hit_sphere(mk3(a0.x, a0.y, a1.x), a3.y, o, d, tmin, best_t, t);
                }
                if (ok) {
                    const long long bits = __double_as_longlong(a1.y);
                    const int32_t prim = (int32_t)(bits & 0xffffffffLL), rank = (int32_t)(bits >> 32);
                    // min t; exact ties go to the primitive the reference's DFS reaches last (bvh.cpp:94-108)
                    if (t < best_t || out.prim < 0 || rank > out.rank) {
                        out.prim = prim; out.rank = rank; out.t = t; out.u = bu; out.v = bv;
                        best_t = t;
                        tbest_f = __double2float_ru(t);
                        if (ANY_HIT) return;
                    }
                }
            }
        }
        for (;;) {
            if (st.sp == 0) return;
            float tn;
            st.pop(node, tn);
            if (tn <= tbest_f * TAKE_SLACK) break;
        }
    }
}


#endif  // TAKE_EXPERIMENTAL (binary-node traversal)

// ---------------------------------------------------------------------------------------------------------
// 4-wide variant of trace_fast: same conservative slab arithmetic, same FP64 leaf test and tie rule, on 128-byte
// nodes that hold four children (bvh_build.h: WideNode).  Half as many dependent node fetches per ray; the up to
// four hits of a node are ordered near-to-far with a 5-comparator network on (distance bits | child slot) keys.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cswap(uint32_t &a, uint32_t &b) {
    const uint32_t lo = min(a, b), hi = max(a, b);
    a = lo; b = hi;
}

// TIES: count the accepted leaf tests whose t equals the current hit's (DevScene::tie_count) -- only compiled into the kernels
// of a render that runs ahead of the tie-break ranks (take_gpu.cu); elsewhere it would cost registers the kernels do not have.
template <bool ANY_HIT, bool COUNT, bool TIES = false>
__device__ __forceinline__ void trace_fast4(const DevScene &sc, D3 o, D3 d, double tmin, double tmax, TravStack &st,
                                            HitOut &out, TravCounters *cnt) {
    out.prim = -1; out.rank = -1; out.t = 0; out.u = 0; out.v = 0;
    if (sc.num_prims <= 0) return;
    const float ox = (float)o.x, oy = (float)o.y, oz = (float)o.z;
    const float idx = safe_rcp((float)d.x), idy = safe_rcp((float)d.y), idz = safe_rcp((float)d.z);
    const float delta = 1.9073486e-6f * fmaxf(fmaxf(fabsf(ox), fabsf(oy)), fmaxf(fabsf(oz), sc.abs_max));  // 2^-19
    const float olx = -(ox + delta) * idx, ohx = -(ox - delta) * idx;
    const float oly = -(oy + delta) * idy, ohy = -(oy - delta) * idy;
    const float olz = -(oz + delta) * idz, ohz = -(oz - delta) * idz;
    const float tmin_f = __double2float_rd(tmin);
    float tbest_f = __double2float_ru(tmax);
    double best_t = tmax;

    st.sp = 0;
    int32_t node = 0;  // root
    for (;;) {
        while (node >= 0) {
            const float4 *N = sc.wide_nodes + 8 * (int64_t)node;
            const F8 nx = ldg_f8(N), ny = ldg_f8(N + 2), nz = ldg_f8(N + 4);
            const float4 lox = nx.a, hix = nx.b, loy = ny.a, hiy = ny.b, loz = nz.a, hiz = nz.b;
            const int4 ch = __ldg((const int4 *)(N + 6));
            if (COUNT) cnt->box += 4;
            uint32_t key[4];
#define TAKE_WIDE_CHILD(K, LX, HX, LY, HY, LZ, HZ, C)                                            \
            {                                                                                      \
                float a = fmaf(LX, idx, olx), b = fmaf(HX, idx, ohx);                              \
                float tn = fminf(a, b), tf = fmaxf(a, b);                                          \
                a = fmaf(LY, idy, oly); b = fmaf(HY, idy, ohy);                                    \
                tn = fmaxf(tn, fminf(a, b)); tf = fminf(tf, fmaxf(a, b));                          \
                a = fmaf(LZ, idz, olz); b = fmaf(HZ, idz, ohz);                                    \
                tn = fmaxf(tn, fminf(a, b)); tf = fminf(tf, fmaxf(a, b));                          \
                tn = fmaxf(tn, tmin_f); tf = fminf(tf, tbest_f);                                   \
                const bool h = (tn <= tf * TAKE_SLACK) && (C != TAKE_WIDE_EMPTY);                  \
                key[K] = h ? ((__float_as_uint(tn) & 0xfffffffcu) | (uint32_t)K) : 0xffffffffu;    \
            }
            TAKE_WIDE_CHILD(0, lox.x, hix.x, loy.x, hiy.x, loz.x, hiz.x, ch.x)
            TAKE_WIDE_CHILD(1, lox.y, hix.y, loy.y, hiy.y, loz.y, hiz.y, ch.y)
            TAKE_WIDE_CHILD(2, lox.z, hix.z, loy.z, hiy.z, loz.z, hiz.z, ch.z)
            TAKE_WIDE_CHILD(3, lox.w, hix.w, loy.w, hiy.w, loz.w, hiz.w, ch.w)
#undef TAKE_WIDE_CHILD
            // sort the four keys ascending: misses (0xffffffff) sink to the end, hits come out near-to-far.  An any-hit
            // query has no use for the order (TAKE_ANYHIT_UNSORTED): its window never shrinks, so every child the ray
            // overlaps is visited unless an occluder ends the search, whichever comes first.
            if (!(ANY_HIT && TAKE_ANYHIT_UNSORTED)) {
                cswap(key[0], key[1]); cswap(key[2], key[3]); cswap(key[0], key[2]); cswap(key[1], key[3]); cswap(key[1], key[2]);
            }
            // child link of the slot in a key's low two bits: three selects, no branches
#define TAKE_WIDE_PICK(KEY) (((KEY) & 2u) ? (((KEY) & 1u) ? ch.w : ch.z) : (((KEY) & 1u) ? ch.y : ch.x))
#if TAKE_PUSH_ALWAYS
            st.put_bits(TAKE_WIDE_PICK(key[3]), key[3] & 0xfffffffcu); st.sp += (key[3] != 0xffffffffu);
            st.put_bits(TAKE_WIDE_PICK(key[2]), key[2] & 0xfffffffcu); st.sp += (key[2] != 0xffffffffu);
            st.put_bits(TAKE_WIDE_PICK(key[1]), key[1] & 0xfffffffcu); st.sp += (key[1] != 0xffffffffu);
#else
            if (key[3] != 0xffffffffu) st.push_bits(TAKE_WIDE_PICK(key[3]), key[3] & 0xfffffffcu);
            if (key[2] != 0xffffffffu) st.push_bits(TAKE_WIDE_PICK(key[2]), key[2] & 0xfffffffcu);
            if (key[1] != 0xffffffffu) st.push_bits(TAKE_WIDE_PICK(key[1]), key[1] & 0xfffffffcu);
#endif
            if (key[0] != 0xffffffffu) {
                node = TAKE_WIDE_PICK(key[0]);
            } else {
                for (;;) {
                    if (st.sp == 0) return;
                    float tn;
                    st.pop(node, tn);
                    if ((ANY_HIT && TAKE_ANYHIT_UNSORTED) || tn <= tbest_f * TAKE_SLACK) break;
                }
            }
#undef TAKE_WIDE_PICK
        }
        {
            const int32_t code = ~node;
            const int64_t first = code >> 3;
            const int count = (code & 7) + 1;
            for (int k = 0; k < count; ++k) {
                const double2 *T = sc.tris + 6 * (first + k);
                const D4 t01 = ldg_d4(T), t23 = ldg_d4(T + 2), t45 = ldg_d4(T + 4);
                const double2 a0 = t01.a, a1 = t01.b, a2 = t23.a, a3 = t23.b, a4 = t45.a, a5 = t45.b;
                if (COUNT) cnt->tri += 1;
                double t, bu = 0, bv = 0;
                bool ok;
                if (a5.y == 0.0)
                    ok = hit_triangle(mk3(a0.x, a0.y, a1.x), mk3(a2.x, a2.y, a3.x), mk3(a4.x, a4.y, a5.x), o, d, tmin, best_t, t,
                                      bu, bv);
                else
                    ok = hit_sphere(mk3(a0.x, a0.y, a1.x), a3.y, o, d, tmin, best_t, t);
                if (ok) {
                    const long long bits = __double_as_longlong(a1.y);
                    const int32_t prim = (int32_t)(bits & 0xffffffffLL), rank = (int32_t)(bits >> 32);
                    // (an accepted t equal to the current hit's: the only place the rank decides anything -- counted, see DevScene)
                    if (TIES && !ANY_HIT && out.prim >= 0 && !(t < best_t)) atomicAdd(sc.tie_count, 1ULL);
                    if (t < best_t || out.prim < 0 || rank > out.rank) {
                        out.prim = prim; out.rank = rank; out.t = t; out.u = bu; out.v = bv;
                        best_t = t;
                        tbest_f = __double2float_ru(t);
                        if (ANY_HIT) return;
                    }
                }
            }
        }
        for (;;) {
            if (st.sp == 0) return;
            float tn;
            st.pop(node, tn);
            if ((ANY_HIT && TAKE_ANYHIT_UNSORTED) || tn <= tbest_f * TAKE_SLACK) break;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Packet traversal for camera rays: the 32 rays of a warp walk the tree TOGETHER.
//
// Pass 0 of a wave gives every warp the samples of one pixel (or of an 8x4 pixel tile): rays that differ by sub-pixel
// jitter.  Traced one ray per thread they visit almost the same nodes, yet each lane keeps its own stack and its own node
// pointer, the warp runs until its longest lane is done (21.7 of 32 lanes active per instruction in the ncu capture of
// config 2) and every node is requested by up to 32 lanes.  Here the warp has ONE node pointer and ONE stack (in shared
// memory): a node is loaded once (same address in all lanes: a broadcast), every lane tests the four children against its
// own ray, a child is entered when ANY lane enters it, the children are ordered by the nearest entry distance over the
// lanes (redux), and a stacked entry is dropped when it is beyond every lane's current hit.  Every lane therefore sees a
// superset of the leaves its own traversal would have seen and applies the same FP64 leaf test and tie rule, whose
// outcome does not depend on the visiting order: the result is bit-identical to trace_fast4.
// `valid` = false marks a lane without a ray (it takes part in the warp-wide operations with an empty interval).
// ---------------------------------------------------------------------------------------------------------
#define TAKE_PACKET_STACK 64   // entries of the shared per-warp stack (3 per level of the 4-wide tree + 1; checked on the host)

template <bool COUNT, bool TIES = false>
__device__ __forceinline__ void trace_packet4(const DevScene &sc, D3 o, D3 d, double tmin, double tmax, bool valid, uint2 *wstack,
                                              HitOut &out, TravCounters *cnt) {
    out.prim = -1; out.rank = -1; out.t = 0; out.u = 0; out.v = 0;
    if (sc.num_prims <= 0) return;
    if (!valid) tmax = -1.0;   // empty interval: no box and no primitive can be hit
    const float ox = (float)o.x, oy = (float)o.y, oz = (float)o.z;
    const float idx = safe_rcp((float)d.x), idy = safe_rcp((float)d.y), idz = safe_rcp((float)d.z);
    const float delta = 1.9073486e-6f * fmaxf(fmaxf(fabsf(ox), fabsf(oy)), fmaxf(fabsf(oz), sc.abs_max));  // 2^-19
    const float olx = -(ox + delta) * idx, ohx = -(ox - delta) * idx;
    const float oly = -(oy + delta) * idy, ohy = -(oy - delta) * idy;
    const float olz = -(oz + delta) * idz, ohz = -(oz - delta) * idz;
    const float tmin_f = __double2float_rd(tmin);
    float tbest_f = __double2float_ru(tmax);
    double best_t = tmax;
    int sp = 0;          // warp-uniform
    int32_t node = 0;    // warp-uniform
    for (;;) {
        while (node >= 0) {
            const float4 *N = sc.wide_nodes + 8 * (int64_t)node;
            const F8 nx = ldg_f8(N), ny = ldg_f8(N + 2), nz = ldg_f8(N + 4);
            const float4 lox = nx.a, hix = nx.b, loy = ny.a, hiy = ny.b, loz = nz.a, hiz = nz.b;
            const int4 ch = __ldg((const int4 *)(N + 6));
            if (COUNT && valid) cnt->box += 4;
            uint32_t key[4];
#define TAKE_PACKET_CHILD(K, LX, HX, LY, HY, LZ, HZ, C)                                          \
            {                                                                                      \
                float a = fmaf(LX, idx, olx), b = fmaf(HX, idx, ohx);                              \
                float tn = fminf(a, b), tf = fmaxf(a, b);                                          \
                a = fmaf(LY, idy, oly); b = fmaf(HY, idy, ohy);                                    \
                tn = fmaxf(tn, fminf(a, b)); tf = fminf(tf, fmaxf(a, b));                          \
                a = fmaf(LZ, idz, olz); b = fmaf(HZ, idz, ohz);                                    \
                tn = fmaxf(tn, fminf(a, b)); tf = fminf(tf, fmaxf(a, b));                          \
                tn = fmaxf(tn, tmin_f); tf = fminf(tf, tbest_f);                                   \
                const bool h = (tn <= tf * TAKE_SLACK) && (C != TAKE_WIDE_EMPTY);                  \
                /* tn >= tmin > 0: the bit pattern orders like the value; nearest entry over the lanes that enter */ \
                const uint32_t m = __reduce_min_sync(0xffffffffu, h ? (__float_as_uint(tn) & 0xfffffffcu) : 0xffffffffu); \
                key[K] = m == 0xffffffffu ? m : (m | (uint32_t)K);                                 \
            }
            TAKE_PACKET_CHILD(0, lox.x, hix.x, loy.x, hiy.x, loz.x, hiz.x, ch.x)
            TAKE_PACKET_CHILD(1, lox.y, hix.y, loy.y, hiy.y, loz.y, hiz.y, ch.y)
            TAKE_PACKET_CHILD(2, lox.z, hix.z, loy.z, hiy.z, loz.z, hiz.z, ch.z)
            TAKE_PACKET_CHILD(3, lox.w, hix.w, loy.w, hiy.w, loz.w, hiz.w, ch.w)
#undef TAKE_PACKET_CHILD
            cswap(key[0], key[1]); cswap(key[2], key[3]); cswap(key[0], key[2]); cswap(key[1], key[3]); cswap(key[1], key[2]);
#define TAKE_WIDE_PICK(KEY) (((KEY) & 2u) ? (((KEY) & 1u) ? ch.w : ch.z) : (((KEY) & 1u) ? ch.y : ch.x))
            // (all lanes hold the same keys and store the same words to the same addresses)
            if (key[3] != 0xffffffffu) wstack[sp++] = make_uint2((uint32_t)TAKE_WIDE_PICK(key[3]), key[3] & 0xfffffffcu);
            if (key[2] != 0xffffffffu) wstack[sp++] = make_uint2((uint32_t)TAKE_WIDE_PICK(key[2]), key[2] & 0xfffffffcu);
            if (key[1] != 0xffffffffu) wstack[sp++] = make_uint2((uint32_t)TAKE_WIDE_PICK(key[1]), key[1] & 0xfffffffcu);
            __syncwarp();
            if (key[0] != 0xffffffffu) {
                node = TAKE_WIDE_PICK(key[0]);
            } else {
                for (;;) {
                    if (sp == 0) return;
                    const uint2 e = wstack[--sp];
                    node = (int32_t)e.x;
                    if (__any_sync(0xffffffffu, __uint_as_float(e.y) <= tbest_f * TAKE_SLACK)) break;
                }
            }
#undef TAKE_WIDE_PICK
        }
        {
            const int32_t code = ~node;
            const int64_t first = code >> 3;
            const int count = (code & 7) + 1;
            for (int k = 0; k < count; ++k) {
                const double2 *T = sc.tris + 6 * (first + k);
                const D4 t01 = ldg_d4(T), t23 = ldg_d4(T + 2), t45 = ldg_d4(T + 4);
                const double2 a0 = t01.a, a1 = t01.b, a2 = t23.a, a3 = t23.b, a4 = t45.a, a5 = t45.b;
                if (COUNT && valid) cnt->tri += 1;
                double t, bu = 0, bv = 0;
                bool ok;
                if (a5.y == 0.0)
                    ok = hit_triangle(mk3(a0.x, a0.y, a1.x), mk3(a2.x, a2.y, a3.x), mk3(a4.x, a4.y, a5.x), o, d, tmin, best_t, t, bu, bv);
                else
                    ok = hit_sphere(mk3(a0.x, a0.y, a1.x), a3.y, o, d, tmin, best_t, t);
                if (ok) {
                    const long long bits = __double_as_longlong(a1.y);
                    const int32_t prim = (int32_t)(bits & 0xffffffffLL), rank = (int32_t)(bits >> 32);
                    if (TIES && out.prim >= 0 && !(t < best_t)) atomicAdd(sc.tie_count, 1ULL);   // the rank decides: counted (DevScene)
                    if (t < best_t || out.prim < 0 || rank > out.rank) {
                        out.prim = prim; out.rank = rank; out.t = t; out.u = bu; out.v = bv;
                        best_t = t;
                        tbest_f = __double2float_ru(t);
                    }
                }
            }
        }
        for (;;) {
            if (sp == 0) return;
            const uint2 e = wstack[--sp];
            node = (int32_t)e.x;
            if (__any_sync(0xffffffffu, __uint_as_float(e.y) <= tbest_f * TAKE_SLACK)) break;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Traversal with lane refill for the bounce and shadow passes (same results as trace_fast4, different schedule).
//
// One ray per thread to completion leaves a warp waiting for its longest ray: in the ncu capture of the first bounce
// pass of config 2 the node loop runs with 8.6 and the leaf tests with 12.2 of 32 lanes (tools/ncu_lines.py) -- two thirds
// of the lanes have FINISHED and idle.  Here a warp stays resident and works in rounds:
//   * node phase: every lane with a ray descends until it holds a leaf or runs out of nodes;
//   * leaf phase: the lanes holding a leaf run the FP64 test together and pop their next node;
//   * a lane whose stack ran dry keeps its result in registers (state 2) -- nothing is handed over yet;
//   * when at least TAKE_REFILL_MIN lanes are without a ray, ALL finished results are handed to `io.retire` in one
//     convergent call and all free lanes take new rays from the queue with one atomicAdd.
// The hand-over (hit record, material histogram, two atomic round trips) therefore always runs for many lanes at once.
// What made it pay (each step measured on five scenes, tools/tune.py; a first version of round 1 -- node phase until every
// lane holds a leaf, then leaf phase, rays retired one by one, FP64 ray re-read with six 8-byte loads -- was 8 % SLOWER):
//   * the warp does one step at a time, a node visit or a leaf test, whichever more of its rays are waiting for: in
//     "node phase until all lanes hold a leaf" the lanes that found their leaf after one or two visits wait for the one
//     that descends eight levels (11 of 27 live lanes active in the node code, ncu);
//   * the kernels are L1-bound once the lanes are busy (L1TEX 84 % of peak): every divergent load instruction costs one
//     L1 wavefront per lane, so the FP64 origin / direction live in registers at 80 registers per thread (6 resident
//     blocks) instead of being re-read per leaf;
//   * the hits are put into queue order by the sort (k_scatter_ordered), not into the order in which the rays finished.
// Measured and not kept: one primitive of a leaf per step instead of the loop over its 1 to 4 primitives (+2 ... +10 % slower
// -- the loop keeps the loads of consecutive leaf records in flight together); a 64-byte copy of the nodes with 8-bit
// quantised child boxes, i.e. two L1 wavefronts per visit instead of four (identical images, extend +3 ... +7 % slower: the
// decode -- 24 byte-to-float conversions, a per-node grid to set up, near / far selection -- costs more issue slots than the
// L1 gives back; any-hit passes -2 ... -4 %).
// The closest hit and the tie rule do not depend on the order in which a ray's leaves are visited, and each ray is still
// traced by one lane with trace_fast4's arithmetic: identical results.
// ---------------------------------------------------------------------------------------------------------
#ifndef TAKE_REFILL_MIN
#define TAKE_REFILL_MIN 20   // measured 4 / 8 / 12 / 16 / 20 / 24 / 28 on five scenes: 20 is best or within 0.5 % of the best on all
#endif
#ifndef TAKE_REFILL_OD_REGS
#define TAKE_REFILL_OD_REGS 0
#endif
#ifndef TAKE_REFILL_NODE_BIAS
#define TAKE_REFILL_NODE_BIAS 4   // a node visit is taken when 4 * (rays at a leaf) <= BIAS * (rays at an inner node)
#endif

struct LaneRay {        // FP32 image of the ray for the box tests + the FP64 window for the leaf test
    float idx, idy, idz, olx, ohx, oly, ohy, olz, ohz, tbest_f;
    double best_t;
#if TAKE_REFILL_OD_REGS
    D3 o, d;
#else
    // The FP64 origin and direction stay in memory and are re-read per leaf test (registers): two 32-byte blocks, fetched
    // with one 256-bit load each.  o = the first three doubles of *o4; d = the last double of *o4 and the first two of *d4
    // (an extend ray: RayRec is ox oy oz dx | dy dz ...) or the first three of *d4 (a shadow ray: ShadowRec is dx dy dz tmax).
    const double2 *o4, *d4;
#endif
};

template <bool D_SPLIT>
__device__ __forceinline__ void lane_ray_od(const double2 *o4, const double2 *d4, D3 &o, D3 &d) {
    const D4 a = ldg_d4(o4), b = ldg_d4(d4);
    o = mk3(a.a.x, a.a.y, a.b.x);
    d = D_SPLIT ? mk3(a.b.y, b.a.x, b.a.y) : mk3(b.a.x, b.a.y, b.b.x);
}

template <bool D_SPLIT>
__device__ __forceinline__ void lane_ray_setup(LaneRay &r, const DevScene &sc, const double2 *o4, const double2 *d4, double tmax) {
    D3 o, d;
    lane_ray_od<D_SPLIT>(o4, d4, o, d);
    const float ox = (float)o.x, oy = (float)o.y, oz = (float)o.z;
    r.idx = safe_rcp((float)d.x); r.idy = safe_rcp((float)d.y); r.idz = safe_rcp((float)d.z);
    const float delta = 1.9073486e-6f * fmaxf(fmaxf(fabsf(ox), fabsf(oy)), fmaxf(fabsf(oz), sc.abs_max));  // 2^-19
    r.olx = -(ox + delta) * r.idx; r.ohx = -(ox - delta) * r.idx;
    r.oly = -(oy + delta) * r.idy; r.ohy = -(oy - delta) * r.idy;
    r.olz = -(oz + delta) * r.idz; r.ohz = -(oz - delta) * r.idz;
    r.tbest_f = __double2float_ru(tmax);
    r.best_t = tmax;
#if TAKE_REFILL_OD_REGS
    r.o = o; r.d = d;
#else
    r.o4 = o4; r.d4 = d4;
#endif
}

// The second half of a node visit: order the (up to four) children the ray enters, push the farther ones, return the nearest
// -- or the next live stack entry, or TAKE_NODE_DONE.  `key[k]` = (entry distance bits | k), 0xffffffff for a child not entered.
template <bool ANY_HIT>
__device__ __forceinline__ int32_t visit_finish(uint32_t (&key)[4], const int4 ch, const LaneRay &r, TravStack &st) {
    if (!(ANY_HIT && TAKE_ANYHIT_UNSORTED)) {
        cswap(key[0], key[1]); cswap(key[2], key[3]); cswap(key[0], key[2]); cswap(key[1], key[3]); cswap(key[1], key[2]);
    }
#define TAKE_WIDE_PICK(KEY) (((KEY) & 2u) ? (((KEY) & 1u) ? ch.w : ch.z) : (((KEY) & 1u) ? ch.y : ch.x))
    if (key[3] != 0xffffffffu) st.push_bits(TAKE_WIDE_PICK(key[3]), key[3] & 0xfffffffcu);
    if (key[2] != 0xffffffffu) st.push_bits(TAKE_WIDE_PICK(key[2]), key[2] & 0xfffffffcu);
    if (key[1] != 0xffffffffu) st.push_bits(TAKE_WIDE_PICK(key[1]), key[1] & 0xfffffffcu);
    if (key[0] != 0xffffffffu) return TAKE_WIDE_PICK(key[0]);
#undef TAKE_WIDE_PICK
    while (st.sp > 0) {
        int32_t nn; float tn;
        st.pop(nn, tn);
        if ((ANY_HIT && TAKE_ANYHIT_UNSORTED) || tn <= r.tbest_f * TAKE_SLACK) return nn;
    }
    return TAKE_NODE_DONE;
}

// One node visit (the body of trace_fast4's inner loop; every ray of the wavefront passes starts at tmin = TAKE_EPS):
// tests the four children, pushes the farther hits and returns the nearest -- or the next live stack entry, or TAKE_NODE_DONE.
template <bool ANY_HIT, bool COUNT>
__device__ __forceinline__ int32_t visit_wide(const DevScene &sc, int32_t node, const LaneRay &r, float tmin_f, TravStack &st, TravCounters *cnt) {
    const float4 *N = sc.wide_nodes + 8 * (int64_t)node;
    const F8 nx = ldg_f8(N), ny = ldg_f8(N + 2), nz = ldg_f8(N + 4);
    const float4 lox = nx.a, hix = nx.b, loy = ny.a, hiy = ny.b, loz = nz.a, hiz = nz.b;
    const int4 ch = __ldg((const int4 *)(N + 6));
    if (COUNT) cnt->box += 4;
    uint32_t key[4];
#define TAKE_WIDE_CHILD(K, LX, HX, LY, HY, LZ, HZ, C)                                            \
    {                                                                                              \
        float a = fmaf(LX, r.idx, r.olx), b = fmaf(HX, r.idx, r.ohx);                              \
        float tn = fminf(a, b), tf = fmaxf(a, b);                                                  \
        a = fmaf(LY, r.idy, r.oly); b = fmaf(HY, r.idy, r.ohy);                                    \
        tn = fmaxf(tn, fminf(a, b)); tf = fminf(tf, fmaxf(a, b));                                  \
        a = fmaf(LZ, r.idz, r.olz); b = fmaf(HZ, r.idz, r.ohz);                                    \
        tn = fmaxf(tn, fminf(a, b)); tf = fminf(tf, fmaxf(a, b));                                  \
        tn = fmaxf(tn, tmin_f); tf = fminf(tf, r.tbest_f);                                         \
        const bool h = (tn <= tf * TAKE_SLACK) && (C != TAKE_WIDE_EMPTY);                          \
        key[K] = h ? ((__float_as_uint(tn) & 0xfffffffcu) | (uint32_t)K) : 0xffffffffu;            \
    }
    TAKE_WIDE_CHILD(0, lox.x, hix.x, loy.x, hiy.x, loz.x, hiz.x, ch.x)
    TAKE_WIDE_CHILD(1, lox.y, hix.y, loy.y, hiy.y, loz.y, hiz.y, ch.y)
    TAKE_WIDE_CHILD(2, lox.z, hix.z, loy.z, hiy.z, loz.z, hiz.z, ch.z)
    TAKE_WIDE_CHILD(3, lox.w, hix.w, loy.w, hiy.w, loz.w, hiz.w, ch.w)
#undef TAKE_WIDE_CHILD
    return visit_finish<ANY_HIT>(key, ch, r, st);
}

// `IO` supplies  void load(uint32_t i, const DevScene&, LaneRay&)  (queue entry i -> this lane's ray) and
// void retire(unsigned mask, bool mine, const HitOut&)  (called by the whole warp; `mask` = the lanes handing over a result).
template <bool ANY_HIT, bool COUNT, bool TIES, typename IO>
__device__ __forceinline__ void trace_refill4(const DevScene &sc, IO &io, uint32_t n, uint32_t *fetch, TravStack &st, TravCounters *cnt) {
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const float tmin_f = __double2float_rd(TAKE_EPS);
    int state = 0;                 // 0: no ray, 1: tracing, 2: finished, result still in `out`
    bool exhausted = n == 0;       // warp-uniform: the queue has nothing left for this warp
    int32_t node = TAKE_NODE_DONE;
    LaneRay r;
    HitOut out;
    out.prim = -1; out.rank = -1; out.t = 0; out.u = 0; out.v = 0;
    st.sp = 0;
    for (;;) {
        const unsigned busy = __ballot_sync(0xffffffffu, state == 1);
        const int n_free = 32 - __popc(busy);
        if (n_free >= (exhausted ? 32 : TAKE_REFILL_MIN)) {
            const unsigned dmask = __ballot_sync(0xffffffffu, state == 2);
            if (dmask) io.retire(dmask, state == 2, out);
            if (state == 2) state = 0;
            if (exhausted) return;   // (only reached with all 32 lanes free)
            const int leader = __ffs(~busy) - 1;
            uint32_t base = 0;
            if (lane == leader) base = atomicAdd(fetch, (uint32_t)n_free);
            base = __shfl_sync(0xffffffffu, base, leader);
            if (base + (uint32_t)n_free >= n) exhausted = true;
            if (state == 0) {
                const uint32_t i = base + (uint32_t)__popc(~busy & lt_mask);
                if (i < n) {
                    io.load(i, sc, r);
                    out.prim = -1; out.rank = -1; out.t = 0; out.u = 0; out.v = 0;
                    st.sp = 0;
                    node = 0;
                    state = sc.num_prims > 0 ? 1 : 2;
                }
            }
            continue;
        }
        // ---- one step: the warp does what most of its rays wait for -- a node visit or a leaf test ----
        const bool want_node = state == 1 && node >= 0;
        const bool want_leaf = state == 1 && node < 0 && node != TAKE_NODE_DONE;
        const int n_node = __popc(__ballot_sync(0xffffffffu, want_node)), n_leaf = __popc(__ballot_sync(0xffffffffu, want_leaf));
        if (n_node * TAKE_REFILL_NODE_BIAS >= n_leaf * 4) {
            if (want_node) node = visit_wide<ANY_HIT, COUNT>(sc, node, r, tmin_f, st, cnt);
        } else if (want_leaf) {
            const int32_t code = ~node;
            const int64_t first = code >> 3;
            const int count = (code & 7) + 1;
#if TAKE_REFILL_OD_REGS
            const D3 o = r.o, d = r.d;
#else
            D3 o, d;
            lane_ray_od<IO::D_SPLIT>(r.o4, r.d4, o, d);
#endif
            bool stop = false;
            for (int k = 0; k < count && !stop; ++k) {
                const double2 *T = sc.tris + 6 * (first + k);
                const D4 t01 = ldg_d4(T), t23 = ldg_d4(T + 2), t45 = ldg_d4(T + 4);
                const double2 a0 = t01.a, a1 = t01.b, a2 = t23.a, a3 = t23.b, a4 = t45.a, a5 = t45.b;
                if (COUNT) cnt->tri += 1;
                double t, bu = 0, bv = 0;
                bool ok;
                if (a5.y == 0.0)
                    ok = hit_triangle(mk3(a0.x, a0.y, a1.x), mk3(a2.x, a2.y, a3.x), mk3(a4.x, a4.y, a5.x), o, d, TAKE_EPS, r.best_t, t, bu, bv);
                else
                    ok = hit_sphere(mk3(a0.x, a0.y, a1.x), a3.y, o, d, TAKE_EPS, r.best_t, t);
                if (ok) {
                    const long long bits = __double_as_longlong(a1.y);
                    const int32_t prim = (int32_t)(bits & 0xffffffffLL), rank = (int32_t)(bits >> 32);
                    if (TIES && !ANY_HIT && out.prim >= 0 && !(t < r.best_t)) atomicAdd(sc.tie_count, 1ULL);   // see trace_fast4
                    if (t < r.best_t || out.prim < 0 || rank > out.rank) {
                        out.prim = prim; out.rank = rank; out.t = t; out.u = bu; out.v = bv;
                        r.best_t = t;
                        r.tbest_f = __double2float_ru(t);
                        if (ANY_HIT) stop = true;
                    }
                }
            }
            node = TAKE_NODE_DONE;
            if (!(ANY_HIT && stop)) {
                while (st.sp > 0) {
                    int32_t nn; float tn;
                    st.pop(nn, tn);
                    if ((ANY_HIT && TAKE_ANYHIT_UNSORTED) || tn <= r.tbest_f * TAKE_SLACK) { node = nn; break; }
                }
            }
        }
        if (state == 1 && node == TAKE_NODE_DONE) state = 2;
    }
}

#if TAKE_EXPERIMENTAL
// ---------------------------------------------------------------------------------------------------------
// Speculative 4-wide traversal (warp-cooperative schedule, same results as trace_fast4).
//
// trace_fast4 alternates "descend until MY ray holds a leaf" and "test MY leaf": a lane that reaches its leaf early
// idles until the slowest lane of the warp gets there, and the long FP64 leaf test then runs for whichever lanes
// happen to hold one.  Here a lane that reaches a leaf *postpones* it and keeps descending from its stack while any
// other lane of the warp is still searching (__any_sync); the leaf phase starts only when every lane holds a leaf or
// has run dry, so the FP64 tests execute with as many lanes as possible.  Speculated visits use the not-yet-shrunk
// best distance, so they may test a few more boxes than trace_fast4 -- never fewer -- and the closest hit and the
// tie rule are order-independent, so the result is identical.
// Must be entered by all lanes of `__activemask()` together (the kernels call it under `if (valid)`).
// ---------------------------------------------------------------------------------------------------------
template <bool ANY_HIT, bool COUNT>
__device__ __forceinline__ void trace_spec4(const DevScene &sc, D3 o, D3 d, double tmin, double tmax, TravStack &st,
                                            HitOut &out, TravCounters *cnt) {
    const unsigned wmask = __activemask();
    out.prim = -1; out.rank = -1; out.t = 0; out.u = 0; out.v = 0;
    const float ox = (float)o.x, oy = (float)o.y, oz = (float)o.z;
    const float idx = safe_rcp((float)d.x), idy = safe_rcp((float)d.y), idz = safe_rcp((float)d.z);
    const float delta = 1.9073486e-6f * fmaxf(fmaxf(fabsf(ox), fabsf(oy)), fmaxf(fabsf(oz), sc.abs_max));  // 2^-19
    const float olx = -(ox + delta) * idx, ohx = -(ox - delta) * idx;
    const float oly = -(oy + delta) * idy, ohy = -(oy - delta) * idy;
    const float olz = -(oz + delta) * idz, ohz = -(oz - delta) * idz;
    const float tmin_f = __double2float_rd(tmin);
    float tbest_f = __double2float_ru(tmax);
    double best_t = tmax;

    st.sp = 0;
    int32_t node = sc.num_prims > 0 ? 0 : TAKE_NODE_DONE;
    int32_t leaf = 0;  // postponed leaf link (< 0), or 0 = none
#define TAKE_SPEC_POP()                                                      \
    {                                                                         \
        node = TAKE_NODE_DONE;                                                \
        while (st.sp > 0) {                                                   \
            int32_t nn; float tn;                                             \
            st.pop(nn, tn);                                                   \
            if (tn <= tbest_f * TAKE_SLACK) { node = nn; break; }             \
        }                                                                     \
    }
    while (__any_sync(wmask, node != TAKE_NODE_DONE)) {
        // ---- node phase: runs while some lane has neither a leaf nor an empty stack ----
        for (;;) {
            if (node >= 0) {
                const float4 *N = sc.wide_nodes + 8 * (int64_t)node;
                const F8 nx = ldg_f8(N), ny = ldg_f8(N + 2), nz = ldg_f8(N + 4);
                const float4 lox = nx.a, hix = nx.b, loy = ny.a, hiy = ny.b, loz = nz.a, hiz = nz.b;
                const int4 ch = __ldg((const int4 *)(N + 6));
                if (COUNT) cnt->box += 4;
                uint32_t key[4];
#define TAKE_WIDE_CHILD(K, LX, HX, LY, HY, LZ, HZ, C)                                            \
                {                                                                                  \
                    float a = fmaf(LX, idx, olx), b = fmaf(HX, idx, ohx);                          \
                    float tn = fminf(a, b), tf = fmaxf(a, b);                                      \
                    a = fmaf(LY, idy, oly); b = fmaf(HY, idy, ohy);                                \
                    tn = fmaxf(tn, fminf(a, b)); tf = fminf(tf, fmaxf(a, b));                      \
                    a = fmaf(LZ, idz, olz); b = fmaf(HZ, idz, ohz);                                \
                    tn = fmaxf(tn, fminf(a, b)); tf = fminf(tf, fmaxf(a, b));                      \
                    tn = fmaxf(tn, tmin_f); tf = fminf(tf, tbest_f);                               \
                    const bool h = (tn <= tf * TAKE_SLACK) && (C != TAKE_WIDE_EMPTY);              \
                    key[K] = h ? ((__float_as_uint(tn) & 0xfffffffcu) | (uint32_t)K) : 0xffffffffu; \
                }
                TAKE_WIDE_CHILD(0, lox.x, hix.x, loy.x, hiy.x, loz.x, hiz.x, ch.x)
                TAKE_WIDE_CHILD(1, lox.y, hix.y, loy.y, hiy.y, loz.y, hiz.y, ch.y)
                TAKE_WIDE_CHILD(2, lox.z, hix.z, loy.z, hiy.z, loz.z, hiz.z, ch.z)
                TAKE_WIDE_CHILD(3, lox.w, hix.w, loy.w, hiy.w, loz.w, hiz.w, ch.w)
#undef TAKE_WIDE_CHILD
                cswap(key[0], key[1]); cswap(key[2], key[3]); cswap(key[0], key[2]); cswap(key[1], key[3]); cswap(key[1], key[2]);
#define TAKE_WIDE_PICK(KEY) (((KEY) & 2u) ? (((KEY) & 1u) ? ch.w : ch.z) : (((KEY) & 1u) ? ch.y : ch.x))
                if (key[3] != 0xffffffffu) st.push_bits(TAKE_WIDE_PICK(key[3]), key[3] & 0xfffffffcu);
                if (key[2] != 0xffffffffu) st.push_bits(TAKE_WIDE_PICK(key[2]), key[2] & 0xfffffffcu);
                if (key[1] != 0xffffffffu) st.push_bits(TAKE_WIDE_PICK(key[1]), key[1] & 0xfffffffcu);
                if (key[0] != 0xffffffffu) node = TAKE_WIDE_PICK(key[0]);
                else TAKE_SPEC_POP()
#undef TAKE_WIDE_PICK
            }
            if (node < 0 && node != TAKE_NODE_DONE && leaf == 0) {  // first leaf: postpone it, keep going
                leaf = node;
                TAKE_SPEC_POP()
            }
            // a lane is still searching if it holds no leaf yet and has nodes left
            if (!__any_sync(wmask, leaf == 0 && node != TAKE_NODE_DONE)) break;
        }
        // ---- leaf phase: the postponed leaf, then any further leaves that are next on this lane's stack ----
        while (leaf < 0) {
            const int32_t code = ~leaf;
            const int64_t first = code >> 3;
            const int count = (code & 7) + 1;
            leaf = 0;
            for (int k = 0; k < count; ++k) {
                const double2 *T = sc.tris + 6 * (first + k);
                const D4 t01 = ldg_d4(T), t23 = ldg_d4(T + 2), t45 = ldg_d4(T + 4);
                const double2 a0 = t01.a, a1 = t01.b, a2 = t23.a, a3 = t23.b, a4 = t45.a, a5 = t45.b;
                if (COUNT) cnt->tri += 1;
                double t, bu = 0, bv = 0;
                bool ok;
                if (a5.y == 0.0)
                    ok = hit_triangle(mk3(a0.x, a0.y, a1.x), mk3(a2.x, a2.y, a3.x), mk3(a4.x, a4.y, a5.x), o, d, tmin, best_t, t,
                                      bu, bv);
                else
                    ok = hit_sphere(mk3(a0.x, a0.y, a1.x), a3.y, o, d, tmin, best_t, t);
                if (ok) {
                    const long long bits = __double_as_longlong(a1.y);
                    const int32_t prim = (int32_t)(bits & 0xffffffffLL), rank = (int32_t)(bits >> 32);
                    if (t < best_t || out.prim < 0 || rank > out.rank) {
                        out.prim = prim; out.rank = rank; out.t = t; out.u = bu; out.v = bv;
                        best_t = t;
                        tbest_f = __double2float_ru(t);
                        if (ANY_HIT) { node = TAKE_NODE_DONE; st.sp = 0; k = count; }
                    }
                }
            }
            if (node < 0 && node != TAKE_NODE_DONE) {
                leaf = node;
                TAKE_SPEC_POP()
            }
        }
        // the hit may have shrunk the window below the entry distance of the inner node picked speculatively
        // (harmless: its boxes are re-tested against the new window when it is visited)
    }
#undef TAKE_SPEC_POP
}

#endif  // TAKE_EXPERIMENTAL (speculative traversal)

#ifndef TAKE_SPECULATE
#define TAKE_SPECULATE 0
#endif

// Dispatch on the tree width chosen at scene creation (the shipped library only has the 4-wide tree).
template <bool ANY_HIT, bool COUNT, bool WIDE, bool TIES = false>
__device__ __forceinline__ void trace_any(const DevScene &sc, D3 o, D3 d, double tmin, double tmax, TravStack &st, HitOut &out,
                                          TravCounters *cnt) {
#if TAKE_EXPERIMENTAL
    if (WIDE) {
        if (TAKE_SPECULATE) trace_spec4<ANY_HIT, COUNT>(sc, o, d, tmin, tmax, st, out, cnt);
        else trace_fast4<ANY_HIT, COUNT, TIES>(sc, o, d, tmin, tmax, st, out, cnt);
    }
    else trace_fast<ANY_HIT, COUNT>(sc, o, d, tmin, tmax, st, out, cnt);
#else
    static_assert(WIDE, "binary-node traversal needs -DTAKE_EXPERIMENTAL=1");
    trace_fast4<ANY_HIT, COUNT, TIES>(sc, o, d, tmin, tmax, st, out, cnt);
#endif
}


}  // namespace take
