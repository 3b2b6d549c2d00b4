// Device builder for the FAST tree: replaces the host's binned-SAH build of bvh_build.cpp on the
// critical path of take_gpu_scene_create.  (construct_bvh, src/bvh.cpp:8-45, defines the reference's tree; that tree -- needed
// only for the tie-break ranks and the exact mode -- keeps being built on the host, in the background, see take_gpu.cu.)
//
// Any tree whose boxes conservatively contain their primitives gives the same closest hit (traverse.cuh), so the builder
// is free to choose topology for speed.  It is PLOC (parallel locally-ordered clustering, Meister & Bittner 2018), a
// bottom-up agglomerative build whose trees are close to full-sweep SAH quality:
//   1. k_prim_boxes      conservative FP32 box per primitive (FP64 bounds rounded outward + one ulp, like the host's
//                        round_down / round_up), centroid bounds and max |coordinate| by warp-reduced atomics
//   2. k_morton          63-bit Morton code of each centroid; cub radix sort of (code, primitive)
//   3. PLOC rounds       every cluster looks at its R neighbours on either side in Morton order and picks the one whose
//                        union has the smallest surface area (total order on (area, lower index, higher index), so the
//                        globally best pair is always mutual and every round merges at least one pair); mutual pairs
//                        merge into a new node; the survivors are compacted in order.  Node ids and positions come from
//                        prefix sums, not atomics: the tree's memory layout is the same on every run.
//                        While merging, each node gets its SAH cost with the host builder's constants (node visit 1,
//                        leaf test 1.2) and subtrees of up to `max_leaf` primitives that are cheaper as a leaf are marked.
//   4. wide collapse     breadth-first from the root: a binary node's two children are opened (largest surface area first)
//                        until there are four -- the same rule as the host's WidePolicy -- leaf slots are numbered in
//                        depth-first order on the way, 128-byte WideNodes are written level by level.
//   5. k_leaf_records    the 96-byte FP64 leaf records in leaf order (rank field filled in later by k_patch_ranks).
// All of it is a few milliseconds for a million primitives; the kernels are simple streaming / gather passes.
#pragma once
#include <cub/cub.cuh>

#include "device_common.cuh"

namespace take {
namespace devbuild {

#ifndef TAKE_PLOC_RADIUS
#define TAKE_PLOC_RADIUS 16
#endif
#define TAKE_DB_BLOCK 256

struct BNode {          // binary node of the PLOC tree (leaves: nodes [0, n) in Morton order; inner nodes follow)
    float4 lo, hi;      // conservative FP32 box; lo.w = surface half-area, hi.w = SAH cost of the subtree
    int32_t left, right;  // children (inner) or -1, primitive id (single-primitive leaf)
    int32_t count;      // primitives below
    int32_t leaf;       // 1: the whole subtree is one leaf of `count` primitives
};
static_assert(sizeof(BNode) == 48, "BNode");

__device__ __forceinline__ float next_down(float f) {   // nextafterf(f, -inf) for finite f
    if (f == 0.0f) return -1.401298464324817e-45f;
    uint32_t b = __float_as_uint(f);
    b = f > 0.0f ? b - 1u : b + 1u;
    return __uint_as_float(b);
}
__device__ __forceinline__ float next_up(float f) { return -next_down(-f); }
// the host's round_down(v, 0) / round_up(v, 0) (bvh_build.cpp): outward rounding, then one more ulp
__device__ __forceinline__ float box_lo(double v) { return next_down(__double2float_rd(v)); }
__device__ __forceinline__ float box_hi(double v) { return next_up(__double2float_ru(v)); }
__device__ __forceinline__ float half_area(float4 lo, float4 hi) {
    const float dx = hi.x - lo.x, dy = hi.y - lo.y, dz = hi.z - lo.z;
    return dx * dy + dy * dz + dz * dx;
}
// monotone float <-> uint maps for atomicMin / atomicMax on floats
__device__ __forceinline__ uint32_t f2ord(float f) { const uint32_t b = __float_as_uint(f); return (b & 0x80000000u) ? ~b : (b | 0x80000000u); }
__device__ __forceinline__ float ord2f(uint32_t u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }

struct Globals {        // zero / identity initialised by the host before k_prim_boxes
    uint32_t cmin[3], cmax[3];          // centroid bounds (ordered-uint encoding)
    unsigned long long abs_max_bits;    // max |coordinate| over the FP64 primitive bounds (bits of a non-negative double)
    uint32_t pad;
};

// ---- 1. primitive boxes -------------------------------------------------------------------------------------------
__global__ void k_prim_boxes(DevScene sc, int64_t n, float4 *plo, float4 *phi, Globals *g) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float c[3] = {INFINITY, INFINITY, INFINITY};
    double amax = 0.0;
    const bool valid = i < n;
    if (valid) {
        double lo[3], hi[3];
        const int32_t *id = sc.indices + 3 * i;
        if (sc.prim_flags[i] & TAKE_PRIM_SPHERE) {    // src/scene.cpp:4-23
            const double *s = sc.spheres + 4 * (int64_t)id[0];
            for (int a = 0; a < 3; ++a) { lo[a] = s[a] - s[3]; hi[a] = s[a] + s[3]; }
        } else {
            const double *p0 = sc.positions + 3 * (int64_t)id[0], *p1 = sc.positions + 3 * (int64_t)id[1], *p2 = sc.positions + 3 * (int64_t)id[2];
            for (int a = 0; a < 3; ++a) { lo[a] = fmin(fmin(p0[a], p1[a]), p2[a]); hi[a] = fmax(fmax(p0[a], p1[a]), p2[a]); }
        }
        const float4 l = make_float4(box_lo(lo[0]), box_lo(lo[1]), box_lo(lo[2]), 0.0f);
        const float4 h = make_float4(box_hi(hi[0]), box_hi(hi[1]), box_hi(hi[2]), 0.0f);
        plo[i] = l; phi[i] = h;
        c[0] = 0.5f * (l.x + h.x); c[1] = 0.5f * (l.y + h.y); c[2] = 0.5f * (l.z + h.z);
        for (int a = 0; a < 3; ++a) amax = fmax(amax, fmax(fabs(lo[a]), fabs(hi[a])));
    }
    // warp reduce, one atomic per warp and quantity
    for (int a = 0; a < 3; ++a) {
        float mn = valid ? c[a] : INFINITY, mx = valid ? c[a] : -INFINITY;
        for (int o = 16; o > 0; o >>= 1) { mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o)); mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o)); }
        if ((threadIdx.x & 31) == 0 && mn <= mx) { atomicMin(&g->cmin[a], f2ord(mn)); atomicMax(&g->cmax[a], f2ord(mx)); }
    }
    for (int o = 16; o > 0; o >>= 1) amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    if ((threadIdx.x & 31) == 0) atomicMax(&g->abs_max_bits, (unsigned long long)__double_as_longlong(amax));
}

// ---- 2. Morton codes ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t spread21(uint32_t v) {   // 21 bits -> every third bit of 63
    uint64_t x = v & 0x1fffffu;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}
__global__ void k_morton(int64_t n, const float4 *plo, const float4 *phi, const Globals *g, uint64_t *keys, uint32_t *vals) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 l = plo[i], h = phi[i];
    const float c[3] = {0.5f * (l.x + h.x), 0.5f * (l.y + h.y), 0.5f * (l.z + h.z)};
    uint32_t q[3];
    for (int a = 0; a < 3; ++a) {
        const float mn = ord2f(g->cmin[a]), mx = ord2f(g->cmax[a]);
        const float ext = mx - mn;
        const float t = ext > 0.0f ? (c[a] - mn) / ext : 0.0f;
        q[a] = (uint32_t)fminf(fmaxf(t * 2097152.0f, 0.0f), 2097151.0f);
    }
    keys[i] = (spread21(q[0]) << 2) | (spread21(q[1]) << 1) | spread21(q[2]);
    vals[i] = (uint32_t)i;
}

// leaves of the PLOC tree: node i = the i-th primitive in Morton order
__global__ void k_init_leaves(int64_t n, const uint32_t *sorted_prim, const float4 *plo, const float4 *phi, BNode *nodes, int32_t *cluster,
                              float c_isect) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t p = sorted_prim[i];
    BNode b;
    b.lo = plo[p]; b.hi = phi[p];
    b.lo.w = half_area(b.lo, b.hi);
    b.hi.w = c_isect;
    b.left = -1; b.right = (int32_t)p; b.count = 1; b.leaf = 1;
    nodes[i] = b;
    cluster[i] = (int32_t)i;
}

// ---- 3. PLOC round ------------------------------------------------------------------------------------------------
// nearest neighbour of every cluster within +-R positions: smallest surface area of the union, ties by (lower, higher) index
template <int R>
__global__ void k_ploc_nn(int32_t m, const int32_t *cluster, const BNode *nodes, int32_t *nn) {
    __shared__ float4 slo[TAKE_DB_BLOCK + 2 * R], shi[TAKE_DB_BLOCK + 2 * R];
    const int32_t base = (int32_t)blockIdx.x * TAKE_DB_BLOCK - R;
    for (int t = threadIdx.x; t < TAKE_DB_BLOCK + 2 * R; t += TAKE_DB_BLOCK) {
        const int32_t j = base + t;
        if (j >= 0 && j < m) {
            const BNode &b = nodes[cluster[j]];
            slo[t] = b.lo; shi[t] = b.hi;
        }
    }
    __syncthreads();
    const int32_t i = (int32_t)blockIdx.x * TAKE_DB_BLOCK + threadIdx.x;
    if (i >= m) return;
    const float4 lo = slo[threadIdx.x + R], hi = shi[threadIdx.x + R];
    float best = INFINITY;
    int32_t best_j = -1;
    // ascending j with strict '<' picks the lowest j among equal areas for j > i; for j < i the pair key is (area, j, i), and
    // ascending j again prefers the smaller lower index: the same total order from both ends of a pair
    for (int d = -R; d <= R; ++d) {
        const int32_t j = i + d;
        if (d == 0 || j < 0 || j >= m) continue;
        const float4 l2 = slo[threadIdx.x + R + d], h2 = shi[threadIdx.x + R + d];
        const float4 ul = make_float4(fminf(lo.x, l2.x), fminf(lo.y, l2.y), fminf(lo.z, l2.z), 0.0f);
        const float4 uh = make_float4(fmaxf(hi.x, h2.x), fmaxf(hi.y, h2.y), fmaxf(hi.z, h2.z), 0.0f);
        const float a = half_area(ul, uh);
        if (a < best) { best = a; best_j = j; }
    }
    nn[i] = best_j;
}

// fallback pairing (2k, 2k+1): used for a round in which no union area was finite
__global__ void k_ploc_pairs(int32_t m, int32_t *nn) {
    const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const int32_t j = i ^ 1;
    nn[i] = j < m ? j : -1;
}

// flags of a round, packed for one prefix sum: low word = the cluster survives (is not the higher half of a merging pair),
// high word = the cluster is the lower half of a merging pair (a new node is created for it)
__global__ void k_ploc_flags(int32_t m, const int32_t *nn, uint64_t *flags) {
    const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const int32_t j = nn[i];
    const bool mutual = j >= 0 && nn[j] == i;
    const uint64_t survives = (mutual && j < i) ? 0ull : 1ull;
    const uint64_t creates = (mutual && i < j) ? 1ull : 0ull;
    flags[i] = survives | (creates << 32);
}

__global__ void k_ploc_merge(int32_t m, const int32_t *nn, const uint64_t *flags, const uint64_t *scan, const int32_t *cluster_in,
                             int32_t *cluster_out, BNode *nodes, int32_t next_node, int max_leaf, float c_trav, float c_isect) {
    const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const uint64_t f = flags[i], s = scan[i];
    if (!(f & 1ull)) return;                     // merged away into its partner
    const int32_t pos = (int32_t)(s & 0xffffffffull);
    int32_t id = cluster_in[i];
    if (f >> 32) {                               // lower half of a mutual pair: create the parent
        const int32_t a = id, b = cluster_in[nn[i]];
        const BNode na = nodes[a], nb = nodes[b];
        BNode p;
        p.lo = make_float4(fminf(na.lo.x, nb.lo.x), fminf(na.lo.y, nb.lo.y), fminf(na.lo.z, nb.lo.z), 0.0f);
        p.hi = make_float4(fmaxf(na.hi.x, nb.hi.x), fmaxf(na.hi.y, nb.hi.y), fmaxf(na.hi.z, nb.hi.z), 0.0f);
        const float area = half_area(p.lo, p.hi);
        p.lo.w = area;
        p.left = a; p.right = b;
        p.count = na.count + nb.count;
        // SAH with the host builder's constants (bvh_build.cpp: c_trav 1, c_isect 1.2)
        const float inner = c_trav + (area > 0.0f ? (na.lo.w * na.hi.w + nb.lo.w * nb.hi.w) / area : na.hi.w + nb.hi.w);
        const float as_leaf = c_isect * (float)p.count;
        p.leaf = (p.count <= max_leaf && as_leaf <= inner) ? 1 : 0;
        p.hi.w = p.leaf ? as_leaf : inner;
        id = next_node + (int32_t)(s >> 32);
        nodes[id] = p;
    }
    cluster_out[pos] = id;
}

// ---- 4. wide collapse (breadth first) -------------------------------------------------------------------------------
struct WorkItem {
    int32_t bnode;   // inner binary node this wide node stands for
    int32_t wide;    // index of the wide node to fill
    int32_t first;   // first leaf slot of the subtree
    int32_t pad;
};
struct Kids {        // result of pass A for one item
    int32_t node[4], first[4];
    int32_t nk, n_inner;
};

__device__ __forceinline__ bool is_leaf_node(const BNode &b) { return b.leaf != 0; }

// pass A: choose the (up to) four children and count the inner ones
__global__ void k_wide_kids(int32_t n_items, const WorkItem *items, const BNode *nodes, Kids *kids, uint32_t *inner_count) {
    const int32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_items) return;
    const WorkItem it = items[t];
    const BNode b = nodes[it.bnode];
    Kids k;
    k.node[0] = b.left; k.first[0] = it.first;
    k.node[1] = b.right; k.first[1] = it.first + nodes[b.left].count;
    k.nk = 2;
    for (int q = 2; q < 4; ++q) { k.node[q] = -1; k.first[q] = 0; }
    while (k.nk < 4) {   // open the inner child with the largest surface area (ties: the first), as WidePolicy::kids does
        int best = -1;
        float best_area = -1.0f;
        for (int q = 0; q < k.nk; ++q) {
            const BNode &c = nodes[k.node[q]];
            if (is_leaf_node(c)) continue;
            if (c.lo.w > best_area) { best_area = c.lo.w; best = q; }
        }
        if (best < 0) break;
        const BNode c = nodes[k.node[best]];
        const int32_t f0 = k.first[best];
        k.node[best] = c.left;
        k.node[k.nk] = c.right; k.first[k.nk] = f0 + nodes[c.left].count;
        k.nk++;
    }
    int ni = 0;
    for (int q = 0; q < k.nk; ++q) ni += is_leaf_node(nodes[k.node[q]]) ? 0 : 1;
    k.n_inner = ni;
    kids[t] = k;
    inner_count[t] = (uint32_t)ni;
}

// primitives of a collapsed leaf subtree (<= 8), written in depth-first order
__device__ inline void emit_leaf_prims(const BNode *nodes, int32_t root, int32_t first, int32_t *leaf_prims) {
    int32_t stack[16];
    int sp = 0;
    stack[sp++] = root;
    int32_t o = first;
    while (sp > 0) {
        const BNode &b = nodes[stack[--sp]];
        if (b.left < 0) { leaf_prims[o++] = b.right; continue; }
        stack[sp++] = b.right;
        stack[sp++] = b.left;
    }
}

// pass B: write the wide node, the leaf slots of its leaf children, and the work items of its inner children
__global__ void k_wide_emit(int32_t n_items, const WorkItem *items, const Kids *kids, const uint32_t *inner_scan, const BNode *nodes,
                            WideNode *wide, int32_t next_wide, WorkItem *next_items, int32_t *leaf_prims) {
    const int32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_items) return;
    const WorkItem it = items[t];
    const Kids k = kids[t];
    WideNode w;
    int32_t slot = (int32_t)inner_scan[t];
    for (int q = 0; q < 4; ++q) {
        if (q >= k.nk) {
            w.lox[q] = w.loy[q] = w.loz[q] = INFINITY;
            w.hix[q] = w.hiy[q] = w.hiz[q] = -INFINITY;
            w.child[q] = TAKE_WIDE_EMPTY; w.count[q] = 0;
            continue;
        }
        const BNode &c = nodes[k.node[q]];
        w.lox[q] = c.lo.x; w.loy[q] = c.lo.y; w.loz[q] = c.lo.z;
        w.hix[q] = c.hi.x; w.hiy[q] = c.hi.y; w.hiz[q] = c.hi.z;
        if (is_leaf_node(c)) {
            w.child[q] = ~(int32_t)(((uint32_t)k.first[q] << 3) | (uint32_t)(c.count - 1));
            w.count[q] = c.count;
            emit_leaf_prims(nodes, k.node[q], k.first[q], leaf_prims);
        } else {
            const int32_t wi = next_wide + slot;
            w.child[q] = wi; w.count[q] = 0;
            WorkItem ni;
            ni.bnode = k.node[q]; ni.wide = wi; ni.first = k.first[q]; ni.pad = 0;
            next_items[slot] = ni;
            ++slot;
        }
    }
    wide[it.wide] = w;
}

// the tree is a single leaf (or empty): a root whose only child is that leaf (bvh_build.cpp: WidePolicy::wrap_root)
__global__ void k_wide_wrap_root(const BNode *nodes, int32_t root, int64_t n, WideNode *wide, int32_t *leaf_prims) {
    if (blockIdx.x || threadIdx.x) return;
    WideNode w;
    for (int q = 0; q < 4; ++q) {
        w.lox[q] = w.loy[q] = w.loz[q] = INFINITY;
        w.hix[q] = w.hiy[q] = w.hiz[q] = -INFINITY;
        w.child[q] = TAKE_WIDE_EMPTY; w.count[q] = 0;
    }
    if (n > 0) {
        const BNode &c = nodes[root];
        w.lox[0] = c.lo.x; w.loy[0] = c.lo.y; w.loz[0] = c.lo.z;
        w.hix[0] = c.hi.x; w.hiy[0] = c.hi.y; w.hiz[0] = c.hi.z;
        w.child[0] = ~(int32_t)(uint32_t)(c.count - 1);
        w.count[0] = c.count;
        emit_leaf_prims(nodes, root, 0, leaf_prims);
    }
    wide[0] = w;
}

// ---- 5. leaf records ----------------------------------------------------------------------------------------------
// v0 | (rank << 32 | prim) | e1 | radius | e2 | kind -- the layout host_build writes (take_gpu.cu), e1 / e2 with the same FP64
// subtractions.  The rank (the reference's DFS order, for equal-t ties) is not known yet: k_patch_ranks fills it in.
__global__ void k_leaf_records(DevScene sc, int64_t n, const int32_t *leaf_prims, double *tris) {
    const int64_t slot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= n) return;
    const int32_t prim = leaf_prims[slot];
    double *T = tris + 12 * slot;
    const int32_t *id = sc.indices + 3 * (int64_t)prim;
    T[3] = __longlong_as_double((long long)(uint32_t)prim);
    if (sc.prim_flags[prim] & TAKE_PRIM_SPHERE) {
        const double *sp = sc.spheres + 4 * (int64_t)id[0];
        T[0] = sp[0]; T[1] = sp[1]; T[2] = sp[2];
        T[4] = T[5] = T[6] = 0; T[7] = sp[3];
        T[8] = T[9] = T[10] = 0; T[11] = 1.0;
    } else {
        const double *p0 = sc.positions + 3 * (int64_t)id[0], *p1 = sc.positions + 3 * (int64_t)id[1], *p2 = sc.positions + 3 * (int64_t)id[2];
        for (int a = 0; a < 3; ++a) {
            T[a] = p0[a];
            T[4 + a] = p1[a] - p0[a];   // e1 = v1 - v0, e2 = v2 - v0 (src/shape.cpp:53-54), computed once
            T[8 + a] = p2[a] - p0[a];
        }
        T[7] = 0; T[11] = 0.0;
    }
}

__global__ void k_patch_ranks(int64_t n, const int32_t *dfs_rank, double *tris) {
    const int64_t slot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= n) return;
    double *T = tris + 12 * slot;
    const uint32_t prim = (uint32_t)__double_as_longlong(T[3]);
    T[3] = __longlong_as_double(((long long)dfs_rank[prim] << 32) | (long long)prim);
}

}  // namespace devbuild
}  // namespace take
