// Device builder for the FAST tree: replaces the host's binned-SAH build of bvh_build.cpp on the
// critical path of take_gpu_scene_create.  (construct_bvh, src/bvh.cpp:8-45, defines the reference's tree; that tree -- needed
// only for the tie-break ranks and the exact mode -- keeps being built on the host, in the background, see take_gpu.cu.)
//
// Any tree whose boxes conservatively contain their primitives gives the same closest hit (traverse.cuh), so the builder
// is free to choose topology; what it must keep is traversal speed.  It is therefore the SAME algorithm as the host builder
// (32-bin SAH over centroids on all three axes, node visit 1, leaf test 1.2, leaves of up to max_leaf primitives when
// cheaper) run level by level on the device.  (A first version clustered bottom-up -- PLOC over Morton-sorted primitives --
// and built in 23 ms, but its trees cost 29 % more box tests per ray on config 2: measured, dropped.)
//   1. k_prim_boxes   conservative FP32 box per primitive (FP64 bounds rounded outward + one ulp, like the host's round_down /
//                     round_up), scene bounds, centroid bounds and max |coordinate| by warp-reduced atomics
//   2. SAH levels     one thread block per node with more than SMALL primitives: bin its range (shared-memory bins), three
//                     threads sweep the three axes, the block partitions the range stably into the other buffer with block
//                     scans while accumulating both children's box and centroid bounds -- two passes over the range, as on
//                     the host.  Nodes of up to SMALL primitives are finished by ONE thread each (the host's sparse sweep,
//                     in-place partition) in a final array.  Positions come from prefix sums and min / max / integer
//                     atomics, so the tree and the leaf order are the same on every run.
//   3. wide collapse  breadth-first from the root: a binary node's two children are opened (largest surface area first)
//                     until there are four -- the same rule as the host's WidePolicy -- 128-byte WideNodes are written
//                     level by level.  A node's leaf slots are its range of the final primitive order.
//   4. k_leaf_records the 96-byte FP64 leaf records in leaf order (rank field filled in later by k_patch_ranks).
#pragma once
#include <cub/cub.cuh>

#include "device_common.cuh"

namespace take {
namespace devbuild {

#define TAKE_DB_BLOCK 256
#define TAKE_SAH_BINS 32
#define TAKE_SAH_SMALL 16   // ranges up to this size are finished by one thread (SahBuilder::SMALL on the host)

struct BNode {          // binary node of the SAH tree
    float4 lo, hi;      // conservative FP32 box; lo.w = surface half-area
    int32_t left, right;  // children, or -1 while the node is a leaf
    int32_t count;      // primitives below
    int32_t leaf;       // 1: a leaf of `count` primitives (its slots are its range of the final primitive order)
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
__device__ __forceinline__ float half_area3(const float *lo, const float *hi) {
    const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
    if (dx < 0.0f || dy < 0.0f || dz < 0.0f) return 0.0f;   // empty box
    return dx * dy + dy * dz + dz * dx;
}
__device__ __forceinline__ float half_area(float4 lo, float4 hi) {
    const float l[3] = {lo.x, lo.y, lo.z}, h[3] = {hi.x, hi.y, hi.z};
    return half_area3(l, h);
}
// monotone float <-> uint maps for min / max on floats through integer atomics and redux
__device__ __forceinline__ uint32_t f2ord(float f) { const uint32_t b = __float_as_uint(f); return (b & 0x80000000u) ? ~b : (b | 0x80000000u); }
__device__ __forceinline__ float ord2f(uint32_t u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }
#define TAKE_ORD_PLUS_INF 0xff800000u    // f2ord(+inf)
#define TAKE_ORD_MINUS_INF 0x007fffffu   // f2ord(-inf)

struct Globals {        // identity-initialised by the host before k_prim_boxes
    uint32_t bmin[3], bmax[3];          // scene bounds (ordered-uint encoding)
    uint32_t cmin[3], cmax[3];          // centroid bounds
    unsigned long long abs_max_bits;    // max |coordinate| over the FP64 primitive bounds (bits of a non-negative double)
    int32_t node_count, huge_count, med_count, small_count, chunk_count, pad[3];   // allocation counters of the SAH levels
};

// ---- 1. primitive boxes -------------------------------------------------------------------------------------------
__global__ void k_prim_boxes(DevScene sc, int64_t n, float4 *plo, float4 *phi, Globals *g) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float c[3] = {INFINITY, INFINITY, INFINITY}, bl[3] = {INFINITY, INFINITY, INFINITY}, bh[3] = {-INFINITY, -INFINITY, -INFINITY};
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
        const float4 l = make_float4(box_lo(lo[0]), box_lo(lo[1]), box_lo(lo[2]), __int_as_float((int32_t)i));   // .w: the primitive id
        const float4 h = make_float4(box_hi(hi[0]), box_hi(hi[1]), box_hi(hi[2]), 0.0f);
        plo[i] = l; phi[i] = h;
        c[0] = 0.5f * (l.x + h.x); c[1] = 0.5f * (l.y + h.y); c[2] = 0.5f * (l.z + h.z);
        bl[0] = l.x; bl[1] = l.y; bl[2] = l.z; bh[0] = h.x; bh[1] = h.y; bh[2] = h.z;
        for (int a = 0; a < 3; ++a) amax = fmax(amax, fmax(fabs(lo[a]), fabs(hi[a])));
    }
    // warp reduce (redux on the ordered-uint images), one atomic per warp and quantity
    const bool lane0 = (threadIdx.x & 31) == 0;
    for (int a = 0; a < 3; ++a) {
        const uint32_t cmn = __reduce_min_sync(0xffffffffu, valid ? f2ord(c[a]) : 0xffffffffu);
        const uint32_t cmx = __reduce_max_sync(0xffffffffu, valid ? f2ord(c[a]) : 0u);
        const uint32_t bmn = __reduce_min_sync(0xffffffffu, valid ? f2ord(bl[a]) : 0xffffffffu);
        const uint32_t bmx = __reduce_max_sync(0xffffffffu, valid ? f2ord(bh[a]) : 0u);
        if (lane0 && cmn <= cmx) {
            atomicMin(&g->cmin[a], cmn); atomicMax(&g->cmax[a], cmx);
            atomicMin(&g->bmin[a], bmn); atomicMax(&g->bmax[a], bmx);
        }
    }
    for (int o = 16; o > 0; o >>= 1) amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    if ((threadIdx.x & 31) == 0) atomicMax(&g->abs_max_bits, (unsigned long long)__double_as_longlong(amax));
}

// ---- 2. SAH levels --------------------------------------------------------------------------------------------------
#define TAKE_SAH_CHUNK 4096                    // primitives per block of the chunked path
#define TAKE_SAH_HUGE (4 * TAKE_SAH_CHUNK)     // nodes above this size are split by several blocks

struct SItem {          // a node whose range still has to be split
    int32_t node, lo, hi, pad;
    float bl[3], bh[3], cl[3], ch[3];   // box of the primitives, box of their centroids
};
static_assert(sizeof(SItem) == 64, "SItem");
struct SChunk { int32_t item, lo, hi, index; };   // one block's share of a huge node (index = chunk number inside the node)

struct SBins {          // shared-memory bins of one node (or one chunk of it)
    uint32_t mn[3][TAKE_SAH_BINS][3], mx[3][TAKE_SAH_BINS][3], cnt[3][TAKE_SAH_BINS];
};
struct HugeState {      // global per huge node of a level
    SBins bins;
    uint32_t child[2][12];      // per side: box min, box max, centroid min, centroid max (ordered uint)
    int32_t axis, best_bin, nl, first_chunk, n_chunks, pad[3];
};

__device__ __forceinline__ int bin_of(float c, float cb_lo, float scale) {
    const int k = (int)((c - cb_lo) * scale);
    return min(max(k, 0), TAKE_SAH_BINS - 1);
}
__device__ __forceinline__ float centroid(float l, float h) { return 0.5f * (l + h); }

__device__ __forceinline__ void write_node(BNode *nodes, int32_t id, const float *bl, const float *bh, int32_t count) {
    BNode b;
    b.lo = make_float4(bl[0], bl[1], bl[2], half_area3(bl, bh));
    b.hi = make_float4(bh[0], bh[1], bh[2], 0.0f);
    b.left = b.right = -1;
    b.count = count;
    b.leaf = count == 1 ? 1 : 0;
    nodes[id] = b;
}

__device__ __forceinline__ void item_scales(const SItem &it, float *scale3, bool *axis_ok) {
    for (int x = 0; x < 3; ++x) {
        const float ext = it.ch[x] - it.cl[x];
        axis_ok[x] = ext > 0.0f;
        scale3[x] = axis_ok[x] ? (float)TAKE_SAH_BINS / ext : 0.0f;
    }
}

__device__ __forceinline__ void bins_clear(SBins &B) {
    for (int t = threadIdx.x; t < 3 * TAKE_SAH_BINS; t += blockDim.x) {
        const int x = t / TAKE_SAH_BINS, b = t % TAKE_SAH_BINS;
        for (int a = 0; a < 3; ++a) { B.mn[x][b][a] = 0xffffffffu; B.mx[x][b][a] = 0u; }
        B.cnt[x][b] = 0u;
    }
}

// Bin [lo, hi) on all three axes into shared bins.  Consecutive primitives of a mesh often sit next to each other in space,
// so all lanes of a warp frequently fall into ONE bin -- same-address shared atomics would then serialise (measured: seconds
// at the top levels of a 10 M-primitive scene).  When a whole warp agrees on the bin (__match_all_sync, one instruction) the
// warp reduces its boxes with redux and one lane updates the bin; otherwise the lanes are spread over bins and update them
// themselves.
__device__ __forceinline__ void bin_range(SBins &B, const SItem &it, const float *scale3, const bool *axis_ok, int32_t lo, int32_t hi,
                                          const float4 *cur_lo, const float4 *cur_hi) {
    const int lane = threadIdx.x & 31;
    for (int32_t base = lo; base < hi; base += blockDim.x) {
        const int32_t i = base + threadIdx.x;
        const bool valid = i < hi;
        const unsigned vm = __ballot_sync(0xffffffffu, valid);
        if (!valid) continue;
        const float4 l = cur_lo[i], h = cur_hi[i];
        const float bl[3] = {l.x, l.y, l.z}, bh[3] = {h.x, h.y, h.z};
        uint32_t omin[3], omax[3];
        for (int a = 0; a < 3; ++a) { omin[a] = f2ord(bl[a]); omax[a] = f2ord(bh[a]); }
        for (int x = 0; x < 3; ++x) {
            if (!axis_ok[x]) continue;
            const int k = bin_of(centroid(bl[x], bh[x]), it.cl[x], scale3[x]);
            int same = 0;
            __match_all_sync(vm, k, &same);
            if (same) {
                uint32_t gmin[3], gmax[3];
                for (int a = 0; a < 3; ++a) { gmin[a] = __reduce_min_sync(vm, omin[a]); gmax[a] = __reduce_max_sync(vm, omax[a]); }
                if (lane == __ffs(vm) - 1) {
                    atomicAdd(&B.cnt[x][k], (uint32_t)__popc(vm));
                    for (int a = 0; a < 3; ++a) { atomicMin(&B.mn[x][k][a], gmin[a]); atomicMax(&B.mx[x][k][a], gmax[a]); }
                }
            } else {
                atomicAdd(&B.cnt[x][k], 1u);
                for (int a = 0; a < 3; ++a) { atomicMin(&B.mn[x][k][a], omin[a]); atomicMax(&B.mx[x][k][a], omax[a]); }
            }
        }
    }
}

// Sweep of one axis (split between bin b and b+1; the first minimum wins, like the host's ascending loops).
__device__ inline void sweep_axis(const SBins &B, int x, bool ok, float &best, int &best_bin, int &nleft) {
    best = INFINITY; best_bin = -1; nleft = 0;
    if (!ok) return;
    float right_area[TAKE_SAH_BINS];
    uint32_t right_cnt[TAKE_SAH_BINS];
    float al[3] = {INFINITY, INFINITY, INFINITY}, ah[3] = {-INFINITY, -INFINITY, -INFINITY};
    uint32_t cnt = 0;
    for (int b = TAKE_SAH_BINS - 1; b > 0; --b) {
        cnt += B.cnt[x][b];
        if (B.cnt[x][b]) for (int a = 0; a < 3; ++a) { al[a] = fminf(al[a], ord2f(B.mn[x][b][a])); ah[a] = fmaxf(ah[a], ord2f(B.mx[x][b][a])); }
        right_area[b] = half_area3(al, ah);
        right_cnt[b] = cnt;
    }
    for (int a = 0; a < 3; ++a) { al[a] = INFINITY; ah[a] = -INFINITY; }
    cnt = 0;
    for (int b = 0; b < TAKE_SAH_BINS - 1; ++b) {
        cnt += B.cnt[x][b];
        if (B.cnt[x][b]) for (int a = 0; a < 3; ++a) { al[a] = fminf(al[a], ord2f(B.mn[x][b][a])); ah[a] = fmaxf(ah[a], ord2f(B.mx[x][b][a])); }
        if (cnt == 0 || right_cnt[b + 1] == 0) continue;
        const float cost = half_area3(al, ah) * (float)cnt + right_area[b + 1] * (float)right_cnt[b + 1];
        if (cost < best) { best = cost; best_bin = b; nleft = (int)cnt; }
    }
}

// The three sweeps on threads 0..2 and the choice of the axis on thread 0; results in shared s_axis / s_best_bin / s_nl.
__device__ __forceinline__ void choose_split(const SBins &B, const bool *axis_ok, int32_t m, float *s_cost, int *s_bin, int *s_nleft,
                                             int &s_axis, int &s_best_bin, int &s_nl) {
    if (threadIdx.x < 3) sweep_axis(B, threadIdx.x, axis_ok[threadIdx.x], s_cost[threadIdx.x], s_bin[threadIdx.x], s_nleft[threadIdx.x]);
    __syncthreads();
    if (threadIdx.x == 0) {
        int axis = -1;
        float best = INFINITY;
        for (int x = 0; x < 3; ++x) if (s_bin[x] >= 0 && s_cost[x] < best) { best = s_cost[x]; axis = x; }
        s_axis = axis;
        s_best_bin = axis >= 0 ? s_bin[axis] : -1;
        s_nl = axis >= 0 ? s_nleft[axis] : m / 2;   // all centroids coincide: split the range in half, order kept
    }
    __syncthreads();
}

// Stable partition of [lo, hi) (a node's range or one chunk of it) into the other buffer.  Lefts go to out_l + rank, rights to
// out_r + rank; the bounds of both sides are accumulated in s_child with warp redux + one shared atomic per warp and quantity.
// node_lo / nl: start of the whole node's range and its left count (for the split-in-half fallback of coincident centroids).
__device__ __forceinline__ void partition_range(const SItem &it, const float *scale3, int axis, int best_bin, int32_t node_lo, int32_t nl,
                                                int32_t lo, int32_t hi, int32_t out_l, int32_t out_r, const float4 *cur_lo, const float4 *cur_hi,
                                                float4 *nxt_lo, float4 *nxt_hi, uint32_t (*s_child)[12], uint32_t *s_wl, uint32_t *s_wv) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int32_t lrun = 0, rrun = 0;
    for (int32_t base = lo; base < hi; base += TAKE_DB_BLOCK) {
        const int32_t i = base + tid;
        const bool valid = i < hi;
        float4 l = make_float4(0, 0, 0, 0), h = l;
        bool left = false;
        if (valid) {
            l = cur_lo[i]; h = cur_hi[i];
            if (axis >= 0) {
                const float bl = axis == 0 ? l.x : axis == 1 ? l.y : l.z, bh = axis == 0 ? h.x : axis == 1 ? h.y : h.z;
                left = bin_of(centroid(bl, bh), it.cl[axis], scale3[axis]) <= best_bin;
            } else {
                left = (i - node_lo) < nl;
            }
        }
        const unsigned wmask = __ballot_sync(0xffffffffu, left), vmask = __ballot_sync(0xffffffffu, valid);
        if (lane == 0) { s_wl[warp] = __popc(wmask); s_wv[warp] = __popc(vmask); }
        __syncthreads();
        int32_t lpre = 0, vpre = 0, ltot = 0, vtot = 0;
        for (int w = 0; w < TAKE_DB_BLOCK / 32; ++w) {
            if (w < warp) { lpre += s_wl[w]; vpre += s_wv[w]; }
            ltot += s_wl[w]; vtot += s_wv[w];
        }
        const unsigned lt = (1u << lane) - 1u;
        const int32_t lrank = lpre + __popc(wmask & lt), vrank = vpre + __popc(vmask & lt);
        if (valid) {
            const int32_t pos = left ? out_l + lrun + lrank : out_r + rrun + (vrank - lrank);
            nxt_lo[pos] = l; nxt_hi[pos] = h;
        }
        const float q[12] = {l.x, l.y, l.z, h.x, h.y, h.z, centroid(l.x, h.x), centroid(l.y, h.y), centroid(l.z, h.z),
                             centroid(l.x, h.x), centroid(l.y, h.y), centroid(l.z, h.z)};
        for (int side = 0; side < 2; ++side) {
            const bool mine = valid && (left == (side == 0));
            if (!__any_sync(0xffffffffu, mine)) continue;
            for (int k = 0; k < 12; ++k) {
                const bool is_min = (k % 6) < 3;
                const uint32_t v = mine ? f2ord(q[k]) : (is_min ? 0xffffffffu : 0u);
                const uint32_t r = is_min ? __reduce_min_sync(0xffffffffu, v) : __reduce_max_sync(0xffffffffu, v);
                if (lane == 0) { if (is_min) atomicMin(&s_child[side][k], r); else atomicMax(&s_child[side][k], r); }
            }
        }
        lrun += ltot; rrun += vtot - ltot;
        __syncthreads();
    }
}

// Children of a split node: allocate their ids, write them, queue them by size (huge nodes also reserve their chunks).
__device__ inline void emit_children(const SItem &it, int32_t nl, const uint32_t (*child)[12], BNode *nodes, Globals *g, SItem *next_huge,
                                     SItem *next_med, SItem *next_small, SChunk *next_chunks, HugeState *next_state) {
    const int32_t first = atomicAdd(&g->node_count, 2);
    nodes[it.node].left = first;
    nodes[it.node].right = first + 1;
    for (int side = 0; side < 2; ++side) {
        SItem c;
        c.node = first + side; c.pad = 0;
        c.lo = side == 0 ? it.lo : it.lo + nl;
        c.hi = side == 0 ? it.lo + nl : it.hi;
        for (int a = 0; a < 3; ++a) {
            c.bl[a] = ord2f(child[side][a]); c.bh[a] = ord2f(child[side][3 + a]);
            c.cl[a] = ord2f(child[side][6 + a]); c.ch[a] = ord2f(child[side][9 + a]);
        }
        const int32_t cm = c.hi - c.lo;
        write_node(nodes, c.node, c.bl, c.bh, cm);
        if (cm > TAKE_SAH_HUGE) {
            const int32_t slot = atomicAdd(&g->huge_count, 1);
            const int32_t nch = (cm + TAKE_SAH_CHUNK - 1) / TAKE_SAH_CHUNK;
            const int32_t c0 = atomicAdd(&g->chunk_count, nch);
            next_huge[slot] = c;
            next_state[slot].first_chunk = c0;
            next_state[slot].n_chunks = nch;
            for (int q = 0; q < nch; ++q) {
                SChunk ch;
                ch.item = slot; ch.index = q;
                ch.lo = c.lo + q * TAKE_SAH_CHUNK;
                ch.hi = min(c.hi, ch.lo + TAKE_SAH_CHUNK);
                next_chunks[c0 + q] = ch;
            }
        } else if (cm > TAKE_SAH_SMALL) {
            next_med[atomicAdd(&g->med_count, 1)] = c;
        } else {
            next_small[atomicAdd(&g->small_count, 1)] = c;
        }
    }
}

// root item from the global bounds (reuses emit's queueing rules)
__global__ void k_sah_root(int32_t n, Globals *g, BNode *nodes, SItem *huge, SItem *med, SItem *small, SChunk *chunks, HugeState *state) {
    if (blockIdx.x || threadIdx.x) return;
    SItem it;
    it.node = 0; it.lo = 0; it.hi = n; it.pad = 0;
    for (int a = 0; a < 3; ++a) {
        it.bl[a] = ord2f(g->bmin[a]); it.bh[a] = ord2f(g->bmax[a]);
        it.cl[a] = ord2f(g->cmin[a]); it.ch[a] = ord2f(g->cmax[a]);
    }
    write_node(nodes, 0, it.bl, it.bh, n);
    g->node_count = 1;
    g->huge_count = g->med_count = g->small_count = g->chunk_count = 0;
    if (n > TAKE_SAH_HUGE) {
        const int32_t nch = (n + TAKE_SAH_CHUNK - 1) / TAKE_SAH_CHUNK;
        huge[0] = it;
        state[0].first_chunk = 0; state[0].n_chunks = nch;
        for (int q = 0; q < nch; ++q) {
            SChunk ch;
            ch.item = 0; ch.index = q; ch.lo = q * TAKE_SAH_CHUNK; ch.hi = min(n, ch.lo + TAKE_SAH_CHUNK);
            chunks[q] = ch;
        }
        g->huge_count = 1; g->chunk_count = nch;
    } else if (n > TAKE_SAH_SMALL) {
        med[0] = it; g->med_count = 1;
    } else {
        small[0] = it; g->small_count = 1;
    }
}

// ---- medium nodes: one block per node: bin, sweep, partition (cur -> nxt), emit the children ----
__global__ void __launch_bounds__(TAKE_DB_BLOCK) k_sah_split(const SItem *items, const float4 *cur_lo, const float4 *cur_hi, float4 *nxt_lo,
                                                             float4 *nxt_hi, BNode *nodes, Globals *g, SItem *next_huge, SItem *next_med,
                                                             SItem *next_small, SChunk *next_chunks, HugeState *next_state) {
    __shared__ SBins B;
    __shared__ float s_cost[3];
    __shared__ int s_bin[3], s_nleft[3];
    __shared__ uint32_t s_child[2][12];
    __shared__ uint32_t s_wl[TAKE_DB_BLOCK / 32], s_wv[TAKE_DB_BLOCK / 32];
    __shared__ int s_axis, s_best_bin, s_nl;
    const SItem it = items[blockIdx.x];
    const int tid = threadIdx.x;
    float scale3[3];
    bool axis_ok[3];
    item_scales(it, scale3, axis_ok);
    bins_clear(B);
    if (tid < 24) s_child[tid / 12][tid % 12] = ((tid % 12) % 6 < 3) ? 0xffffffffu : 0u;
    __syncthreads();
    bin_range(B, it, scale3, axis_ok, it.lo, it.hi, cur_lo, cur_hi);
    __syncthreads();
    choose_split(B, axis_ok, it.hi - it.lo, s_cost, s_bin, s_nleft, s_axis, s_best_bin, s_nl);
    const int32_t nl = s_nl;
    partition_range(it, scale3, s_axis, s_best_bin, it.lo, nl, it.lo, it.hi, it.lo, it.lo + nl, cur_lo, cur_hi, nxt_lo, nxt_hi, s_child, s_wl, s_wv);
    if (tid == 0) emit_children(it, nl, s_child, nodes, g, next_huge, next_med, next_small, next_chunks, next_state);
}

// ---- huge nodes: several blocks per node (one per chunk of TAKE_SAH_CHUNK primitives) ----
__global__ void k_huge_clear(HugeState *state, int32_t n_items) {
    const int32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const int32_t item = t / (3 * TAKE_SAH_BINS), r = t % (3 * TAKE_SAH_BINS);
    if (item >= n_items) return;
    SBins &B = state[item].bins;
    const int x = r / TAKE_SAH_BINS, b = r % TAKE_SAH_BINS;
    for (int a = 0; a < 3; ++a) { B.mn[x][b][a] = 0xffffffffu; B.mx[x][b][a] = 0u; }
    B.cnt[x][b] = 0u;
    if (r < 24) state[item].child[r / 12][r % 12] = ((r % 12) % 6 < 3) ? 0xffffffffu : 0u;
}

// pass 1 of a chunk: bin into shared memory, merge into the node's global bins, keep the chunk's own bin counts
__global__ void __launch_bounds__(TAKE_DB_BLOCK) k_huge_bin(const SItem *items, const SChunk *chunks, const float4 *cur_lo, const float4 *cur_hi,
                                                            HugeState *state, uint32_t *chunk_cnt) {
    __shared__ SBins B;
    const SChunk ch = chunks[blockIdx.x];
    const SItem it = items[ch.item];
    float scale3[3];
    bool axis_ok[3];
    item_scales(it, scale3, axis_ok);
    bins_clear(B);
    __syncthreads();
    bin_range(B, it, scale3, axis_ok, ch.lo, ch.hi, cur_lo, cur_hi);
    __syncthreads();
    SBins &G = state[ch.item].bins;
    uint32_t *cc = chunk_cnt + (size_t)(state[ch.item].first_chunk + ch.index) * (3 * TAKE_SAH_BINS);
    for (int t = threadIdx.x; t < 3 * TAKE_SAH_BINS; t += TAKE_DB_BLOCK) {
        const int x = t / TAKE_SAH_BINS, b = t % TAKE_SAH_BINS;
        const uint32_t c = B.cnt[x][b];
        cc[t] = c;
        if (c) {
            atomicAdd(&G.cnt[x][b], c);
            for (int a = 0; a < 3; ++a) { atomicMin(&G.mn[x][b][a], B.mn[x][b][a]); atomicMax(&G.mx[x][b][a], B.mx[x][b][a]); }
        }
    }
}

// one block per huge node: choose the split from the merged bins, turn the chunks' bin counts into output offsets
// (chunk_off[c] = number of left-going primitives in the chunks before c of the same node)
__global__ void __launch_bounds__(TAKE_DB_BLOCK) k_huge_choose(const SItem *items, HugeState *state, const uint32_t *chunk_cnt, int32_t *chunk_off) {
    __shared__ float s_cost[3];
    __shared__ int s_bin[3], s_nleft[3];
    __shared__ int s_axis, s_best_bin, s_nl;
    const SItem it = items[blockIdx.x];
    HugeState &S = state[blockIdx.x];
    float scale3[3];
    bool axis_ok[3];
    item_scales(it, scale3, axis_ok);
    choose_split(S.bins, axis_ok, it.hi - it.lo, s_cost, s_bin, s_nleft, s_axis, s_best_bin, s_nl);
    if (threadIdx.x == 0) {
        S.axis = s_axis; S.best_bin = s_best_bin; S.nl = s_nl;
        int32_t acc = 0;
        for (int c = 0; c < S.n_chunks; ++c) {
            chunk_off[S.first_chunk + c] = acc;
            if (s_axis >= 0) {
                const uint32_t *cc = chunk_cnt + (size_t)(S.first_chunk + c) * (3 * TAKE_SAH_BINS) + s_axis * TAKE_SAH_BINS;
                for (int b = 0; b <= s_best_bin; ++b) acc += (int32_t)cc[b];
            } else {   // coincident centroids: the first nl primitives of the range go left
                const int32_t c_lo = it.lo + c * TAKE_SAH_CHUNK, c_hi = min(it.hi, c_lo + TAKE_SAH_CHUNK);
                acc += max(0, min(c_hi, it.lo + s_nl) - c_lo);
            }
        }
    }
}

// pass 2 of a chunk: stable partition into the node's output ranges, merge the children's bounds into the node's state
__global__ void __launch_bounds__(TAKE_DB_BLOCK) k_huge_partition(const SItem *items, const SChunk *chunks, const float4 *cur_lo, const float4 *cur_hi,
                                                                  float4 *nxt_lo, float4 *nxt_hi, HugeState *state, const int32_t *chunk_off) {
    __shared__ uint32_t s_child[2][12];
    __shared__ uint32_t s_wl[TAKE_DB_BLOCK / 32], s_wv[TAKE_DB_BLOCK / 32];
    const SChunk ch = chunks[blockIdx.x];
    const SItem it = items[ch.item];
    HugeState &S = state[ch.item];
    const int tid = threadIdx.x;
    float scale3[3];
    bool axis_ok[3];
    item_scales(it, scale3, axis_ok);
    if (tid < 24) s_child[tid / 12][tid % 12] = ((tid % 12) % 6 < 3) ? 0xffffffffu : 0u;
    __syncthreads();
    const int32_t left_before = chunk_off[S.first_chunk + ch.index];
    const int32_t right_before = (ch.lo - it.lo) - left_before;
    partition_range(it, scale3, S.axis, S.best_bin, it.lo, S.nl, ch.lo, ch.hi, it.lo + left_before, it.lo + S.nl + right_before, cur_lo, cur_hi,
                    nxt_lo, nxt_hi, s_child, s_wl, s_wv);
    if (tid < 24) {
        const int side = tid / 12, k = tid % 12;
        if ((k % 6) < 3) atomicMin(&S.child[side][k], s_child[side][k]);
        else atomicMax(&S.child[side][k], s_child[side][k]);
    }
}

__global__ void k_huge_emit(const SItem *items, int32_t n_items, const HugeState *state, BNode *nodes, Globals *g, SItem *next_huge,
                            SItem *next_med, SItem *next_small, SChunk *next_chunks, HugeState *next_state) {
    const int32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_items) return;
    emit_children(items[t], state[t].nl, state[t].child, nodes, g, next_huge, next_med, next_small, next_chunks, next_state);
}

// One thread per node of up to SMALL primitives: copy the range into the final array and finish the subtree there (the
// host's sparse sweep for small ranges, in-place partition from both ends).
struct STask { int32_t node, lo, hi; float bl[3], bh[3], cl[3], ch[3]; };

__global__ void k_sah_small(const SItem *items, int32_t n_items, const float4 *cur_lo, const float4 *cur_hi, float4 *fin_lo, float4 *fin_hi,
                            BNode *nodes, Globals *g, int max_leaf, float c_trav, float c_isect) {
    const int32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_items) return;
    const SItem it = items[t];
    for (int32_t i = it.lo; i < it.hi; ++i) { fin_lo[i] = cur_lo[i]; fin_hi[i] = cur_hi[i]; }
    STask stack[TAKE_SAH_SMALL + 2];
    int sp = 0;
    {
        STask r;
        r.node = it.node; r.lo = it.lo; r.hi = it.hi;
        for (int a = 0; a < 3; ++a) { r.bl[a] = it.bl[a]; r.bh[a] = it.bh[a]; r.cl[a] = it.cl[a]; r.ch[a] = it.ch[a]; }
        stack[sp++] = r;
    }
    while (sp > 0) {
        const STask k = stack[--sp];
        const int32_t lo = k.lo, hi = k.hi, m = hi - lo;
        if (m <= 1) { nodes[k.node].leaf = 1; continue; }
        float best_cost = INFINITY;
        int best_axis = -1, best_bin = -1;
        float scale3[3];
        for (int x = 0; x < 3; ++x) {
            const float ext = k.ch[x] - k.cl[x];
            scale3[x] = ext > 0.0f ? (float)TAKE_SAH_BINS / ext : 0.0f;
            if (!(ext > 0.0f)) continue;
            int key[TAKE_SAH_SMALL], ord[TAKE_SAH_SMALL];
            for (int i = 0; i < m; ++i) {
                const float4 l = fin_lo[lo + i], h = fin_hi[lo + i];
                const float bl = x == 0 ? l.x : x == 1 ? l.y : l.z, bh = x == 0 ? h.x : x == 1 ? h.y : h.z;
                key[i] = bin_of(centroid(bl, bh), k.cl[x], scale3[x]);
                int j = i;
                while (j > 0 && key[ord[j - 1]] > key[i]) { ord[j] = ord[j - 1]; --j; }
                ord[j] = i;
            }
            float right_area[TAKE_SAH_SMALL + 1];
            float al[3] = {INFINITY, INFINITY, INFINITY}, ah[3] = {-INFINITY, -INFINITY, -INFINITY};
            for (int i = m - 1; i >= 0; --i) {
                const float4 l = fin_lo[lo + ord[i]], h = fin_hi[lo + ord[i]];
                al[0] = fminf(al[0], l.x); al[1] = fminf(al[1], l.y); al[2] = fminf(al[2], l.z);
                ah[0] = fmaxf(ah[0], h.x); ah[1] = fmaxf(ah[1], h.y); ah[2] = fmaxf(ah[2], h.z);
                right_area[i] = half_area3(al, ah);
            }
            for (int a = 0; a < 3; ++a) { al[a] = INFINITY; ah[a] = -INFINITY; }
            for (int i = 0; i < m - 1; ++i) {
                const float4 l = fin_lo[lo + ord[i]], h = fin_hi[lo + ord[i]];
                al[0] = fminf(al[0], l.x); al[1] = fminf(al[1], l.y); al[2] = fminf(al[2], l.z);
                ah[0] = fmaxf(ah[0], h.x); ah[1] = fmaxf(ah[1], h.y); ah[2] = fmaxf(ah[2], h.z);
                if (key[ord[i + 1]] == key[ord[i]]) continue;   // inside a bin group: not a split position
                const float cost = half_area3(al, ah) * (float)(i + 1) + right_area[i + 1] * (float)(m - i - 1);
                if (cost < best_cost) { best_cost = cost; best_axis = x; best_bin = key[ord[i]]; }
            }
        }
        int32_t mid = -1;
        float cl_[2][3], ch_[2][3], bl_[2][3], bh_[2][3];
        for (int s2 = 0; s2 < 2; ++s2) for (int a = 0; a < 3; ++a) { cl_[s2][a] = bl_[s2][a] = INFINITY; ch_[s2][a] = bh_[s2][a] = -INFINITY; }
        auto grow = [&](int side, float4 l, float4 h) {
            const float pl[3] = {l.x, l.y, l.z}, ph[3] = {h.x, h.y, h.z};
            for (int a = 0; a < 3; ++a) {
                bl_[side][a] = fminf(bl_[side][a], pl[a]); bh_[side][a] = fmaxf(bh_[side][a], ph[a]);
                const float c = centroid(pl[a], ph[a]);
                cl_[side][a] = fminf(cl_[side][a], c); ch_[side][a] = fmaxf(ch_[side][a], c);
            }
        };
        if (best_axis >= 0) {
            const float parent_area = half_area3(k.bl, k.bh);
            const float split_cost = c_trav + c_isect * best_cost / (parent_area > 0.0f ? parent_area : 1.0f);
            const float leaf_cost = c_isect * (float)m;
            if (m <= max_leaf && leaf_cost <= split_cost) { nodes[k.node].leaf = 1; continue; }
            auto goes_left = [&](int32_t i) {
                const float4 l = fin_lo[i], h = fin_hi[i];
                const float bl = best_axis == 0 ? l.x : best_axis == 1 ? l.y : l.z, bh = best_axis == 0 ? h.x : best_axis == 1 ? h.y : h.z;
                return bin_of(centroid(bl, bh), k.cl[best_axis], scale3[best_axis]) <= best_bin;
            };
            int32_t i = lo, j = hi - 1;
            for (;;) {
                while (i <= j && goes_left(i)) ++i;
                while (i <= j && !goes_left(j)) --j;
                if (i >= j) break;
                const float4 a0 = fin_lo[i], a1 = fin_hi[i];
                fin_lo[i] = fin_lo[j]; fin_hi[i] = fin_hi[j];
                fin_lo[j] = a0; fin_hi[j] = a1;
                ++i; --j;
            }
            mid = i;
        }
        if (mid <= lo || mid >= hi) {
            if (m <= max_leaf) { nodes[k.node].leaf = 1; continue; }
            mid = lo + m / 2;
        }
        for (int32_t i = lo; i < hi; ++i) grow(i < mid ? 0 : 1, fin_lo[i], fin_hi[i]);
        const int32_t first = atomicAdd(&g->node_count, 2);
        nodes[k.node].left = first;
        nodes[k.node].right = first + 1;
        for (int side = 1; side >= 0; --side) {    // (left child is processed first: popped last-in first-out)
            STask c;
            c.node = first + side;
            c.lo = side == 0 ? lo : mid;
            c.hi = side == 0 ? mid : hi;
            for (int a = 0; a < 3; ++a) { c.bl[a] = bl_[side][a]; c.bh[a] = bh_[side][a]; c.cl[a] = cl_[side][a]; c.ch[a] = ch_[side][a]; }
            write_node(nodes, c.node, c.bl, c.bh, c.hi - c.lo);
            if (c.hi - c.lo > 1) stack[sp++] = c;
        }
    }
}

// primitive id of every leaf slot (the final order of the SAH partitions)
__global__ void k_leaf_prims(int64_t n, const float4 *fin_lo, int32_t *leaf_prims) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) leaf_prims[i] = __float_as_int(fin_lo[i].w);
}

// ---- 3. wide collapse (breadth first) -------------------------------------------------------------------------------
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

// pass B: write the wide node and the work items of its inner children
__global__ void k_wide_emit(int32_t n_items, const WorkItem *items, const Kids *kids, const uint32_t *inner_scan, const BNode *nodes,
                            WideNode *wide, int32_t next_wide, WorkItem *next_items) {
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
__global__ void k_wide_wrap_root(const BNode *nodes, int32_t root, int64_t n, WideNode *wide) {
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
    }
    wide[0] = w;
}

// ---- 4. leaf records ----------------------------------------------------------------------------------------------
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
