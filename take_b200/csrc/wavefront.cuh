// Wavefront path tracing: generate -> [extend -> sort-by-material -> shade -> shadow-connect] x bounces -> accumulate.
// Replaces the per-pixel loop of src/render.cpp:59-82 and the integrators of src/integrator/path_tracing.h.
//
// A "wave" is a batch of path samples, one per slot.  Per-slot state lives in 32-byte-multiple records (one or two
// full sectors per gathered access); queues are arrays of slot indices filled through warp-aggregated atomics.
// All queue lengths stay on the device: each pass `b` owns its own counters (no resets, no host round trips), and the
// traversal kernels are persistent -- warps pull batches of 32 queue entries until the queue is dry -- so a fixed
// launch configuration serves every pass.
#pragma once
#include "shading.cuh"
#include "traverse.cuh"

namespace take {

struct __align__(16) RayRec {  // 64 B : the extend ray of the slot (tmin is always c_EPSILON)
    double ox, oy, oz, dx, dy, dz, tmax;
    int32_t aux0, aux1;
};
struct __align__(16) HitRec {  // 32 B
    int32_t prim;
    uint32_t keyrank;  // sort key (TAKE_KEY_BITS) | rank inside the key's bin (the remaining low bits)
    double t, u, v;
};
struct __align__(16) PathRec {  // 64 B = two sectors: the radiance sector is all that k_shadow, k_accumulate and a finished path touch
    double rad[3];
    uint32_t k;      // random_real draws consumed so far
    int32_t depth;   // index of the next integrator loop iteration
    double thr[3];
    int32_t flags;   // PEND_* describing the extend ray in flight
    int32_t pad;
};
struct __align__(16) PendRec {  // 32 B : BSDF sample waiting for its extend ray (FG and pdf of src/integrator/path_tracing.h:70-73)
    double fg[3], bpdf;
};
struct __align__(16) ShadowRec {  // 64 B : NEE connection waiting for its shadow ray (origin = RayRec.o)
    double dx, dy, dz, tmax, cx, cy, cz, pad;
};
static_assert(sizeof(RayRec) == 64 && sizeof(HitRec) == 32 && sizeof(PathRec) == 64 && sizeof(PendRec) == 32 &&
                  sizeof(ShadowRec) == 64, "record sizes");

// Streaming (evict-first) copies of whole records for the traversal kernels: the per-path records are touched once per
// pass, the tree is re-read by every ray, so the records should not push the tree out of L2.
#ifndef TAKE_STREAM_HINTS
#define TAKE_STREAM_HINTS 1
#endif
template <typename T>
__device__ __forceinline__ T ld_stream(const T *p) {
    static_assert(sizeof(T) % 16 == 0, "records are multiples of 16 bytes");
#if TAKE_STREAM_HINTS
    T r;
    const int4 *s = reinterpret_cast<const int4 *>(p);
    int4 *d = reinterpret_cast<int4 *>(&r);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(T) / 16); ++i) d[i] = __ldcs(s + i);
    return r;
#else
    return *p;
#endif
}
template <typename T>
__device__ __forceinline__ void st_stream(T *p, const T &v) {
#if TAKE_STREAM_HINTS
    const int4 *s = reinterpret_cast<const int4 *>(&v);
    int4 *d = reinterpret_cast<int4 *>(p);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(T) / 16); ++i) __stcs(d + i, s[i]);
#else
    *p = v;
#endif
}

enum { PEND_PRIMARY = 0, PEND_BSDF = 1, PEND_LIGHT = 2, PEND_SPECULAR = 4, PEND_ENV = 8 };

// Shading sort key: 0 = miss, else (1 + material type) | (branch << 4).  `branch` is the one-sample-MIS integrator's
// light-or-BSDF coin of the vertex about to be shaded (path_tracing.h:187): the coin is the next draw of the sample's
// counter-based stream, so the kernel that emits the ray can look at it ahead of time (RayRec.aux0) and the sort makes
// warps uniform in it as well as in the material -- without it half of every warp idles through the other branch.
#define TAKE_KEY_BITS 5
#define TAKE_RANK_BITS (32 - TAKE_KEY_BITS)
#define TAKE_RANK_MASK ((1u << TAKE_RANK_BITS) - 1u)
#define TAKE_NBINS (1 << TAKE_KEY_BITS)
__device__ __forceinline__ uint32_t sort_key(int32_t prim, const uint8_t *prim_mtype, int32_t branch) {
    return prim < 0 ? 0u : ((1u + (uint32_t)prim_mtype[prim]) | ((uint32_t)branch << 4));
}
#define TAKE_MAX_DEPTH 65533  // sanity bound on -max_depth (a wave runs max_depth + 2 passes)

struct PassCounters {  // one per pass, zeroed once per wave
    uint32_t n_extend;          // entries in this pass's extend queue
    uint32_t n_shadow;          // entries in this pass's shadow queue
    uint32_t fetch_extend;      // persistent-kernel work cursors
    uint32_t fetch_shadow;
    uint32_t bins[TAKE_NBINS];  // histogram of sort keys
    uint32_t fetch_sort;        // work cursor of the ordered scatter (k_scatter_ordered)
    uint32_t pad0[27];
    uint32_t fill[TAKE_NBINS];  // entries of each bin the ordered scatter has placed so far
    uint32_t pad1[32];
};
static_assert(sizeof(PassCounters) == 512 && TAKE_NBINS == 32, "PassCounters");

struct Totals {  // running totals over a render call
    unsigned long long samples, extend_rays, shadow_rays, shaded, box_tests, tri_tests, miss_after_light_sample;
    unsigned long long shadow_box_tests, shadow_tri_tests, pad[7];
};

struct Wave {
    RayRec *ray;
    HitRec *hit;         // by slot (written by extend)
    HitRec *hit_sorted;  // in material order (copied by the sort, `keyrank` replaced by the slot): shade streams through it
    PathRec *path;
    PendRec *pend;
    ShadowRec *shadow;
    int32_t *q_extend[2];
    int32_t *q_shadow;
    PassCounters *pass;  // [max_depth + 3]
    Totals *totals;
    // slot -> (pixel, sample): see slot_of() below, unless an explicit list is given
    int32_t chunk_pixels, chunk_base;
    int32_t gs_log2;  // log2 of the samples of one pixel that share a warp (0: a warp = 32 pixels of one sample index)
    int64_t sample0;
    int32_t n_slots, samples_in_wave;
    const int32_t *list_pixel;  // optional explicit (pixel, sample) list (take_gpu_radiance_samples)
    const int64_t *list_sample;
    int32_t integrator, max_depth, sort_enabled;
    int32_t sort_branch;    // 1: the sort key carries the one-sample integrator's light/BSDF coin (RayRec.aux0)
    int32_t fused_primary;  // 1: no k_generate -- pass 0 of extend and shade compute the camera ray themselves
    int32_t miss_fast;      // 1: camera rays that leave the scene are finished by k_extend itself (see k_extend)
    int32_t packet;         // 1: pass 0 runs k_extend_primary (the warp's camera rays traverse as a packet)
    int32_t count_ties;     // 1: a render ahead of the tie-break ranks: the extend kernels count rank-decided ties
    int32_t rr_start;       // EXTENSION: Russian roulette from this loop iteration on (0 = off, the reference's behaviour)
    int32_t ordered_sort;   // 1: k_scatter_ordered places the hits of a bin in queue order (see there); 0: in the order the extend kernel finished them
    int32_t tile_w;  // > 0: image width, pixels are enumerated in 8x4 tiles (one warp = one tile); 0: row-major
    uint64_t seed;
};

// Pixel enumeration.  Consecutive indices walk 8x4 tiles so that the 32 primary rays of a warp form a compact bundle
// (one warp = one tile) instead of a 32x1 strip; any bijection is fine for the result because a sample's random
// stream is keyed by the pixel, not by the slot.
__device__ __forceinline__ uint32_t pixel_of_index(const Wave &w, uint32_t idx) {
    if (w.tile_w <= 0) return idx;
    const uint32_t tiles_x = (uint32_t)w.tile_w >> 3;
    const uint32_t tile = idx >> 5, in = idx & 31u;
    const uint32_t ty = tile / tiles_x, tx = tile - ty * tiles_x;
    return (ty * 4u + (in >> 3)) * (uint32_t)w.tile_w + tx * 8u + (in & 7u);
}

// Slot layout of a wave.  A warp (32 consecutive slots) holds G_s = 2^gs_log2 samples of each of G_p = 32 / G_s
// neighbouring pixels: with G_s = 1 it is one sample index of a whole 8x4 tile, with larger G_s the camera rays of a warp
// differ by sub-pixel jitter only, follow (nearly) the same path through the tree and fetch the same nodes.  Warps are
// ordered pixel group first, so the sample groups of a pixel group are adjacent.  Any bijection gives the same image
// (streams are keyed by pixel and sample index); the host falls back to G_s = 1 when the wave's sample count or pixel
// count is not a multiple of the group sizes.
__device__ __forceinline__ void slot_to_local(const Wave &w, uint32_t slot, uint32_t &p_local, uint32_t &s_local) {
    if (w.gs_log2 == 0) {
        p_local = slot % (uint32_t)w.chunk_pixels;
        s_local = slot / (uint32_t)w.chunk_pixels;
        return;
    }
    const uint32_t gs = 1u << w.gs_log2, gp_log2 = 5u - (uint32_t)w.gs_log2, gp = 1u << gp_log2;
    const uint32_t sample_groups = (uint32_t)w.samples_in_wave >> w.gs_log2;
    const uint32_t wi = slot >> 5, lane = slot & 31u;
    const uint32_t pg = wi / sample_groups, sg = wi - pg * sample_groups;
    p_local = (pg << gp_log2) + (lane & (gp - 1u));
    s_local = sg * gs + (lane >> gp_log2);
}
__device__ __forceinline__ size_t slot_of(const Wave &w, uint32_t p_local, uint32_t s_local) {
    if (w.gs_log2 == 0) return (size_t)s_local * (size_t)w.chunk_pixels + p_local;
    const uint32_t gp_log2 = 5u - (uint32_t)w.gs_log2, gp = 1u << gp_log2, gs = 1u << w.gs_log2;
    const uint32_t sample_groups = (uint32_t)w.samples_in_wave >> w.gs_log2;
    const size_t wi = (size_t)(p_local >> gp_log2) * sample_groups + (s_local >> w.gs_log2);
    return (wi << 5) + ((s_local & (gs - 1u)) << gp_log2) + (p_local & (gp - 1u));
}

__device__ __forceinline__ void slot_identity(const Wave &w, int slot, uint32_t &pixel, uint64_t &sample) {
    if (w.list_pixel) {
        pixel = (uint32_t)w.list_pixel[slot];
        sample = (uint64_t)w.list_sample[slot];
    } else {
        uint32_t p_local, s_local;
        slot_to_local(w, (uint32_t)slot, p_local, s_local);
        pixel = pixel_of_index(w, (uint32_t)w.chunk_base + p_local);
        sample = (uint64_t)(w.sample0 + s_local);
    }
}

// Append to a device queue with one atomic per warp: lanes that want to push are ranked with ballot/popc,
// the first of them reserves the range, shfl broadcasts the base.
__device__ __forceinline__ void queue_push(bool want, int32_t *queue, uint32_t *counter, int32_t value) {
    const unsigned active = __activemask();
    const unsigned mask = __ballot_sync(active, want);
    if (mask == 0) return;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(mask) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(counter, (uint32_t)__popc(mask));
    base = __shfl_sync(active, base, leader);
    if (want) queue[base + __popc(mask & ((1u << lane) - 1u))] = value;
}

#ifndef TAKE_V1_MIN_BLOCKS
#define TAKE_V1_MIN_BLOCKS 1
#endif
#ifndef TAKE_BOUNCE_MIN_BLOCKS   // resident blocks (-> register cap) of the one-ray-per-thread traversal kernels (bounce / shadow passes)
#define TAKE_BOUNCE_MIN_BLOCKS TAKE_V1_MIN_BLOCKS
#endif

// Camera ray of a slot: src/render.cpp:69-75 (jittered pinhole; first draw -> x, second -> y).  Consumes 2 draws.
// The coin the one-sample integrator will flip first at the vertex this stream reaches next: draw number rng.k.
// (`skip`: the iteration the ray arrives in starts with a Russian-roulette draw, so the coin is the draw after it)
__device__ __forceinline__ int32_t peek_branch(Rng rng, bool skip = false) {
    if (skip) rng.next();
    return rng.next() <= 0.5 ? 1 : 0;
}

__device__ __forceinline__ void primary_ray(const DevScene &sc, const Wave &w, int slot, D3 &o, D3 &dir, Rng &rng) {
    uint32_t pixel;
    uint64_t sample;
    slot_identity(w, slot, pixel, sample);
    const int col = (int)(pixel % (uint32_t)sc.width), row = (int)(pixel / (uint32_t)sc.width);
    const int x = col, y = sc.height - 1 - row;  // the reference's y-up loop variable; image row = H - y - 1
    rng.seed = w.seed; rng.sample = sample; rng.pixel = pixel; rng.k = 0;
    const double jx = rng.next();
    const double jy = rng.next();
    dir = sub(add(mul(mul(sc.cam_u, (x + jx) / sc.width - 0.5), sc.viewport_w),
                  mul(mul(sc.cam_v, (y + jy) / sc.height - 0.5), sc.viewport_h)),
              sc.cam_w);
    dir = normalize(dir);
    o = sc.lookfrom;
}
__device__ __forceinline__ void primary_ray(const DevScene &sc, const Wave &w, int slot, D3 &o, D3 &dir) {
    Rng rng;
    primary_ray(sc, w, slot, o, dir, rng);
}

// Length of pass `pass`'s extend queue.  With fused primaries pass 0 has no queue: entry i is slot i.
__device__ __forceinline__ uint32_t pass_count(const Wave &w, int pass) {
    return (pass == 0 && w.fused_primary) ? (uint32_t)w.n_slots : w.pass[pass].n_extend;
}

// Number of entries the sort and the shade kernel of pass `pass` work on.  With `miss_fast` pass 0 keeps only the camera
// rays that hit something: k_extend lists their slots in q_extend[0] and counts them in pass[0].n_extend (which the
// fused camera-ray pass does not otherwise use).
__device__ __forceinline__ uint32_t shade_count(const Wave &w, int pass) {
    return (pass == 0 && w.miss_fast) ? w.pass[0].n_extend : pass_count(w, pass);
}

#if TAKE_EXPERIMENTAL
// ---- generate: src/render.cpp:65-75 (only when the primaries are not fused into pass 0) ----------------------
__global__ void k_generate(DevScene sc, Wave w) {
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot == 0) w.pass[0].n_extend = (uint32_t)w.n_slots;
    if (slot >= w.n_slots) return;
    D3 org, dir;
    Rng rng;
    primary_ray(sc, w, slot, org, dir, rng);
    RayRec r;
    r.ox = org.x; r.oy = org.y; r.oz = org.z;
    r.dx = dir.x; r.dy = dir.y; r.dz = dir.z;
    r.tmax = INFINITY;
    r.aux0 = w.sort_branch ? peek_branch(rng) : 0;
    r.aux1 = 0;
    w.ray[slot] = r;
    PathRec p;
    p.thr[0] = p.thr[1] = p.thr[2] = 1.0;
    p.rad[0] = p.rad[1] = p.rad[2] = 0.0;
    p.k = 2;
    p.depth = 0;
    p.flags = PEND_PRIMARY;
    p.pad = 0;
    w.path[slot] = p;
    w.q_extend[0][slot] = slot;
}

#endif  // TAKE_EXPERIMENTAL (separate generate pass)

// Work distribution of the persistent traversal kernels: a warp takes `batches` batches of 32 queue entries with ONE
// atomicAdd on the pass's cursor and works through them.  With one batch per atomic every warp hits the same address
// once per 32 rays; at 12 G camera rays/s that is a same-address atomic every 2.5 ns and the wait for the returned value
// shows up as 27 % of the stall samples of pass 0 (ncu, r01c).  Measured (tools/tune.py): 4 batches per atomic make the
// camera-ray pass 4 % faster, but every other pass SLOWER (-6 % on the first bounce, 2x on the thin late passes): rays of
// very different lengths need the fine granularity for balance.  So only the camera-ray pass uses TAKE_FETCH_BATCHES.
// (Reserving the next single batch ahead of time instead was also measured: -2 %.)
#ifndef TAKE_FETCH_BATCHES
#define TAKE_FETCH_BATCHES 4
#endif
struct WorkCursor {
    uint32_t cur, end;
    // warp-uniform: false when the queue is exhausted; `base` = first entry of the next batch of 32
    __device__ __forceinline__ bool next(uint32_t *cursor, uint32_t n, int lane, uint32_t &base, uint32_t batches = 1) {
        if (cur >= end) {
            uint32_t b = 0;
            if (lane == 0) b = atomicAdd(cursor, 32u * batches);
            cur = __shfl_sync(0xffffffffu, b, 0);
            if (cur >= n) return false;
            end = min(cur + 32u * batches, n);
        }
        base = cur;
        cur += 32u;
        return true;
    }
};

// ---- extend: closest hit for every queued ray + histogram of the shading sort key ------------------------
// What happens to a traced ray (shared by the per-thread kernel and the camera-ray packet kernel); called by whole warps.
__device__ __forceinline__ void extend_finish(const DevScene &sc, const Wave &w, PassCounters &pc, int lane, bool valid, int slot, const HitOut &h,
                                              int32_t branch, bool primary, bool miss_fast) {
    uint32_t key = 0;
    bool rec = false;  // this lane has a hit record for the sort
    if (valid) {
        if (miss_fast && h.prim < 0) {
            // A camera ray that left the scene: the sample is the background colour (path_tracing.h:8, :117, :164 -- the same
            // line of the other two integrators).  Finishing it here keeps it out of the sort and the shade pass --
            // on an open scene most camera rays end this way, and streaming their records through two more kernels
            // was the larger part of the pass-0 shade time.
            // (only the radiance sector of the record: nothing reads the rest of a finished path)
            int4 *ps = reinterpret_cast<int4 *>(w.path + slot);
            __stcs(ps, make_int4(__double2loint(sc.background.x), __double2hiint(sc.background.x), __double2loint(sc.background.y),
                                 __double2hiint(sc.background.y)));
            __stcs(ps + 1, make_int4(__double2loint(sc.background.z), __double2hiint(sc.background.z), 2, 0));
        } else {
            rec = true;
            if (primary && w.sort_branch) {  // the coin of the first vertex: draw number 2 of the sample's stream
                Rng rng;
                uint32_t pixel;
                uint64_t sample;
                slot_identity(w, slot, pixel, sample);
                rng.seed = w.seed; rng.sample = sample; rng.pixel = pixel; rng.k = 2;
                branch = peek_branch(rng);
            }
            key = sort_key(h.prim, sc.prim_mtype, branch);
        }
    }
    // warp-aggregated histogram: lanes with the same key elect a leader that bumps the bin once
    const unsigned vmask = __ballot_sync(0xffffffffu, rec);
    if (rec) {
        const unsigned peers = __match_any_sync(vmask, key);
        const int leader = __ffs(peers) - 1;
        uint32_t rbase = 0;
        if (lane == leader) rbase = atomicAdd(&pc.bins[key], (uint32_t)__popc(peers));
        rbase = __shfl_sync(peers, rbase, leader);
        const uint32_t rank = rbase + __popc(peers & ((1u << lane) - 1u));
        HitRec hr;
        hr.prim = h.prim;
        hr.keyrank = (key << TAKE_RANK_BITS) | rank;
        hr.t = h.t; hr.u = h.u; hr.v = h.v;
        st_stream(w.hit + slot, hr);
    }
    if (miss_fast && vmask) {  // list the surviving slots for the sort (one atomic per warp)
        uint32_t qb = 0;
        if (lane == __ffs(vmask) - 1) qb = atomicAdd(&pc.n_extend, (uint32_t)__popc(vmask));
        qb = __shfl_sync(0xffffffffu, qb, __ffs(vmask) - 1);
        if (rec) w.q_extend[0][qb + __popc(vmask & ((1u << lane) - 1u))] = slot;
    }
}

template <bool COUNT, bool WIDE, bool TIES = false>
__global__ void __launch_bounds__(128, TAKE_BOUNCE_MIN_BLOCKS) k_extend(DevScene sc, Wave w, int pass) {
    TAKE_DECLARE_STACK(st);
    PassCounters &pc = w.pass[pass];
    const uint32_t n = pass_count(w, pass);
    const bool primary = pass == 0 && w.fused_primary;
    const bool miss_fast = pass == 0 && w.miss_fast;
    const int32_t *queue = w.q_extend[pass & 1];
    const int lane = threadIdx.x & 31;
    TravCounters cnt = {0, 0};
    WorkCursor wc = {0, 0};
    for (;;) {
        uint32_t base;
        if (!wc.next(&pc.fetch_extend, n, lane, base, primary ? TAKE_FETCH_BATCHES : 1)) break;
        const uint32_t i = base + lane;
        const bool valid = i < n;
        int slot = -1;
        int32_t branch = 0;
        HitOut h;
        if (valid) {
            D3 o, d;
            double tmax = INFINITY;
            if (primary) {
                slot = (int)i;
                primary_ray(sc, w, slot, o, d);
            } else {
                slot = queue[i];
                const RayRec r = ld_stream(w.ray + slot);
                o = mk3(r.ox, r.oy, r.oz); d = mk3(r.dx, r.dy, r.dz); tmax = r.tmax;
                branch = r.aux0;
            }
            trace_any<false, COUNT, WIDE, TIES>(sc, o, d, TAKE_EPS, tmax, st, h, &cnt);
        }
        extend_finish(sc, w, pc, lane, valid, slot, h, branch, primary, miss_fast);
    }
    if (COUNT) {
        atomicAdd(&w.totals->box_tests, cnt.box);
        atomicAdd(&w.totals->tri_tests, cnt.tri);
    }
}

// Pass 0 with fused camera rays: the warp's 32 rays (the samples of one pixel, or of one 8x4 tile) traverse as a packet.
template <bool COUNT, bool TIES = false>
__global__ void __launch_bounds__(128, TAKE_V1_MIN_BLOCKS) k_extend_primary(DevScene sc, Wave w) {
    __shared__ uint2 pstack[128 / 32][TAKE_PACKET_STACK];
    PassCounters &pc = w.pass[0];
    const uint32_t n = (uint32_t)w.n_slots;
    const bool miss_fast = w.miss_fast != 0;
    const int lane = threadIdx.x & 31;
    uint2 *wstack = pstack[threadIdx.x >> 5];
    TravCounters cnt = {0, 0};
    WorkCursor wc = {0, 0};
    for (;;) {
        uint32_t base;
        if (!wc.next(&pc.fetch_extend, n, lane, base, TAKE_FETCH_BATCHES)) break;
        const uint32_t i = base + lane;
        const bool valid = i < n;
        D3 o = mk3(0, 0, 0), d = mk3(0, 0, 0);
        if (valid) primary_ray(sc, w, (int)i, o, d);
        HitOut h;
        trace_packet4<COUNT, TIES>(sc, o, d, TAKE_EPS, INFINITY, valid, wstack, h, &cnt);
        extend_finish(sc, w, pc, lane, valid, (int)i, h, 0, true, miss_fast);
    }
    if (COUNT) {
        atomicAdd(&w.totals->box_tests, cnt.box);
        atomicAdd(&w.totals->tri_tests, cnt.tri);
    }
}

// ---- sort: scatter slots into material-contiguous order ---------------------------------------------------
__global__ void k_scatter(Wave w, int pass) {
    __shared__ uint32_t offs[TAKE_NBINS];
    const PassCounters &pc = w.pass[pass];
    if (threadIdx.x == 0) {
        uint32_t acc = 0;
        for (int b = 0; b < TAKE_NBINS; ++b) { offs[b] = acc; acc += pc.bins[b]; }
    }
    __syncthreads();
    const uint32_t n = shade_count(w, pass);
    const bool primary = pass == 0 && w.fused_primary && !w.miss_fast;  // no list of slots: entry i is slot i
    const int32_t *queue = w.q_extend[pass & 1];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int slot = primary ? (int)i : queue[i];
        HitRec h = w.hit[slot];
        const uint32_t dst = offs[h.keyrank >> TAKE_RANK_BITS] + (h.keyrank & TAKE_RANK_MASK);
        h.keyrank = (uint32_t)slot;  // the sorted copy carries the slot in place of the key: shade needs no queue
        w.hit_sorted[dst] = h;
    }
}

// The same scatter with the position inside a bin assigned HERE, in queue order, instead of by the extend kernel in the
// order its rays happened to finish.  With one ray per thread the two nearly coincide (a warp finishes its 32 consecutive
// queue entries together); the lane-refill kernels finish rays in an order unrelated to the queue, and the shade kernel
// that then streams through a bin gathers its path / ray records from all over the wave (measured: shade +16 % on the
// 10 M-triangle scene, +41 % on the Cornell box).  Blocks take chunks of TAKE_SORT_CHUNK queue entries from a cursor, rank
// them per key in shared memory and reserve each key's share of its bin with one global atomic per key and chunk, so a bin
// is filled chunk by chunk in (nearly) queue order.  The order changes no result -- only where the records lie.
#define TAKE_SORT_CHUNK 1024
__global__ void __launch_bounds__(256) k_scatter_ordered(Wave w, int pass) {
    __shared__ uint32_t offs[TAKE_NBINS], cnt[TAKE_NBINS], base_of[TAKE_NBINS];
    __shared__ uint32_t chunk_base;
    PassCounters &pc = w.pass[pass];
    if (threadIdx.x == 0) {
        uint32_t acc = 0;
        for (int b = 0; b < TAKE_NBINS; ++b) { offs[b] = acc; acc += pc.bins[b]; }
    }
    const uint32_t n = shade_count(w, pass);
    const bool primary = pass == 0 && w.fused_primary && !w.miss_fast;
    const int32_t *queue = w.q_extend[pass & 1];
    const int lane = threadIdx.x & 31;
    constexpr int PER = TAKE_SORT_CHUNK / 256;
    for (;;) {
        __syncthreads();   // (also: offs ready; the previous chunk's cnt / base_of no longer needed)
        if (threadIdx.x == 0) chunk_base = atomicAdd(&pc.fetch_sort, (uint32_t)TAKE_SORT_CHUNK);
        if (threadIdx.x < TAKE_NBINS) cnt[threadIdx.x] = 0;
        __syncthreads();
        const uint32_t c0 = chunk_base;
        if (c0 >= n) break;
        HitRec h[PER];
        uint32_t key[PER], lrank[PER];
        int slot[PER];
#pragma unroll
        for (int j = 0; j < PER; ++j) {   // entry c0 + j * 256 + thread: a warp's 32 entries are consecutive
            const uint32_t i = c0 + (uint32_t)j * 256u + threadIdx.x;
            slot[j] = i < n ? (primary ? (int)i : queue[i]) : -1;
        }
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            if (slot[j] >= 0) h[j] = w.hit[slot[j]];
            key[j] = slot[j] >= 0 ? (h[j].keyrank >> TAKE_RANK_BITS) : 0xffffffffu;
        }
#pragma unroll
        for (int j = 0; j < PER; ++j) {   // rank inside the chunk: one shared atomic per distinct key and warp
            const unsigned peers = __match_any_sync(0xffffffffu, key[j]);
            const int leader = __ffs(peers) - 1;
            uint32_t b = 0;
            if (lane == leader && slot[j] >= 0) b = atomicAdd(&cnt[key[j]], (uint32_t)__popc(peers));
            b = __shfl_sync(0xffffffffu, b, leader);
            lrank[j] = b + __popc(peers & ((1u << lane) - 1u));
        }
        __syncthreads();
        if (threadIdx.x < TAKE_NBINS) {
            const uint32_t c = cnt[threadIdx.x];
            base_of[threadIdx.x] = c ? atomicAdd(&pc.fill[threadIdx.x], c) : 0u;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            if (slot[j] < 0) continue;
            const uint32_t dst = offs[key[j]] + base_of[key[j]] + lrank[j];
            h[j].keyrank = (uint32_t)slot[j];
            w.hit_sorted[dst] = h[j];
        }
    }
}

// ---- shade -----------------------------------------------------------------------------------------------
struct ShadeCtx {
    const DevScene &sc;
    const Wave &w;
    int slot, pass;
    D3 thr, rad;
    Rng rng;
    bool emit_extend, emit_shadow;
    D3 org;        // origin of the rays leaving this vertex (v.pos)
    D3 ext_dir;    // extend ray
    D3 sh_dir, sh_contrib;
    double sh_tmax;
    D3 pend_fg;
    double pend_pdf;
    int pend_flags;
    int depth;
    int shaded;    // integrator loop iterations entered (statistics)
};

// EXTENSION (README.md:19-24 of the reference lists Russian roulette as a goal; it has none): at the top of loop iteration
// c.depth, from rr_start on, the path survives with probability q = min(max component of the throughput, 0.95) and is divided
// by q -- one extra draw per iteration, the same in oracle/take_oracle.cpp (rr_survives).  false = the path ends here.
// RR = false compiles it out of the kernels that run the reference's integrators as they are (the check and its live state cost
// the shade kernels 1-2 %).
template <bool RR>
__device__ __forceinline__ bool russian_roulette(ShadeCtx &c) {
    if (!RR || c.w.rr_start <= 0 || c.depth < c.w.rr_start) return true;
    const double q = fmin(fmax(fmax(c.thr.x, c.thr.y), c.thr.z), 0.95);
    if (!(q > 0)) return false;
    if (c.rng.next() >= q) return false;
    c.thr = divs(c.thr, q);
    return true;
}

// NEE light-sample geometry shared by the integrators: path_tracing.h:31-43 / :189-200.
// Returns false when the reference `break`s (light_pdf <= 0).
__device__ __forceinline__ bool nee_sample(const DevScene &sc, const TakeLightDesc &l, int light_id, const Isect &v, Rng &rng,
                                           D3 &light_dir, double &dist, double &lpdf) {
    D3 lp, ln;
    sample_on_light(sc, light_id, l.prim_id, v.pos, rng, lp, ln);
    dist = length(sub(lp, v.pos));
    light_dir = normalize(sub(lp, v.pos));
    lpdf = light_pdf_area(sc, light_id, lp, v.pos) * (dist * dist) / (fmax(dot(neg(ln), light_dir), 0.0) * sc.pick_count);
    return !(lpdf <= 0);
}

// light pdf of a BSDF-sampled hit on an emitter: path_tracing.h:88-96 / :255-263
__device__ __forceinline__ bool hit_light_pdf(const DevScene &sc, const Isect &nv, D3 prev_pos, double &lpdf) {
    double d = length(sub(nv.pos, prev_pos));
    D3 light_dir = normalize(sub(nv.pos, prev_pos));
    lpdf = light_pdf_area(sc, nv.light, nv.pos, prev_pos) * (d * d) / (fmax(dot(neg(nv.gn), light_dir), 0.0) * sc.pick_count);
    return !(lpdf <= 0);
}

// Radiance carried by a ray that leaves the scene.  ENV = false compiles the environment-map extension out of the
// kernels used for scenes without a map (the reference's case), keeping their register budget.
template <bool ENV>
__device__ __forceinline__ D3 miss_rad(const DevScene &sc, D3 d) { return ENV ? env_radiance(sc, d) : sc.background; }

// Multi-sample MIS: src/integrator/path_tracing.h:5-111.  One call = "finish iteration depth-1 with the hit that
// just arrived, then run iteration depth up to the point where it needs rays".
template <bool ENV, bool RR>
__device__ __forceinline__ void shade_mis(ShadeCtx &c, const RayRec &ray, const HitRec &hit, const PathRec &path, const PendRec &pend) {
    const DevScene &sc = c.sc;
    const D3 o = mk3(ray.ox, ray.oy, ray.oz), d = mk3(ray.dx, ray.dy, ray.dz);
    Isect v;
    if (path.flags == PEND_PRIMARY) {
        if (hit.prim < 0) { c.rad = miss_rad<ENV>(sc, d); return; }  // :8
        fill_isect<ENV>(sc, o, d, hit.prim, hit.t, hit.u, hit.v, v);
        if (v.light != -1 && sc.lights[v.light].kind == TAKE_LIGHT_AREA)  // :14-18
            c.rad = add(c.rad, mulv(c.thr, light_intensity(sc.lights[v.light])));
    } else {
        const D3 FG = mk3(pend.fg[0], pend.fg[1], pend.fg[2]);
        const double bpdf = pend.bpdf;
        const bool spec = (path.flags & PEND_SPECULAR) != 0;
        if (hit.prim < 0) {  // :82-87
            if (ENV && sc.env_light) {  // EXTENSION: the miss found the sampled environment -> MIS weight as for an emitter hit (:99)
                const double lpdf = env_pdf(sc, d) / sc.pick_count;
                D3 Ce = mul(mulv(FG, env_radiance(sc, d)), spec ? (1 / bpdf) : (bpdf / (lpdf * lpdf + bpdf * bpdf)));
                c.rad = add(c.rad, mulv(c.thr, Ce));
                return;
            }
            c.thr = mulv(c.thr, divs(FG, bpdf));
            c.rad = add(c.rad, mulv(c.thr, miss_rad<ENV>(sc, d)));
            return;
        }
        fill_isect<ENV>(sc, o, d, hit.prim, hit.t, hit.u, hit.v, v);
        if (v.light != -1) {  // :88-101
            double lpdf;
            if (!hit_light_pdf(sc, v, o, lpdf)) return;
            const TakeLightDesc &l = sc.lights[v.light];
            if (l.kind == TAKE_LIGHT_AREA) {
                D3 C2 = mul(mulv(FG, light_intensity(l)), spec ? (1 / bpdf) : (bpdf / (lpdf * lpdf + bpdf * bpdf)));
                c.rad = add(c.rad, mulv(c.thr, C2));
            }
        }
        c.thr = mulv(c.thr, divs(FG, bpdf));  // :107
    }
    if (c.depth > c.w.max_depth) return;  // loop bound :20
    if (!russian_roulette<RR>(c)) return;
    c.shaded += 1;
    const D3 dir_in = neg(d);
    const TakeMaterialDesc &m = sc.materials[v.material];
    const bool spec = is_specular(m.type);
    c.org = v.pos;
    if (sc.pick_count > 0 && !spec) {  // :30-59
        const int light_id = (int)floor(c.rng.next() * sc.pick_count);
        if (ENV && sc.env_light && light_id == sc.num_lights) {  // EXTENSION: the environment as a light
            const double u1 = c.rng.next();
            const double u2 = c.rng.next();
            D3 light_dir;
            double pdf_w;
            env_sample_dir(sc, u1, u2, light_dir, pdf_w);
            const double lpdf = pdf_w / sc.pick_count;
            if (lpdf <= 0) return;
            const double bpdf = bsdf_pdf(m, dir_in, light_dir, v);
            if (bpdf > 0 && !isinf(lpdf)) {
                D3 FG = bsdf_eval(sc, m, dir_in, light_dir, 0.0, v);
                D3 C1 = divs(mul(mulv(FG, env_radiance(sc, light_dir)), lpdf), lpdf * lpdf + bpdf * bpdf);
                c.emit_shadow = true;
                c.sh_dir = light_dir;
                c.sh_tmax = INFINITY;
                c.sh_contrib = mulv(c.thr, C1);
            }
        } else if (sc.lights[light_id].kind == TAKE_LIGHT_AREA) {
            const TakeLightDesc &l = sc.lights[light_id];
            D3 light_dir;
            double dist, lpdf;
            if (!nee_sample(sc, l, light_id, v, c.rng, light_dir, dist, lpdf)) return;
            const double bpdf = bsdf_pdf(m, dir_in, light_dir, v);
            if (bpdf > 0 && !isinf(lpdf)) {
                D3 FG = bsdf_eval(sc, m, dir_in, light_dir, 0.0, v);
                D3 C1 = divs(mul(mulv(FG, light_intensity(l)), lpdf), lpdf * lpdf + bpdf * bpdf);
                c.emit_shadow = true;
                c.sh_dir = light_dir;
                c.sh_tmax = (1 - TAKE_EPS) * dist;
                c.sh_contrib = mulv(c.thr, C1);
            }
        }
    }
    D3 rec_dir;
    double rec_pdf;
    if (!sample_bsdf(m, dir_in, v, c.rng, rec_dir, rec_pdf)) return;  // :65-69
    c.pend_fg = bsdf_eval(sc, m, dir_in, rec_dir, rec_pdf, v);
    c.ext_dir = normalize(rec_dir);
    c.pend_pdf = rec_pdf;
    if (rec_pdf <= 0) return;  // :75-78
    c.pend_flags = PEND_BSDF | (spec ? PEND_SPECULAR : 0);
    c.emit_extend = true;
    c.depth += 1;
}

// No MIS: src/integrator/path_tracing.h:114-157
template <bool ENV, bool RR>
__device__ __forceinline__ void shade_raw(ShadeCtx &c, const RayRec &ray, const HitRec &hit, const PathRec &path, const PendRec &) {
    const DevScene &sc = c.sc;
    const D3 o = mk3(ray.ox, ray.oy, ray.oz), d = mk3(ray.dx, ray.dy, ray.dz);
    if (hit.prim < 0) {
        if (path.flags == PEND_PRIMARY) c.rad = miss_rad<ENV>(sc, d);  // :117
        else c.rad = add(c.rad, mulv(c.thr, miss_rad<ENV>(sc, d)));     // :148-152 (throughput already updated, :145)
        return;
    }
    Isect v;
    fill_isect<ENV>(sc, o, d, hit.prim, hit.t, hit.u, hit.v, v);
    if (c.depth > c.w.max_depth) return;
    if (v.light != -1) {  // :123-129
        if (sc.lights[v.light].kind == TAKE_LIGHT_AREA) c.rad = add(c.rad, mulv(c.thr, light_intensity(sc.lights[v.light])));
        return;
    }
    if (!russian_roulette<RR>(c)) return;
    c.shaded += 1;
    const D3 dir_in = neg(d);
    const TakeMaterialDesc &m = sc.materials[v.material];
    D3 rec_dir;
    double pdf;
    if (!sample_bsdf(m, dir_in, v, c.rng, rec_dir, pdf)) return;
    D3 FG = bsdf_eval(sc, m, dir_in, rec_dir, pdf, v);
    c.ext_dir = normalize(rec_dir);
    if (pdf <= 0) return;
    c.thr = mulv(c.thr, divs(FG, pdf));
    c.org = v.pos;
    c.pend_fg = FG;
    c.pend_pdf = pdf;
    c.pend_flags = PEND_BSDF;
    c.emit_extend = true;
    c.depth += 1;
}

// light pdf with the light picked by power: get_light_pdf * d^2 * pmf / cos (path_tracing.h:309, :366)
__device__ __forceinline__ bool nee_sample_power(const DevScene &sc, const TakeLightDesc &l, int light_id, const Isect &v, Rng &rng,
                                                 D3 &light_dir, double &dist, double &lpdf) {
    D3 lp, ln;
    sample_on_light(sc, light_id, l.prim_id, v.pos, rng, lp, ln);
    dist = length(sub(lp, v.pos));
    light_dir = normalize(sub(lp, v.pos));
    lpdf = light_pdf_area(sc, light_id, lp, v.pos) * (dist * dist) * sc.light_pmf[light_id] / (fmax(dot(neg(ln), light_dir), 0.0));
    return !(lpdf <= 0);
}
__device__ __forceinline__ bool hit_light_pdf_power(const DevScene &sc, const Isect &nv, D3 prev_pos, double &lpdf) {
    double d = length(sub(nv.pos, prev_pos));
    D3 light_dir = normalize(sub(nv.pos, prev_pos));
    lpdf = light_pdf_area(sc, nv.light, nv.pos, prev_pos) * (d * d) * sc.light_pmf[nv.light] / fmax(dot(neg(nv.gn), light_dir), 0.0);
    return !(lpdf <= 0);
}

// One-sample MIS: src/integrator/path_tracing.h:161-271.  POWER = true is its sibling that picks the light in proportion to its
// power, path_tracing_one_sample_MIS_power (:274-380): same structure; the light comes from the power CDF (light.cpp:9-17),
// the pmf stands where 1 / N stood, and the light-aimed ray is checked on arrival -- a miss adds the background, a
// non-emissive hit ends the path, and only then is the throughput updated (:326-335).
template <bool ENV, bool POWER, bool RR>
__device__ __forceinline__ void shade_one_sample(ShadeCtx &c, const RayRec &ray, const HitRec &hit, const PathRec &path, const PendRec &pend) {
    const DevScene &sc = c.sc;
    const D3 o = mk3(ray.ox, ray.oy, ray.oz), d = mk3(ray.dx, ray.dy, ray.dz);
    const int nl = (ENV && !POWER) ? sc.pick_count : sc.num_lights;
    Isect v;
    if (path.flags == PEND_PRIMARY) {
        if (hit.prim < 0) { c.rad = miss_rad<ENV>(sc, d); return; }
        fill_isect<ENV>(sc, o, d, hit.prim, hit.t, hit.u, hit.v, v);
    } else if (path.flags & PEND_LIGHT) {
        if (POWER) {
            if (hit.prim < 0) { c.rad = add(c.rad, mulv(c.thr, miss_rad<ENV>(sc, d))); return; }   // :326-330
            fill_isect<ENV>(sc, o, d, hit.prim, hit.t, hit.u, hit.v, v);
            if (v.light == -1) return;                                                              // :332-334
            c.thr = mulv(c.thr, divs(mk3(pend.fg[0], pend.fg[1], pend.fg[2]), pend.bpdf));          // :335 (bpdf slot: 0.5 lpdf + 0.5 bpdf)
        } else {
            if (ENV && hit.prim < 0 && (path.flags & PEND_ENV)) {  // EXTENSION: reached the environment we aimed at
                c.rad = add(c.rad, mulv(c.thr, env_radiance(sc, d)));
                return;
            }
            if (hit.prim < 0) {
                // the reference dereferences an empty optional here (:220, undefined behaviour): terminate and count
                atomicAdd(&c.w.totals->miss_after_light_sample, 1ULL);
                return;
            }
            fill_isect<ENV>(sc, o, d, hit.prim, hit.t, hit.u, hit.v, v);
        }
    } else {
        const D3 FG = mk3(pend.fg[0], pend.fg[1], pend.fg[2]);
        const bool spec = (path.flags & PEND_SPECULAR) != 0;
        double pdf = (nl == 0 || spec) ? pend.bpdf : 0.5 * pend.bpdf;  // :245
        if (hit.prim < 0) {  // :247-252
            if (ENV && !POWER && sc.env_light && !spec) pdf += 0.5 * (env_pdf(sc, d) / nl);  // EXTENSION: mixture pdf, as :255-265 for emitters
            c.thr = mulv(c.thr, divs(FG, pdf));
            c.rad = add(c.rad, mulv(c.thr, miss_rad<ENV>(sc, d)));
            return;
        }
        fill_isect<ENV>(sc, o, d, hit.prim, hit.t, hit.u, hit.v, v);
        if (!spec && v.light != -1) {  // :253-265
            double lpdf;
            if (!(POWER ? hit_light_pdf_power(sc, v, o, lpdf) : hit_light_pdf(sc, v, o, lpdf))) return;
            pdf += 0.5 * lpdf;
        }
        c.thr = mulv(c.thr, divs(FG, pdf));
    }
    const D3 dir_in = neg(d);
    for (; c.depth <= c.w.max_depth; c.depth += 1) {
        if (v.light != -1 && sc.lights[v.light].kind == TAKE_LIGHT_AREA) {  // :170-177
            c.rad = add(c.rad, mulv(c.thr, light_intensity(sc.lights[v.light])));
            return;
        }
        if (!russian_roulette<RR>(c)) return;
        c.shaded += 1;
        const TakeMaterialDesc &m = sc.materials[v.material];
        const bool spec = is_specular(m.type);
        c.org = v.pos;
        if (nl > 0 && !spec && c.rng.next() <= 0.5) {  // :187
            int light_id;
            if (POWER) {   // sample_light_power, light.cpp:9-17
                const double u = c.rng.next();
                light_id = min(max(upper_bound_idx(sc.light_cdf, nl + 1, u) - 1, 0), nl - 1);
            } else {
                light_id = (int)floor(c.rng.next() * nl);
            }
            if (ENV && !POWER && sc.env_light && light_id == sc.num_lights) {  // EXTENSION: the environment as a light
                const double u1 = c.rng.next();
                const double u2 = c.rng.next();
                D3 light_dir;
                double pdf_w;
                env_sample_dir(sc, u1, u2, light_dir, pdf_w);
                const double lpdf = pdf_w / nl;
                if (lpdf <= 0) return;
                const double bpdf = bsdf_pdf(m, dir_in, light_dir, v);
                if (bpdf <= 0) return;
                D3 FG = bsdf_eval(sc, m, dir_in, light_dir, 0.0, v);
                c.thr = mulv(c.thr, divs(FG, 0.5 * lpdf + 0.5 * bpdf));
                c.ext_dir = light_dir;
                c.pend_fg = FG;
                c.pend_pdf = bpdf;
                c.pend_flags = PEND_LIGHT | PEND_ENV;
                c.emit_extend = true;
                c.depth += 1;
                return;
            }
            const TakeLightDesc &l = sc.lights[light_id];
            if (l.kind != TAKE_LIGHT_AREA) continue;  // a point light: the iteration does nothing (:190)
            D3 light_dir;
            double dist, lpdf;
            if (!(POWER ? nee_sample_power(sc, l, light_id, v, c.rng, light_dir, dist, lpdf) : nee_sample(sc, l, light_id, v, c.rng, light_dir, dist, lpdf)))
                return;
            const double bpdf = bsdf_pdf(m, dir_in, light_dir, v);
            if (bpdf <= 0) return;
            D3 FG = bsdf_eval(sc, m, dir_in, light_dir, 0.0, v);
            if (POWER) {
                c.pend_pdf = 0.5 * lpdf + 0.5 * bpdf;   // applied when the ray has arrived at an emitter (:335)
            } else {
                c.thr = mulv(c.thr, divs(FG, 0.5 * lpdf + 0.5 * bpdf));  // :225
                c.pend_pdf = bpdf;
            }
            c.ext_dir = light_dir;
            c.pend_fg = FG;
            c.pend_flags = PEND_LIGHT;
            c.emit_extend = true;
            c.depth += 1;
            return;
        }
        D3 rec_dir;
        double rec_pdf;
        if (!sample_bsdf(m, dir_in, v, c.rng, rec_dir, rec_pdf)) return;
        c.pend_fg = bsdf_eval(sc, m, dir_in, rec_dir, rec_pdf, v);
        c.ext_dir = normalize(rec_dir);
        c.pend_pdf = rec_pdf;
        if (rec_pdf <= 0) return;
        c.pend_flags = PEND_BSDF | (spec ? PEND_SPECULAR : 0);
        c.emit_extend = true;
        c.depth += 1;
        return;
    }
}

#ifndef TAKE_SHADE_PREFETCH
#define TAKE_SHADE_PREFETCH 2   // 0: off, 1: next iteration's hit record -> L2, 2: -> L1 (measured: -4 % on one-sample shade).
                                // (A second stage -- reading the next iteration's (prim, slot) early and prefetching the ray /
                                // path / pending-sample / shading records they point to into L2 -- was measured: shade +4 ... +9 %
                                // SLOWER on all four scenes; the extra requests compete with the demand loads.)
#endif
#ifndef TAKE_SHADE_EARLY
#define TAKE_SHADE_EARLY 0   // 0: records fetched when needed, 1: ray+path right after the slot, 2: + the pending BSDF sample
#endif
#ifndef TAKE_SHADE_MIN_BLOCKS
#define TAKE_SHADE_MIN_BLOCKS 1
#endif
template <int INTEGRATOR, bool ENV, bool RR = false>
__global__ void __launch_bounds__(128, TAKE_SHADE_MIN_BLOCKS) k_shade(DevScene sc, Wave w, int pass) {
    PassCounters &pc = w.pass[pass];
    const uint32_t n = shade_count(w, pass);
    const bool primary = pass == 0 && w.fused_primary;
    const int32_t *queue = w.q_extend[pass & 1];  // only read when the sort is off
    int32_t *q_next = w.q_extend[(pass + 1) & 1];
    // grid-stride with whole warps, so that the queue pushes below always see converged warps
    const uint32_t n_round = (n + 31u) & ~31u;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += gridDim.x * blockDim.x) {
        const bool valid = i < n;
        bool emit_extend = false, emit_shadow = false;
        int slot = -1, shaded = 0;
#if TAKE_SHADE_PREFETCH
        // the record this thread needs in its next iteration: requested now, so that the ~1 us DRAM round trip at the head
        // of the dependent chain (hit -> slot -> ray/path -> vertex data) overlaps this iteration's shading
        if (w.sort_enabled && i + gridDim.x * blockDim.x < n) {
#if TAKE_SHADE_PREFETCH == 1
            asm volatile("prefetch.global.L2 [%0];" ::"l"(w.hit_sorted + i + gridDim.x * blockDim.x));
#else
            asm volatile("prefetch.global.L1 [%0];" ::"l"(w.hit_sorted + i + gridDim.x * blockDim.x));
#endif
        }
#endif
        if (valid) {
            const HitRec hit = w.sort_enabled ? w.hit_sorted[i] : w.hit[primary ? (int)i : queue[i]];
            slot = w.sort_enabled ? (int)hit.keyrank : (primary ? (int)i : queue[i]);
            // Everything that depends only on the slot is requested before the hit is looked at, so that the gathers
            // overlap instead of forming a chain (slot -> hit -> ray -> path -> pend): the kernel is latency-bound.
            RayRec ray;
            PathRec path;
            PendRec pend;
#if TAKE_SHADE_EARLY >= 1
            if (!primary) {
                ray = w.ray[slot];
                path = w.path[slot];
#if TAKE_SHADE_EARLY >= 2
                if (INTEGRATOR != TAKE_INTEGRATOR_RAW) pend = w.pend[slot];
#endif
            }
#endif
            if (primary && !ENV && hit.prim < 0) {
                // a camera ray that left the scene: the sample is the background colour (path_tracing.h:8); no need to
                // recompute the ray (the environment-map build needs its direction and takes the general path)
                PathRec p;
                p.thr[0] = p.thr[1] = p.thr[2] = 1.0;
                p.rad[0] = sc.background.x; p.rad[1] = sc.background.y; p.rad[2] = sc.background.z;
                p.k = 2; p.depth = 0; p.flags = 0; p.pad = 0;
                w.path[slot] = p;
            } else {
                if (primary) {  // nothing was stored for the camera ray: recompute it (2 draws) and start the path
                    D3 o, d;
                    primary_ray(sc, w, slot, o, d);
                    ray.ox = o.x; ray.oy = o.y; ray.oz = o.z; ray.dx = d.x; ray.dy = d.y; ray.dz = d.z;
                    ray.tmax = INFINITY; ray.aux0 = ray.aux1 = 0;
                    path.thr[0] = path.thr[1] = path.thr[2] = 1.0;
                    path.rad[0] = path.rad[1] = path.rad[2] = 0.0;
                    path.k = 2; path.depth = 0; path.flags = PEND_PRIMARY; path.pad = 0;
                }
#if TAKE_SHADE_EARLY < 1
                else { ray = w.ray[slot]; path = w.path[slot]; }
#endif
#if TAKE_SHADE_EARLY < 2
                if (!primary && INTEGRATOR != TAKE_INTEGRATOR_RAW &&
                    ((path.flags & PEND_BSDF) || (INTEGRATOR == TAKE_INTEGRATOR_ONE_SAMPLE_MIS_POWER && (path.flags & PEND_LIGHT))))
                    pend = w.pend[slot];
#endif
                uint32_t pixel;
                uint64_t sample;
                slot_identity(w, slot, pixel, sample);
                ShadeCtx c = {sc, w, slot, pass};
                c.thr = mk3(path.thr[0], path.thr[1], path.thr[2]);
                c.rad = mk3(path.rad[0], path.rad[1], path.rad[2]);
                c.rng.seed = w.seed; c.rng.sample = sample; c.rng.pixel = pixel; c.rng.k = path.k;
                c.emit_extend = c.emit_shadow = false;
                c.depth = path.depth;
                c.pend_flags = 0;
                c.shaded = 0;
                c.org = mk3(ray.ox, ray.oy, ray.oz);
                if (INTEGRATOR == TAKE_INTEGRATOR_MIS) shade_mis<ENV, RR>(c, ray, hit, path, pend);
                else if (INTEGRATOR == TAKE_INTEGRATOR_RAW) shade_raw<ENV, RR>(c, ray, hit, path, pend);
                else if (INTEGRATOR == TAKE_INTEGRATOR_ONE_SAMPLE_MIS) shade_one_sample<ENV, false, RR>(c, ray, hit, path, pend);
                else shade_one_sample<ENV, true, RR>(c, ray, hit, path, pend);
                emit_extend = c.emit_extend;
                emit_shadow = c.emit_shadow;
                shaded = c.shaded;
                PathRec p;
                p.thr[0] = c.thr.x; p.thr[1] = c.thr.y; p.thr[2] = c.thr.z;
                p.rad[0] = c.rad.x; p.rad[1] = c.rad.y; p.rad[2] = c.rad.z;
                p.k = c.rng.k;
                p.depth = c.depth;
                p.flags = c.pend_flags;
                p.pad = 0;
                w.path[slot] = p;
                if (emit_extend || emit_shadow) {
                    RayRec r;
                    r.ox = c.org.x; r.oy = c.org.y; r.oz = c.org.z;
                    r.dx = c.ext_dir.x; r.dy = c.ext_dir.y; r.dz = c.ext_dir.z;
                    r.tmax = INFINITY;
                    // the coin of the vertex this ray will reach is the stream's next draw (see TAKE_KEY_BITS)
                    r.aux0 = ((INTEGRATOR == TAKE_INTEGRATOR_ONE_SAMPLE_MIS || INTEGRATOR == TAKE_INTEGRATOR_ONE_SAMPLE_MIS_POWER) && w.sort_branch &&
                              emit_extend) ? peek_branch(c.rng, RR && w.rr_start > 0 && c.depth >= w.rr_start) : 0;
                    r.aux1 = 0;
                    w.ray[slot] = r;
                }
                if (emit_extend) {
                    PendRec pe;
                    pe.fg[0] = c.pend_fg.x; pe.fg[1] = c.pend_fg.y; pe.fg[2] = c.pend_fg.z;
                    pe.bpdf = c.pend_pdf;
                    w.pend[slot] = pe;
                }
                if (emit_shadow) {
                    ShadowRec s;
                    s.dx = c.sh_dir.x; s.dy = c.sh_dir.y; s.dz = c.sh_dir.z;
                    s.tmax = c.sh_tmax;
                    s.cx = c.sh_contrib.x; s.cy = c.sh_contrib.y; s.cz = c.sh_contrib.z;
                    s.pad = 0;
                    w.shadow[slot] = s;
                }
                    }
}
        queue_push(emit_extend, q_next, &w.pass[pass + 1].n_extend, slot);
        queue_push(emit_shadow, w.q_shadow, &pc.n_shadow, slot);
        shaded = __reduce_add_sync(0xffffffffu, shaded);
        if ((threadIdx.x & 31) == 0 && shaded) atomicAdd(&w.totals->shaded, (unsigned long long)shaded);
    }
}

// ---- shadow-connect: any-hit query; unoccluded connections add throughput * C1 (path_tracing.h:53-60) ----------
template <bool COUNT, bool WIDE>
__global__ void __launch_bounds__(128, TAKE_BOUNCE_MIN_BLOCKS) k_shadow(DevScene sc, Wave w, int pass) {
    TAKE_DECLARE_STACK(st);
    PassCounters &pc = w.pass[pass];
    const uint32_t n = pc.n_shadow;
    const int lane = threadIdx.x & 31;
    TravCounters cnt = {0, 0};
    WorkCursor wc = {0, 0};
    for (;;) {
        uint32_t base;
        if (!wc.next(&pc.fetch_shadow, n, lane, base)) break;
        const uint32_t i = base + lane;
        if (i < n) {
            const int slot = w.q_shadow[i];
            const RayRec r = ld_stream(w.ray + slot);
            const ShadowRec s = ld_stream(w.shadow + slot);
            HitOut h;
            trace_any<true, COUNT, WIDE>(sc, mk3(r.ox, r.oy, r.oz), mk3(s.dx, s.dy, s.dz), TAKE_EPS, s.tmax, st, h, &cnt);
            if (h.prim < 0) {
                PathRec *p = w.path + slot;
                p->rad[0] += s.cx; p->rad[1] += s.cy; p->rad[2] += s.cz;
            }
        }
    }
    if (COUNT) {
        atomicAdd(&w.totals->shadow_box_tests, cnt.box);
        atomicAdd(&w.totals->shadow_tri_tests, cnt.tri);
    }
}

// ---- bounce / shadow passes with lane refill (trace_refill4): same results, rays handed to lanes as lanes free up -----
struct ExtendRefillIO {
    const Wave &w;
    PassCounters &pc;
    const int32_t *queue;
    const uint8_t *mtype;
    int slot;
    int32_t branch;
    static constexpr bool D_SPLIT = true;   // RayRec: ox oy oz dx | dy dz tmax aux
    __device__ __forceinline__ void load(uint32_t i, const DevScene &sc, LaneRay &r) {
        slot = queue[i];
        const RayRec *rr = w.ray + slot;
        branch = rr->aux0;
        lane_ray_setup<true>(r, sc, reinterpret_cast<const double2 *>(rr), reinterpret_cast<const double2 *>(rr) + 2, rr->tmax);
    }
    // the bounce-pass part of extend_finish, for the lanes in `mask`
    __device__ __forceinline__ void retire(unsigned mask, bool mine, const HitOut &h) {
        if (!mine) return;
        const int lane = threadIdx.x & 31;
        const uint32_t key = sort_key(h.prim, mtype, branch);
        const unsigned peers = __match_any_sync(mask, key);
        const int leader = __ffs(peers) - 1;
        uint32_t rbase = 0;
        if (lane == leader) rbase = atomicAdd(&pc.bins[key], (uint32_t)__popc(peers));
        rbase = __shfl_sync(peers, rbase, leader);
        HitRec hr;
        hr.prim = h.prim;
        hr.keyrank = (key << TAKE_RANK_BITS) | (rbase + __popc(peers & ((1u << lane) - 1u)));
        hr.t = h.t; hr.u = h.u; hr.v = h.v;
        st_stream(w.hit + slot, hr);
    }
};

#ifndef TAKE_REFILL_MIN_BLOCKS
#define TAKE_REFILL_MIN_BLOCKS TAKE_BOUNCE_MIN_BLOCKS
#endif
#ifndef TAKE_REFILL_EXTEND_MIN_BLOCKS
#define TAKE_REFILL_EXTEND_MIN_BLOCKS TAKE_REFILL_MIN_BLOCKS
#endif
template <bool COUNT, bool TIES = false>
__global__ void __launch_bounds__(128, TAKE_REFILL_EXTEND_MIN_BLOCKS) k_extend_refill(DevScene sc, Wave w, int pass) {
    TAKE_DECLARE_STACK(st);
    PassCounters &pc = w.pass[pass];
    ExtendRefillIO io = {w, pc, w.q_extend[pass & 1], sc.prim_mtype, -1, 0};
    TravCounters cnt = {0, 0};
    trace_refill4<false, COUNT, TIES>(sc, io, pc.n_extend, &pc.fetch_extend, st, &cnt);
    if (COUNT) {
        atomicAdd(&w.totals->box_tests, cnt.box);
        atomicAdd(&w.totals->tri_tests, cnt.tri);
    }
}

struct ShadowRefillIO {
    const Wave &w;
    int slot;
    static constexpr bool D_SPLIT = false;  // origin from RayRec, direction = the first three doubles of ShadowRec
    __device__ __forceinline__ void load(uint32_t i, const DevScene &sc, LaneRay &r) {
        slot = w.q_shadow[i];
        const RayRec *rr = w.ray + slot;
        const ShadowRec *sr = w.shadow + slot;
        lane_ray_setup<false>(r, sc, reinterpret_cast<const double2 *>(rr), reinterpret_cast<const double2 *>(sr), sr->tmax);
    }
    __device__ __forceinline__ void retire(unsigned, bool mine, const HitOut &h) {
        if (!mine || h.prim >= 0) return;
        const ShadowRec *sr = w.shadow + slot;
        PathRec *p = w.path + slot;
        p->rad[0] += sr->cx; p->rad[1] += sr->cy; p->rad[2] += sr->cz;
    }
};

template <bool COUNT>
__global__ void __launch_bounds__(128, TAKE_REFILL_MIN_BLOCKS) k_shadow_refill(DevScene sc, Wave w, int pass) {
    TAKE_DECLARE_STACK(st);
    PassCounters &pc = w.pass[pass];
    ShadowRefillIO io = {w, -1};
    TravCounters cnt = {0, 0};
    trace_refill4<true, COUNT, false>(sc, io, pc.n_shadow, &pc.fetch_shadow, st, &cnt);
    if (COUNT) {
        atomicAdd(&w.totals->shadow_box_tests, cnt.box);
        atomicAdd(&w.totals->shadow_tri_tests, cnt.tri);
    }
}

// ---- accumulate: per pixel, add the wave's samples in ascending sample order (src/render.cpp:67-78) ------------
__global__ void k_accumulate(Wave w, double *sum, double *sumsq, int n_passes) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p == 0) {
        unsigned long long ext = 0, sh = 0;
        for (int b = 0; b < n_passes; ++b) { ext += pass_count(w, b); sh += w.pass[b].n_shadow; }
        atomicAdd(&w.totals->extend_rays, ext);
        atomicAdd(&w.totals->shadow_rays, sh);
        atomicAdd(&w.totals->samples, (unsigned long long)w.n_slots);
    }
    if (p >= w.chunk_pixels) return;
    const size_t o = 3 * (size_t)pixel_of_index(w, (uint32_t)(w.chunk_base + p));
    double a0 = sum[o], a1 = sum[o + 1], a2 = sum[o + 2];
    double b0 = 0, b1 = 0, b2 = 0;
    if (sumsq) { b0 = sumsq[o]; b1 = sumsq[o + 1]; b2 = sumsq[o + 2]; }
    for (int s = 0; s < w.samples_in_wave; ++s) {
        const PathRec *pr = w.path + slot_of(w, (uint32_t)p, (uint32_t)s);
        const double r0 = pr->rad[0], r1 = pr->rad[1], r2 = pr->rad[2];
        a0 += r0; a1 += r1; a2 += r2;
        b0 += r0 * r0; b1 += r1 * r1; b2 += r2 * r2;
    }
    sum[o] = a0; sum[o + 1] = a1; sum[o + 2] = a2;
    if (sumsq) { sumsq[o] = b0; sumsq[o + 1] = b1; sumsq[o + 2] = b2; }
}

// explicit-list mode: radiance of each slot
__global__ void k_gather_radiance(Wave w, double *out, int n_passes) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s == 0) {
        unsigned long long ext = 0, sh = 0;
        for (int b = 0; b < n_passes; ++b) { ext += pass_count(w, b); sh += w.pass[b].n_shadow; }
        atomicAdd(&w.totals->extend_rays, ext);
        atomicAdd(&w.totals->shadow_rays, sh);
        atomicAdd(&w.totals->samples, (unsigned long long)w.n_slots);
    }
    if (s >= w.n_slots) return;
    out[3 * (size_t)s] = w.path[s].rad[0];
    out[3 * (size_t)s + 1] = w.path[s].rad[1];
    out[3 * (size_t)s + 2] = w.path[s].rad[2];
}

// ---- direct intersection entry points (take_gpu_intersect / take_gpu_occluded) -------------------------------
template <bool ANY_HIT, bool WIDE>
__global__ void __launch_bounds__(128, TAKE_V1_MIN_BLOCKS) k_intersect_fast(DevScene sc, const TakeRay *rays, int64_t n, TakeHit *hits, uint8_t *occ,
                                                        uint32_t *fetch) {
    TAKE_DECLARE_STACK(st);
    const int lane = threadIdx.x & 31;
    WorkCursor wc = {0, 0};
    const uint32_t n32 = (uint32_t)n;
    for (;;) {
        uint32_t base;
        if (!wc.next(fetch, n32, lane, base)) break;
        const int64_t i = (int64_t)base + lane;
        if (i < n) {
            const TakeRay r = rays[i];
            HitOut h;
            trace_any<ANY_HIT, false, WIDE>(sc, mk3(r.origin[0], r.origin[1], r.origin[2]), mk3(r.dir[0], r.dir[1], r.dir[2]),
                                            r.tmin, r.tmax, st, h, nullptr);
            if (ANY_HIT) {
                occ[i] = h.prim >= 0 ? 1 : 0;
            } else {
                TakeHit o;
                o.prim_id = h.prim; o.pad = 0; o.t = h.t; o.u = h.u; o.v = h.v;
                hits[i] = o;
            }
        }
    }
}


__global__ void k_intersect_exact(DevScene sc, const TakeRay *rays, int64_t n, TakeHit *hits) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const TakeRay r = rays[i];
    HitOut h;
    trace_exact(sc, mk3(r.origin[0], r.origin[1], r.origin[2]), mk3(r.dir[0], r.dir[1], r.dir[2]), r.tmin, r.tmax, h);
    TakeHit o;
    o.prim_id = h.prim; o.pad = 0; o.t = h.t; o.u = h.u; o.v = h.v;
    hits[i] = o;
}

}  // namespace take
