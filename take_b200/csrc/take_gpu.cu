// C-ABI of the B200 rendering core (include/take_gpu.h): host orchestration of the CUDA kernels.
// There is no CPU fallback anywhere in this file: every entry point needs a CUDA device.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <new>
#include <string>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/take_gpu.h"
#include "bvh_build.h"
#include "wavefront.cuh"
#include "exr_out.cuh"
#include "bvh_device.cuh"

using namespace take;

namespace {

thread_local std::string g_error;

int fail(int code, const std::string &msg) {
    g_error = msg;
    return code;
}

#define CU(expr)                                                                                             \
    do {                                                                                                     \
        cudaError_t e_ = (expr);                                                                             \
        if (e_ != cudaSuccess)                                                                               \
            return fail(e_ == cudaErrorMemoryAllocation ? TAKE_E_NOMEM : TAKE_E_CUDA,                        \
                        std::string(#expr) + ": " + cudaGetErrorString(e_));                                 \
    } while (0)

#ifndef TAKE_WARP_SAMPLES_DEFAULT
#define TAKE_WARP_SAMPLES_DEFAULT 32
#endif

int env_int(const char *name, int dflt) {
    const char *v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

// Host threads the builders may use: all cores, unless TAKE_HOST_THREADS says otherwise (one process per GPU: every rank of a
// node would otherwise start a full set of threads on the same cores).
int host_threads() {
    const int e = env_int("TAKE_HOST_THREADS", 0);
    return e > 0 ? e : (int)std::max(1u, std::thread::hardware_concurrency());
}

// Device memory comes from the device's stream-ordered pool with the release threshold lifted: what a scene or a builder
// frees stays mapped in the pool and the next allocation of the process -- the next scene, the next wave size -- gets it back
// in microseconds.  (Plain cudaMalloc / cudaFree cost ~0.1 s per GB in a process that has just released tens of GB of wave
// buffers: scene creation measured 4x slower in a process that renders several scenes than in a fresh one.)  Semantics are
// those of cudaMalloc / cudaFree: the pointer is usable on any stream at once, and a free waits for the device first.
// TAKE_MEMPOOL=0 goes back to cudaMalloc / cudaFree.
struct PoolStreams {
    std::mutex mu;
    cudaStream_t st[64] = {};
    bool off = false, init = false;
    cudaStream_t get(int dev) {
        std::lock_guard<std::mutex> g(mu);
        if (!init) { init = true; off = env_int("TAKE_MEMPOOL", 1) == 0; }
        if (off || dev < 0 || dev >= 64) return nullptr;
        if (!st[dev]) {
            cudaMemPool_t pool;
            if (cudaDeviceGetDefaultMemPool(&pool, dev) != cudaSuccess) { cudaGetLastError(); off = true; return nullptr; }
            uint64_t keep = UINT64_MAX;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            if (cudaStreamCreateWithFlags(&st[dev], cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); st[dev] = nullptr; off = true; }
        }
        return st[dev];
    }
};
PoolStreams g_pool_streams;

struct DeviceBuffer {
    void *p = nullptr;
    size_t bytes = 0;
    int dev = -1;
    cudaStream_t pool_stream = nullptr;   // non-null: p came from the pool
    void release() {
        if (!p) return;
        if (pool_stream) {
            int cur = -1;
            cudaGetDevice(&cur);
            if (cur != dev) cudaSetDevice(dev);
            cudaDeviceSynchronize();      // what cudaFree does implicitly: nobody is using the buffer any more
            cudaFreeAsync(p, pool_stream);
            if (cur != dev && cur >= 0) cudaSetDevice(cur);
        } else {
            cudaFree(p);
        }
        p = nullptr; bytes = 0; pool_stream = nullptr;
    }
    ~DeviceBuffer() { release(); }
    cudaError_t ensure(size_t n) {
        if (n <= bytes) return cudaSuccess;
        release();
        cudaGetDevice(&dev);
        cudaStream_t ps = g_pool_streams.get(dev);
        cudaError_t e;
        if (ps) {
            e = cudaMallocAsync(&p, n, ps);
            if (e == cudaSuccess) e = cudaStreamSynchronize(ps);   // usable on every stream from here on
            if (e == cudaSuccess) pool_stream = ps;
        } else {
            e = cudaMalloc(&p, n);
        }
        if (e == cudaSuccess) bytes = n; else p = nullptr;
        return e;
    }
    template <typename T> T *as() { return (T *)p; }
};

}  // namespace

// The reference-order tree of a device-built scene: built by a host thread from its own copy of the primitive boxes, shared
// by all replicas of a multi-GPU handle, joined by the first call that traces rays.
struct RefJob {
    std::thread th;
    std::once_flag joined;
    RefTree tree;
    std::vector<Aabb> boxes;
    double t0 = 0, ms = 0;
    std::atomic<bool> done{false};
    void wait() {
        std::call_once(joined, [this] { if (th.joinable()) th.join(); });
    }
    ~RefJob() { if (th.joinable()) th.join(); }
};

struct TakeScene {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    DevScene dev{};
    int width = 0, height = 0;
    // scene storage
    DeviceBuffer env_rgb, env_marg, env_cond;
    DeviceBuffer nodes, wide_nodes, tris, ref_nodes, positions, normals, uvs, indices, prim_material, prim_light, dfs_rank, prim_flags,
        prim_mtype, spheres, materials, lights, textures, shade_recs, light_recs, light_pmf, light_cdf;
    std::vector<DeviceBuffer *> tex_data;
    DeviceBuffer exr_packed;
    // wave storage
    struct WaveBuffers {
        DeviceBuffer ray, hit, hit_sorted, path, pend, shadow, q0, q1, q_shadow, pass;
    } wb[2];  // two waves in flight: set i is driven by streams[i]
    DeviceBuffer totals, scratch_a, scratch_b, scratch_c, fetch;
    cudaStream_t stream2 = nullptr;       // second wave stream (stream is the first and the API's stream)
    // take_gpu_render_async: two result slots, filled by a copy stream behind the render streams
    struct AsyncSlot {
        DeviceBuffer sum, sq, totals;
        cudaEvent_t e0 = nullptr, e1 = nullptr, done = nullptr;
        int64_t launches = 0, waves = 0, ticket = -1;
        bool busy = false;
        take::Totals *h_totals = nullptr;
    } async_slot[2];
    cudaStream_t copy_stream = nullptr;
    int64_t next_ticket = 0;
    cudaEvent_t ev_acc[2] = {nullptr, nullptr}, ev_begin = nullptr;
    int64_t wave_capacity = 0;
    int wave_sets = 0, wave_passes = 0;
    int blocks_extend = 0, blocks_shadow = 0, blocks_isect = 0, blocks_occl = 0, blocks_extend_primary = 0;
    int blocks_extend_refill = 0, blocks_shadow_refill = 0;
    int refill_extend = 0, refill_shadow = 0;   // bounce / shadow passes on the lane-refill kernels (TAKE_REFILL, see scene creation)
    bool wide = true;   // 4-wide nodes (TAKE_BVH_WIDTH=4, default) or binary nodes (TAKE_BVH_WIDTH=2)
    // Device-built scenes: the reference-order tree (tie-break ranks, exact mode) is built by a host thread while the device
    // builds the fast tree and while the caller goes on; finish_reference_tree() joins it before the first query.
    std::shared_ptr<RefJob> ref_job;
    bool ref_pending = false;
    DeviceBuffer tie_count, snap_sum, snap_sq;   // rank-decided ties of a render; output snapshots of a provisional render
    int64_t provisional_renders = 0, provisional_reruns = 0;
    DeviceBuffer leaf_prims;
    bool device_built = false;
    bool power_ok = false;   // the scene has emitters with non-zero total power (the power-sampling integrator needs them)
    double create_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // validate, upload, prim boxes (host), device build, records, total, -, -
    // diagnostics
    double build_ms_ref = 0, build_ms_fast = 0;
    int fast_depth = 0;
    int wide_depth = 0;     // levels of the 4-wide tree (bounds the packet traversal's shared stack)
    double sah_cost = 0;
    int64_t num_fast_nodes = 0;
    ~TakeScene() {
        if (ref_job) ref_job->wait();
        for (auto *b : tex_data) delete b;
        for (auto e : ev_acc) if (e) cudaEventDestroy(e);
        if (ev_begin) cudaEventDestroy(ev_begin);
        for (auto &a : async_slot) {
            if (a.e0) cudaEventDestroy(a.e0);
            if (a.e1) cudaEventDestroy(a.e1);
            if (a.done) cudaEventDestroy(a.done);
            if (a.h_totals) cudaFreeHost(a.h_totals);
        }
        if (copy_stream) cudaStreamDestroy(copy_stream);
        if (stream2) cudaStreamDestroy(stream2);
        if (stream) cudaStreamDestroy(stream);
    }
};

namespace {

template <typename T>
int upload(DeviceBuffer &buf, const T *src, size_t count, cudaStream_t s) {
    size_t bytes = std::max<size_t>(count * sizeof(T), 16);
    CU(buf.ensure(bytes));
    if (count) CU(cudaMemcpyAsync(buf.p, src, count * sizeof(T), cudaMemcpyHostToDevice, s));
    return TAKE_OK;
}

double now_ms() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

int validate(const TakeSceneDesc *d) {
    if (!d) return fail(TAKE_E_INVALID, "null scene description");
    if (d->camera.width <= 0 || d->camera.height <= 0) return fail(TAKE_E_INVALID, "camera resolution must be positive");
    if (d->num_prims < 0 || d->num_vertices < 0 || d->num_spheres < 0 || d->num_materials < 0 || d->num_lights < 0 ||
        d->num_textures < 0)
        return fail(TAKE_E_INVALID, "negative count");
    if (d->num_prims >= (1 << 28)) return fail(TAKE_E_INVALID, "too many primitives (limit 2^28)");
    if (d->num_prims > 0 && (!d->indices || !d->prim_material || !d->prim_light || !d->prim_flags))
        return fail(TAKE_E_INVALID, "primitive arrays missing");
    if (d->num_vertices > 0 && (!d->positions || !d->normals || !d->uvs)) return fail(TAKE_E_INVALID, "vertex arrays missing");
    if (d->num_spheres > 0 && !d->spheres) return fail(TAKE_E_INVALID, "sphere array missing");
    if (d->num_materials > 0 && !d->materials) return fail(TAKE_E_INVALID, "material array missing");
    if (d->num_lights > 0 && !d->lights) return fail(TAKE_E_INVALID, "light array missing");
    if (d->num_textures > 0 && !d->textures) return fail(TAKE_E_INVALID, "texture array missing");
    // coordinates must be finite: an infinite or NaN extent would poison the conservative padding of every box test
    for (int64_t i = 0; i < 3 * d->num_vertices; ++i)
        if (!std::isfinite(d->positions[i])) return fail(TAKE_E_INVALID, "non-finite vertex position");
    for (int64_t i = 0; i < 4 * d->num_spheres; ++i)
        if (!std::isfinite(d->spheres[i])) return fail(TAKE_E_INVALID, "non-finite sphere");
    for (int64_t i = 0; i < d->num_prims; ++i) {
        const int32_t *id = d->indices + 3 * i;
        if (d->prim_flags[i] & TAKE_PRIM_SPHERE) {
            if (id[0] < 0 || id[0] >= d->num_spheres) return fail(TAKE_E_INVALID, "sphere index out of range");
        } else {
            for (int k = 0; k < 3; ++k)
                if (id[k] < 0 || id[k] >= d->num_vertices) return fail(TAKE_E_INVALID, "vertex index out of range");
        }
        if (d->prim_material[i] < 0 || d->prim_material[i] >= d->num_materials)
            return fail(TAKE_E_INVALID, "material id out of range");  // the reference would index out of bounds
        if (d->prim_light[i] < -1 || d->prim_light[i] >= d->num_lights) return fail(TAKE_E_INVALID, "light id out of range");
    }
    for (int i = 0; i < d->num_materials; ++i) {
        const TakeMaterialDesc &m = d->materials[i];
        if (m.type < 0 || m.type > TAKE_MAT_GGX) return fail(TAKE_E_INVALID, "unknown material type");
        if (m.tex_id >= d->num_textures) return fail(TAKE_E_INVALID, "texture id out of range");
    }
    for (int i = 0; i < d->num_textures; ++i)
        if (d->textures[i].width <= 0 || d->textures[i].height <= 0 || !d->textures[i].rgb) return fail(TAKE_E_INVALID, "bad texture");
    if (d->env_rgb && (d->env_width <= 0 || d->env_height <= 0)) return fail(TAKE_E_INVALID, "bad environment map size");
    for (int i = 0; i < d->num_lights; ++i) {
        const TakeLightDesc &l = d->lights[i];
        if (l.kind != TAKE_LIGHT_AREA && l.kind != TAKE_LIGHT_POINT) return fail(TAKE_E_INVALID, "unknown light kind");
        if (l.kind == TAKE_LIGHT_AREA && (l.prim_id < 0 || l.prim_id >= d->num_prims))
            return fail(TAKE_E_INVALID, "area light primitive out of range");
    }
    return TAKE_OK;
}

// f(a, b) over [0, n) in contiguous chunks on `threads` host threads (the per-primitive loops of host_build)
template <typename F>
void parallel_chunks(int64_t n, int threads, F f, int64_t min_parallel = 1 << 16) {
    const int parts = n >= min_parallel ? (int)std::min<int64_t>(std::max(1, threads), n) : 1;
    if (parts <= 1) { f((int64_t)0, n); return; }
    std::vector<std::thread> pool;
    for (int t = 1; t < parts; ++t) pool.emplace_back(f, n * t / parts, n * (t + 1) / parts);
    f((int64_t)0, n / parts);
    for (auto &t : pool) t.join();
}

struct HostBuild {
    RefTree ref;
    FastTree fast;
    std::vector<double> tris;  // 12 doubles per leaf slot
    double abs_max = 0, ms_ref = 0, ms_fast = 0;
    uint64_t geom_hash = 0;    // of the primitive data the structures were built from (geometry_hash)
    int64_t num_prims = 0;
};

// FNV-1a over the arrays the builders read: which primitives, where.  Chunks are hashed on all threads and the chunk
// hashes are hashed in order, so the value does not depend on the thread count.
uint64_t fnv1a(const void *data, size_t bytes, uint64_t h = 1469598103934665603ULL) {
    const unsigned char *p = (const unsigned char *)data;
    for (size_t i = 0; i < bytes; ++i) { h ^= p[i]; h *= 1099511628211ULL; }
    return h;
}
uint64_t hash_array(const void *data, size_t bytes, int threads) {
    const size_t CH = 1 << 22;
    const size_t nch = (bytes + CH - 1) / CH;
    std::vector<uint64_t> part(nch);
    parallel_chunks((int64_t)nch, threads, [&](int64_t a0, int64_t a1) {
        for (int64_t c = a0; c < a1; ++c)
            part[c] = fnv1a((const char *)data + (size_t)c * CH, std::min(CH, bytes - (size_t)c * CH));
    }, 8);
    return fnv1a(part.data(), part.size() * sizeof(uint64_t));
}
uint64_t geometry_hash(const TakeSceneDesc *d, int threads) {
    uint64_t h[5] = {0, 0, 0, 0, 0};
    if (d->num_prims > 0) {
        h[0] = hash_array(d->indices, (size_t)d->num_prims * 3 * sizeof(int32_t), threads);
        h[1] = hash_array(d->prim_flags, (size_t)d->num_prims, threads);
    }
    if (d->num_vertices > 0) h[2] = hash_array(d->positions, (size_t)d->num_vertices * 3 * sizeof(double), threads);
    if (d->num_spheres > 0) h[3] = hash_array(d->spheres, (size_t)d->num_spheres * 4 * sizeof(double), threads);
    h[4] = (uint64_t)d->num_prims;
    return fnv1a(h, sizeof(h));
}

// Everything scene_create does before touching CUDA: primitive boxes, both trees, leaf-ordered primitive records.
int host_build(const TakeSceneDesc *d, int threads, HostBuild &hb) {
    const int64_t n = d->num_prims;
    // primitive boxes exactly as build_bvh (src/scene.cpp:4-23)
    std::vector<Aabb> boxes((size_t)n);
    double abs_max = 0;
    std::mutex abs_mu;
    parallel_chunks(n, threads, [&](int64_t a0, int64_t a1) {
        double amax = 0;
        for (int64_t i = a0; i < a1; ++i) {
            Aabb &b = boxes[i];
            const int32_t *id = d->indices + 3 * i;
            if (d->prim_flags[i] & TAKE_PRIM_SPHERE) {
                const double *sp = d->spheres + 4 * (int64_t)id[0];
                for (int a = 0; a < 3; ++a) { b.lo[a] = sp[a] - sp[3]; b.hi[a] = sp[a] + sp[3]; }
            } else {
                const double *p0 = d->positions + 3 * (int64_t)id[0], *p1 = d->positions + 3 * (int64_t)id[1],
                             *p2 = d->positions + 3 * (int64_t)id[2];
                for (int a = 0; a < 3; ++a) {
                    b.lo[a] = std::min(std::min(p0[a], p1[a]), p2[a]);
                    b.hi[a] = std::max(std::max(p0[a], p1[a]), p2[a]);
                }
            }
            for (int a = 0; a < 3; ++a) amax = std::max(amax, std::max(fabs(b.lo[a]), fabs(b.hi[a])));
        }
        std::lock_guard<std::mutex> g(abs_mu);
        abs_max = std::max(abs_max, amax);
    });
    if (!std::isfinite(abs_max)) return fail(TAKE_E_INVALID, "non-finite scene extent");
    hb.abs_max = abs_max;
    hb.num_prims = n;
    hb.geom_hash = geometry_hash(d, threads);
    RefTree &ref = hb.ref;
    FastTree &fast = hb.fast;
    // the two trees are independent: build them side by side (the reference-order tree is bound by its serial
    // top-level std::sort calls, which must stay libstdc++'s to reproduce the reference's tie order)
    const int max_leaf = std::min(8, std::max(1, env_int("TAKE_BVH_MAX_LEAF", 4)));
    double t0 = now_ms(), t_ref = 0;
    std::thread ref_thread([&] { build_reference_tree(boxes.data(), n, threads, ref); t_ref = now_ms() - t0; });
    build_fast_tree(boxes.data(), n, max_leaf, 0.0f, threads, fast);
    hb.ms_fast = now_ms() - t0;
    ref_thread.join();
    hb.ms_ref = t_ref;
    if (fast.depth > TAKE_STACK_SMEM + TAKE_STACK_LOCAL)
        return fail(TAKE_E_INVALID, "acceleration tree too deep (" + std::to_string(fast.depth) + ")");

    // leaf-ordered FP64 primitive records: v0 | idbits | e1 | aux | e2 | kind
    std::vector<double> &tris = hb.tris;
    tris.resize((size_t)n * 12);
    parallel_chunks(n, threads, [&](int64_t a0, int64_t a1) {
        for (int64_t slot = a0; slot < a1; ++slot) {
            const int32_t prim = fast.leaf_prims[slot];
            double *T = tris.data() + 12 * slot;
            const int32_t *id = d->indices + 3 * (int64_t)prim;
            long long bits = ((long long)ref.dfs_rank[prim] << 32) | (long long)(uint32_t)prim;
            memcpy(&T[3], &bits, 8);
            if (d->prim_flags[prim] & TAKE_PRIM_SPHERE) {
                const double *sp = d->spheres + 4 * (int64_t)id[0];
                T[0] = sp[0]; T[1] = sp[1]; T[2] = sp[2];
                T[4] = T[5] = T[6] = 0; T[7] = sp[3];
                T[8] = T[9] = T[10] = 0; T[11] = 1.0;
            } else {
                const double *p0 = d->positions + 3 * (int64_t)id[0], *p1 = d->positions + 3 * (int64_t)id[1],
                             *p2 = d->positions + 3 * (int64_t)id[2];
                for (int a = 0; a < 3; ++a) {
                    T[a] = p0[a];
                    T[4 + a] = p1[a] - p0[a];  // e1 = v1 - v0, e2 = v2 - v0 (src/shape.cpp:53-54), computed once
                    T[8 + a] = p2[a] - p0[a];
                }
                T[7] = 0; T[11] = 0.0;
            }
        }
    });
    return TAKE_OK;
}

// `passes`: extend passes of a wave (max_depth + 2); every pass owns a PassCounters block, and the shade kernel of the last
// pass still addresses the block after it.
int ensure_wave(TakeScene *s, int64_t capacity, int sets, int passes) {
    if (capacity <= s->wave_capacity && sets <= s->wave_sets && passes <= s->wave_passes) return TAKE_OK;
    capacity = std::max(capacity, s->wave_capacity);
    sets = std::max(sets, s->wave_sets);
    passes = std::max(passes, s->wave_passes);
    const double t_alloc = now_ms();
    for (int i = 0; i < sets; ++i) {
        TakeScene::WaveBuffers &b = s->wb[i];
        CU(b.ray.ensure(capacity * sizeof(RayRec)));
        CU(b.hit.ensure(capacity * sizeof(HitRec)));
        CU(b.hit_sorted.ensure(capacity * sizeof(HitRec)));
        CU(b.path.ensure(capacity * sizeof(PathRec)));
        CU(b.pend.ensure(capacity * sizeof(PendRec)));
        CU(b.shadow.ensure(capacity * sizeof(ShadowRec)));
        CU(b.q0.ensure(capacity * 4));
        CU(b.q1.ensure(capacity * 4));
        CU(b.q_shadow.ensure(capacity * 4));
        CU(b.pass.ensure(sizeof(PassCounters) * (size_t)(passes + 1)));
    }
    CU(s->totals.ensure(sizeof(Totals)));
    s->wave_capacity = capacity;
    s->wave_sets = sets;
    s->wave_passes = passes;
    if (env_int("TAKE_TIMING", 0))
        fprintf(stderr, "[take_gpu] wave buffers: %d set(s) x %lld slots (%.2f GB) allocated in %.1f ms\n", sets, (long long)capacity,
                sets * capacity * 332.0 / 1e9, now_ms() - t_alloc);
    return TAKE_OK;
}

struct StageTimer {
    bool on = false;
    cudaStream_t stream;
    std::vector<cudaEvent_t> ev;
    std::vector<int> stage, pass_of;
    double ms[6] = {0, 0, 0, 0, 0, 0};
    enum { PASS_COLS = 16 };  // per-pass breakdown of the first passes (development aid); later passes fold into the last column
    double ms_pass[6][PASS_COLS] = {};
    int cur_pass = 0;
    void begin(int st) {
        if (!on) return;
        pass_of.push_back(cur_pass);
        cudaEvent_t a;
        cudaEventCreate(&a);
        cudaEventRecord(a, stream);
        ev.push_back(a);
        stage.push_back(st);
    }
    void end() {
        if (!on) return;
        cudaEvent_t b;
        cudaEventCreate(&b);
        cudaEventRecord(b, stream);
        ev.push_back(b);
    }
    void collect() {
        if (!on) return;
        cudaStreamSynchronize(stream);
        for (size_t i = 0; i < stage.size(); ++i) {
            float t = 0;
            cudaEventElapsedTime(&t, ev[2 * i], ev[2 * i + 1]);
            ms[stage[i]] += t;
            ms_pass[stage[i]][std::min(pass_of[i], (int)PASS_COLS - 1)] += t;
        }
        if (env_int("TAKE_PASS_TIMES", 0)) {  // development aid: per-pass stage times on stderr
            static const char *names[6] = {"generate", "extend", "shade", "shadow", "sort", "other"};
            for (int st = 1; st <= 3; ++st) {
                fprintf(stderr, "[take_gpu] %-8s per pass (ms):", names[st]);
                for (int b = 0; b < 12; ++b) fprintf(stderr, " %.3f", ms_pass[st][b]);
                fprintf(stderr, "\n");
            }
        }
        for (auto e : ev) cudaEventDestroy(e);
        ev.clear();
        stage.clear();
        pass_of.clear();
    }
};
enum { ST_GENERATE = 0, ST_EXTEND, ST_SHADE, ST_SHADOW, ST_SORT, ST_OTHER };

// `acc_after`: event the accumulation must wait for (the previous wave's accumulation: keeps the per-pixel summation
// order and avoids concurrent read-modify-write of the image); `acc_done`: recorded after this wave's accumulation.
int launch_wave(TakeScene *s, Wave &w, const TakeRenderOpts *o, double *d_sum, double *d_sumsq, double *d_list_out,
                StageTimer &tm, bool count, int64_t &launches, cudaStream_t st, cudaEvent_t acc_after = nullptr,
                cudaEvent_t acc_done = nullptr) {
    tm.stream = st;
    const int n_passes = o->max_depth + 2;
    CU(cudaMemsetAsync(w.pass, 0, sizeof(PassCounters) * (size_t)(n_passes + 1), st));
#if TAKE_EXPERIMENTAL
    if (!w.fused_primary) {
        tm.begin(ST_GENERATE);
        k_generate<<<(w.n_slots + 255) / 256, 256, 0, st>>>(s->dev, w);
        tm.end();
        launches++;
    }
#endif
    const int shade_blocks = std::max(1, std::min((w.n_slots + 127) / 128, s->sm_count * 64));
    const int scatter_blocks = std::max(1, std::min((w.n_slots + 255) / 256, s->sm_count * 16));
    for (int b = 0; b < n_passes; ++b) {
        tm.cur_pass = b;
        tm.begin(ST_EXTEND);
        if (b == 0 && w.packet) {
            if (w.count_ties) k_extend_primary<false, true><<<s->blocks_extend_primary, 128, 0, st>>>(s->dev, w);
            else if (count) k_extend_primary<true><<<s->blocks_extend_primary, 128, 0, st>>>(s->dev, w);
            else k_extend_primary<false><<<s->blocks_extend_primary, 128, 0, st>>>(s->dev, w);
        } else if (b >= 1 && s->refill_extend && s->wide) {
            // bounce rays have very different lengths: lanes are refilled as their rays finish (trace_refill4)
            if (w.count_ties) k_extend_refill<false, true><<<s->blocks_extend_refill, 128, 0, st>>>(s->dev, w, b);
            else if (count) k_extend_refill<true><<<s->blocks_extend_refill, 128, 0, st>>>(s->dev, w, b);
            else k_extend_refill<false><<<s->blocks_extend_refill, 128, 0, st>>>(s->dev, w, b);
        } else if (w.count_ties && s->wide) {
            k_extend<false, true, true><<<s->blocks_extend, 128, 0, st>>>(s->dev, w, b);
        } else
#if TAKE_EXPERIMENTAL
        if (!s->wide) {
            if (count) k_extend<true, false><<<s->blocks_extend, 128, 0, st>>>(s->dev, w, b);
            else k_extend<false, false><<<s->blocks_extend, 128, 0, st>>>(s->dev, w, b);
        } else
#endif
        {
            if (count) k_extend<true, true><<<s->blocks_extend, 128, 0, st>>>(s->dev, w, b);
            else k_extend<false, true><<<s->blocks_extend, 128, 0, st>>>(s->dev, w, b);
        }
        tm.end();
        if (w.sort_enabled) {
            tm.begin(ST_SORT);
            // (pass 0 is traced one ray per thread or as packets: its hits already finish in slot order)
            if (w.ordered_sort && b >= 1) k_scatter_ordered<<<s->sm_count * 8, 256, 0, st>>>(w, b);
            else k_scatter<<<scatter_blocks, 256, 0, st>>>(w, b);
            tm.end();
            launches++;
        }
        tm.begin(ST_SHADE);
        const bool env = s->dev.env_rgb != nullptr;
#define TAKE_SHADE(I) (env ? k_shade<I, true><<<shade_blocks, 128, 0, st>>>(s->dev, w, b)                         \
                           : w.rr_start > 0 ? k_shade<I, false, true><<<shade_blocks, 128, 0, st>>>(s->dev, w, b) \
                                            : k_shade<I, false><<<shade_blocks, 128, 0, st>>>(s->dev, w, b))
        if (o->integrator == TAKE_INTEGRATOR_MIS) TAKE_SHADE(TAKE_INTEGRATOR_MIS);
        else if (o->integrator == TAKE_INTEGRATOR_RAW) TAKE_SHADE(TAKE_INTEGRATOR_RAW);
        else if (o->integrator == TAKE_INTEGRATOR_ONE_SAMPLE_MIS) TAKE_SHADE(TAKE_INTEGRATOR_ONE_SAMPLE_MIS);
        else TAKE_SHADE(TAKE_INTEGRATOR_ONE_SAMPLE_MIS_POWER);
#undef TAKE_SHADE
        tm.end();
        launches += 2;
        if (o->integrator == TAKE_INTEGRATOR_MIS && b <= o->max_depth) {
            tm.begin(ST_SHADOW);
            if (s->refill_shadow && s->wide) {
                if (count) k_shadow_refill<true><<<s->blocks_shadow_refill, 128, 0, st>>>(s->dev, w, b);
                else k_shadow_refill<false><<<s->blocks_shadow_refill, 128, 0, st>>>(s->dev, w, b);
            } else
#if TAKE_EXPERIMENTAL
            if (!s->wide) {
                if (count) k_shadow<true, false><<<s->blocks_shadow, 128, 0, st>>>(s->dev, w, b);
                else k_shadow<false, false><<<s->blocks_shadow, 128, 0, st>>>(s->dev, w, b);
            } else
#endif
            {
                if (count) k_shadow<true, true><<<s->blocks_shadow, 128, 0, st>>>(s->dev, w, b);
                else k_shadow<false, true><<<s->blocks_shadow, 128, 0, st>>>(s->dev, w, b);
            }
            tm.end();
            launches++;
        }
    }
    if (acc_after) CU(cudaStreamWaitEvent(st, acc_after, 0));
    tm.begin(ST_OTHER);
    if (d_list_out) k_gather_radiance<<<(w.n_slots + 255) / 256, 256, 0, st>>>(w, d_list_out, n_passes);
    else k_accumulate<<<(w.chunk_pixels + 255) / 256, 256, 0, st>>>(w, d_sum, d_sumsq, n_passes);
    tm.end();
    launches++;
    if (acc_done) CU(cudaEventRecord(acc_done, st));
    CU(cudaGetLastError());
    return TAKE_OK;
}

void fill_wave_ptrs(TakeScene *s, Wave &w, const TakeRenderOpts *o, int set = 0) {
    memset(&w, 0, sizeof(w));
    TakeScene::WaveBuffers &b = s->wb[set];
    w.ray = b.ray.as<RayRec>();
    w.hit = b.hit.as<HitRec>();
    w.hit_sorted = b.hit_sorted.as<HitRec>();
    w.path = b.path.as<PathRec>();
    w.pend = b.pend.as<PendRec>();
    w.shadow = b.shadow.as<ShadowRec>();
    w.q_extend[0] = b.q0.as<int32_t>();
    w.q_extend[1] = b.q1.as<int32_t>();
    w.q_shadow = b.q_shadow.as<int32_t>();
    w.pass = b.pass.as<PassCounters>();
    w.totals = s->totals.as<Totals>();
    w.integrator = o->integrator;
    w.max_depth = o->max_depth;
    w.sort_enabled = (o->flags & TAKE_RENDER_NO_SORT) ? 0 : 1;
    const int n_pick = s->dev.env_light ? s->dev.pick_count : s->dev.num_lights;
    w.sort_branch = (w.sort_enabled && (o->integrator == TAKE_INTEGRATOR_ONE_SAMPLE_MIS || o->integrator == TAKE_INTEGRATOR_ONE_SAMPLE_MIS_POWER) && n_pick > 0 &&
                     !env_int("TAKE_NO_SORT_BRANCH", 0)) ? 1 : 0;
    w.seed = o->seed;
    w.rr_start = (o->flags & TAKE_RENDER_RUSSIAN_ROULETTE) ? (o->reserved > 0 ? o->reserved : 3) : 0;
#if TAKE_EXPERIMENTAL
    w.fused_primary = env_int("TAKE_NO_FUSE", 0) ? 0 : 1;
#else
    w.fused_primary = 1;  // pass 0 of extend / shade computes the camera ray itself (no generate pass)
#endif
    w.tile_w = (s->width % 8 == 0 && s->height % 4 == 0 && !env_int("TAKE_NO_TILES", 0)) ? s->width : 0;
    // camera rays that miss are finished inside k_extend (the environment-map extension needs their direction in the
    // shade kernel, and the unsorted debugging mode walks every slot, so both keep the general path)
    w.miss_fast = (w.fused_primary && w.sort_enabled && s->dev.env_rgb == nullptr && !env_int("TAKE_NO_MISS_FAST", 0)) ? 1 : 0;
    // camera rays as warp packets (traverse.cuh: trace_packet4): needs the 4-wide tree and a tree shallow enough for the shared
    // per-warp stack; TAKE_PACKET=0 keeps one ray per thread in pass 0 (A/B runs, and the test that both give the same image)
    {   // TAKE_REFILL: bit 0 = bounce passes, bit 1 = shadow passes on the lane-refill kernels (default: both)
        const int r = env_int("TAKE_REFILL", 3);
        s->refill_extend = r & 1;
        s->refill_shadow = (r >> 1) & 1;
    }
    w.ordered_sort = env_int("TAKE_ORDERED_SORT", s->refill_extend) ? 1 : 0;
    w.packet = (w.fused_primary && s->wide && 3 * s->wide_depth + 1 <= TAKE_PACKET_STACK && env_int("TAKE_PACKET", 1)) ? 1 : 0;
}

// Keep the 4-wide tree resident in L2: the wavefront kernels stream gigabytes of per-path records through the cache
// while every ray keeps returning to the same 36 MB (1 M triangles) of nodes.  The node array is marked "persisting"
// on the streams that run the traversal kernels (as much of it as the device sets aside), misses are streaming.
// Measured on B200 (tools/tune.py): k_extend gains 1-2 %, but the set-aside shrinks the L2 left for the shade and sort
// kernels, which lose more (config 2: +3 % shade; the 10 M-triangle scene: +50 % shade) -- so it is OFF unless
// TAKE_L2_PERSIST=1.  The evict-first hints on the record traffic of the traversal kernels (ld_stream / st_stream in
// wavefront.cuh) get most of the benefit without taking cache away from anyone.
void apply_l2_policy(TakeScene *s, cudaStream_t st) {
    if (!env_int("TAKE_L2_PERSIST", 0) || !s->wide_nodes.p) return;
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, s->device) != cudaSuccess || p.persistingL2CacheMaxSize <= 0) return;
    const size_t want = std::min<size_t>(s->wide_nodes.bytes, (size_t)p.persistingL2CacheMaxSize);
    cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want);
    cudaStreamAttrValue a;
    memset(&a, 0, sizeof(a));
    a.accessPolicyWindow.base_ptr = s->wide_nodes.p;
    a.accessPolicyWindow.num_bytes = std::min<size_t>(s->wide_nodes.bytes, (size_t)p.accessPolicyMaxWindowSize);
    a.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)want / (double)a.accessPolicyWindow.num_bytes);
    a.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    a.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &a);
    cudaGetLastError();
}

int check_opts(const TakeScene *s, const TakeRenderOpts *o) {
    if (!s || !o) return fail(TAKE_E_INVALID, "null argument");
    if (o->integrator < TAKE_INTEGRATOR_MIS || o->integrator > TAKE_INTEGRATOR_ONE_SAMPLE_MIS_POWER)
        return fail(TAKE_E_INVALID, "unknown integrator");
    if (o->integrator == TAKE_INTEGRATOR_ONE_SAMPLE_MIS_POWER) {
        if (s->dev.env_light) return fail(TAKE_E_INVALID, "the power-sampling integrator does not take a sampled environment map");
        if (s->dev.num_lights > 0 && !s->power_ok)   // (the reference would divide by a zero total power)
            return fail(TAKE_E_INVALID, "the power-sampling integrator needs emitters with non-zero power");
    }
    // the reference accepts any -max_depth (render.cpp:14-19); every pass of a wave costs four (mostly empty) launches and a
    // 256-byte counter block, so the bound here is only a sanity limit
    if (o->max_depth < -1 || o->max_depth > TAKE_MAX_DEPTH)
        return fail(TAKE_E_INVALID, "max_depth out of range (-1 .. " + std::to_string(TAKE_MAX_DEPTH) + ")");
    if (o->spp_end < o->spp_begin) return fail(TAKE_E_INVALID, "spp_end < spp_begin");
    if ((o->flags & TAKE_RENDER_RUSSIAN_ROULETTE) && s->dev.env_rgb)
        return fail(TAKE_E_INVALID, "the Russian-roulette extension is not compiled for scenes with an environment map");
    return TAKE_OK;
}

void fill_stats(const Totals &t, TakeStats *stats, const StageTimer &tm, double ms_total, int64_t launches, int64_t waves) {
    if (!stats) return;
    memset(stats, 0, sizeof(*stats));
    stats->samples = (int64_t)t.samples;
    stats->extend_rays = (int64_t)t.extend_rays;
    stats->shadow_rays = (int64_t)t.shadow_rays;
    stats->shaded = (int64_t)t.shaded;
    stats->box_tests = (int64_t)t.box_tests;
    stats->tri_tests = (int64_t)t.tri_tests;
    stats->shadow_box_tests = (int64_t)t.shadow_box_tests;
    stats->shadow_tri_tests = (int64_t)t.shadow_tri_tests;
    stats->miss_after_light_sample = (int64_t)t.miss_after_light_sample;
    stats->kernel_launches = launches;
    stats->waves = waves;
    stats->ms_total = ms_total;
    stats->ms_generate = tm.ms[ST_GENERATE];
    stats->ms_extend = tm.ms[ST_EXTEND];
    stats->ms_shade = tm.ms[ST_SHADE];
    stats->ms_shadow = tm.ms[ST_SHADOW];
    stats->ms_sort = tm.ms[ST_SORT];
    stats->ms_other = tm.ms[ST_OTHER];
}

void read_totals(TakeScene *s, TakeStats *stats, const StageTimer &tm, double ms_total, int64_t launches, int64_t waves) {
    if (!stats) return;
    Totals t;
    cudaMemcpy(&t, s->totals.p, sizeof(t), cudaMemcpyDeviceToHost);
    fill_stats(t, stats, tm, ms_total, launches, waves);
}

}  // namespace

extern "C" {

// Nothing may propagate through the C boundary: the entry points that size host containers by caller-supplied counts or do
// file I/O are function-try-blocks ending in this.
static int translate_exception() noexcept {
    try { throw; }
    catch (const std::bad_alloc &) { return fail(TAKE_E_NOMEM, "out of host memory"); }
    catch (const std::exception &e) { return fail(TAKE_E_INVALID, std::string("internal error: ") + e.what()); }
    catch (...) { return fail(TAKE_E_INVALID, "internal error"); }
}

const char *take_gpu_last_error(void) { return g_error.c_str(); }
int take_exr_fail(int code, const std::string &msg) { return fail(code, msg); }  // for exr_write.cpp
const char *take_gpu_version(void) { return "take_b200 0.1 (sm_100a)"; }

int take_gpu_device_count(int *count) {
    if (!count) return fail(TAKE_E_INVALID, "null argument");
    *count = 0;
    CU(cudaGetDeviceCount(count));
    return TAKE_OK;
}

static int scene_create_from(int device, const TakeSceneDesc *d, HostBuild *hb, TakeScene **out, std::shared_ptr<RefJob> job = nullptr);

// primitive boxes exactly as build_bvh (src/scene.cpp:4-23), then the reference-order tree on a host thread that works on ITS
// OWN copy (the caller's arrays are not touched after the creating call returns)
static std::shared_ptr<RefJob> start_reference_tree(const TakeSceneDesc *d, int threads) {
    auto job = std::make_shared<RefJob>();
    const int64_t n = d->num_prims;
    job->boxes.resize((size_t)n);
    RefJob *j = job.get();
    parallel_chunks(n, threads, [&](int64_t a0, int64_t a1) {
        for (int64_t i = a0; i < a1; ++i) {
            Aabb &b = j->boxes[i];
            const int32_t *id = d->indices + 3 * i;
            if (d->prim_flags[i] & TAKE_PRIM_SPHERE) {
                const double *sp = d->spheres + 4 * (int64_t)id[0];
                for (int a = 0; a < 3; ++a) { b.lo[a] = sp[a] - sp[3]; b.hi[a] = sp[a] + sp[3]; }
            } else {
                const double *p0 = d->positions + 3 * (int64_t)id[0], *p1 = d->positions + 3 * (int64_t)id[1],
                             *p2 = d->positions + 3 * (int64_t)id[2];
                for (int a = 0; a < 3; ++a) {
                    b.lo[a] = std::min(std::min(p0[a], p1[a]), p2[a]);
                    b.hi[a] = std::max(std::max(p0[a], p1[a]), p2[a]);
                }
            }
        }
    });
    job->t0 = now_ms();
    // (leaves a core to the caller: the device build and the first waves need the launching thread)
    const int ref_threads = std::max(1, threads - 1);
    const int delay_ms = env_int("TAKE_REF_DELAY_MS", 0);   // test knob: keeps the tree "still being built" for a while
    job->th = std::thread([j, n, ref_threads, delay_ms] {
        if (delay_ms > 0) std::this_thread::sleep_for(std::chrono::milliseconds(delay_ms));
        build_reference_tree(j->boxes.data(), n, ref_threads, j->tree);
        std::vector<Aabb>().swap(j->boxes);
        j->ms = now_ms() - j->t0;
        j->done.store(true);
    });
    return job;
}

// The fast tree is built on the device (bvh_device.cuh) unless TAKE_DEVICE_BUILD=0 asks for the host's binned-SAH builder
// (the A/B baseline for tree quality; also what the prebuilt / saved-build paths use).
int take_gpu_scene_create(int device, const TakeSceneDesc *d, TakeScene **out) try {
    if (!out) return fail(TAKE_E_INVALID, "null argument");
    *out = nullptr;
    const double t0 = now_ms();
    if (int rc = validate(d)) return rc;
    const double t_validate = now_ms() - t0;
    int rc;
    if (env_int("TAKE_DEVICE_BUILD", 1) && !TAKE_EXPERIMENTAL) {
        rc = scene_create_from(device, d, nullptr, out);
    } else {
        HostBuild hb;
        if ((rc = host_build(d, host_threads(), hb))) return rc;
        rc = scene_create_from(device, d, &hb, out);
    }
    if (rc == TAKE_OK) { (*out)->create_ms[0] = t_validate; (*out)->create_ms[5] = now_ms() - t0; }
    return rc;
} catch (...) { return translate_exception(); }

// Join the background build of the reference-order tree (device-built scenes), upload it and write the tie-break ranks into
// the leaf records.  Every entry point that traces rays calls this first; it is a no-op afterwards.
static int finish_reference_tree(TakeScene *s) {
    if (!s->ref_pending) return TAKE_OK;
    s->ref_pending = false;
    s->ref_job->wait();
    s->build_ms_ref = s->ref_job->ms;
    CU(cudaSetDevice(s->device));
    RefTree &ref = s->ref_job->tree;
    if (int rc = upload(s->ref_nodes, ref.nodes.data(), ref.nodes.size(), s->stream)) return rc;
    if (int rc = upload(s->dfs_rank, ref.dfs_rank.data(), ref.dfs_rank.size(), s->stream)) return rc;
    s->dev.ref_nodes = s->ref_nodes.as<RefNode>();
    s->dev.ref_root = ref.root;
    s->dev.dfs_rank = s->dfs_rank.as<int32_t>();
    const int64_t n = s->dev.num_prims;
    if (n > 0) {
        devbuild::k_patch_ranks<<<(unsigned)((n + 255) / 256), 256, 0, s->stream>>>(n, s->dfs_rank.as<int32_t>(), s->tris.as<double>());
        CU(cudaGetLastError());
    }
    CU(cudaStreamSynchronize(s->stream));
    s->ref_job.reset();                     // the host copy goes with the last scene that used it
    return TAKE_OK;
}

// Fast tree on the device: see bvh_device.cuh.  Needs the scene arrays (positions, indices, flags, spheres) resident and
// s->dev pointing at them.  Fills s->wide_nodes, s->leaf_prims, s->tris (leaf records without ranks).
static int device_build_fast_tree(TakeScene *s, int max_leaf, double &abs_max) {
    using namespace devbuild;
    cudaStream_t st = s->stream;
    const int64_t n = s->dev.num_prims;
    abs_max = 0;
    s->sah_cost = 0;
    CU(s->leaf_prims.ensure(std::max<size_t>((size_t)n * 4, 16)));
    CU(s->tris.ensure(std::max<size_t>((size_t)n * 96, 32)));
    DeviceBuffer wide_tmp, nodes, lo_a, hi_a, lo_b, hi_b, lo_f, hi_f, glob, small, items_a, items_b, kids, icount, iscan, cubtmp;
    CU(wide_tmp.ensure((size_t)std::max<int64_t>(n, 1) * sizeof(WideNode)));
    if (n == 0) {
        k_wide_wrap_root<<<1, 32, 0, st>>>(nullptr, 0, 0, wide_tmp.as<WideNode>());
        CU(cudaGetLastError());
        CU(s->wide_nodes.ensure(sizeof(WideNode)));
        CU(cudaMemcpyAsync(s->wide_nodes.p, wide_tmp.p, sizeof(WideNode), cudaMemcpyDeviceToDevice, st));
        CU(cudaStreamSynchronize(st));
        s->fast_depth = s->wide_depth = 1; s->num_fast_nodes = 1;
        return TAKE_OK;
    }
    const unsigned gb = (unsigned)((n + TAKE_DB_BLOCK - 1) / TAKE_DB_BLOCK);
    CU(lo_a.ensure((size_t)n * 16)); CU(hi_a.ensure((size_t)n * 16));
    CU(lo_b.ensure((size_t)n * 16)); CU(hi_b.ensure((size_t)n * 16));
    CU(lo_f.ensure((size_t)n * 16)); CU(hi_f.ensure((size_t)n * 16));
    CU(glob.ensure(sizeof(Globals)));
    {
        Globals g;
        memset(&g, 0, sizeof(g));
        for (int a = 0; a < 3; ++a) { g.cmin[a] = g.bmin[a] = 0xffffffffu; g.cmax[a] = g.bmax[a] = 0u; }
        CU(cudaMemcpyAsync(glob.p, &g, sizeof(g), cudaMemcpyHostToDevice, st));
    }
    Globals *g = glob.as<Globals>();
    k_prim_boxes<<<gb, TAKE_DB_BLOCK, 0, st>>>(s->dev, n, lo_a.as<float4>(), hi_a.as<float4>(), g);
    CU(cudaGetLastError());
    // SAH levels: huge nodes by several blocks each (chunks), medium nodes by one block each, small nodes by one thread each
    const float c_trav = 1.0f, c_isect = 1.2f;   // SahBuilder's constants (bvh_build.cpp)
    const size_t cap_huge = (size_t)n / TAKE_SAH_HUGE + 4, cap_med = (size_t)n / TAKE_SAH_SMALL + 4, cap_small = (size_t)n / (TAKE_SAH_SMALL / 2) + 8;
    const size_t cap_chunks = (size_t)n / TAKE_SAH_CHUNK + cap_huge + 4;
    DeviceBuffer huge_a, huge_b, med_a, med_b, chunks_a, chunks_b, state_a, state_b, chunk_cnt, chunk_off;
    CU(nodes.ensure((size_t)(2 * n) * sizeof(BNode)));
    CU(huge_a.ensure(cap_huge * sizeof(SItem))); CU(huge_b.ensure(cap_huge * sizeof(SItem)));
    CU(med_a.ensure(cap_med * sizeof(SItem))); CU(med_b.ensure(cap_med * sizeof(SItem)));
    CU(small.ensure(cap_small * sizeof(SItem)));
    CU(chunks_a.ensure(cap_chunks * sizeof(SChunk))); CU(chunks_b.ensure(cap_chunks * sizeof(SChunk)));
    CU(state_a.ensure(cap_huge * sizeof(HugeState))); CU(state_b.ensure(cap_huge * sizeof(HugeState)));
    CU(chunk_cnt.ensure(cap_chunks * 3 * TAKE_SAH_BINS * sizeof(uint32_t))); CU(chunk_off.ensure(cap_chunks * sizeof(int32_t)));
    int32_t *h_cnt = nullptr;   // pinned: {node, huge, medium, small, chunk counts} read back after every level
    CU(cudaMallocHost((void **)&h_cnt, 16 * sizeof(int32_t)));
    struct HostFree { void *p; ~HostFree() { cudaFreeHost(p); } } host_free{h_cnt};
    SItem *hcur = huge_a.as<SItem>(), *hnext = huge_b.as<SItem>(), *mcur = med_a.as<SItem>(), *mnext = med_b.as<SItem>();
    SChunk *ccur = chunks_a.as<SChunk>(), *cnext = chunks_b.as<SChunk>();
    HugeState *scur = state_a.as<HugeState>(), *snext = state_b.as<HugeState>();
    k_sah_root<<<1, 32, 0, st>>>((int32_t)n, g, nodes.as<BNode>(), hcur, mcur, small.as<SItem>(), ccur, scur);
    CU(cudaGetLastError());
    float4 *cur_lo = lo_a.as<float4>(), *cur_hi = hi_a.as<float4>(), *nxt_lo = lo_b.as<float4>(), *nxt_hi = hi_b.as<float4>();
    auto read_counts = [&]() -> int {
        CU(cudaMemcpyAsync(h_cnt, &g->node_count, 8 * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        if (h_cnt[0] < 1 || h_cnt[0] > 2 * n || h_cnt[1] < 0 || (size_t)h_cnt[1] > cap_huge || h_cnt[2] < 0 || (size_t)h_cnt[2] > cap_med ||
            h_cnt[3] < 0 || (size_t)h_cnt[3] > cap_small || h_cnt[4] < 0 || (size_t)h_cnt[4] > cap_chunks)
            return fail(TAKE_E_CUDA, "device BVH build: inconsistent level");
        return TAKE_OK;
    };
    auto run_small = [&](int32_t n_small, const float4 *src_lo, const float4 *src_hi) -> int {
        if (n_small <= 0) return TAKE_OK;
        k_sah_small<<<(unsigned)((n_small + 63) / 64), 64, 0, st>>>(small.as<SItem>(), n_small, src_lo, src_hi, lo_f.as<float4>(), hi_f.as<float4>(),
                                                                  nodes.as<BNode>(), g, max_leaf, c_trav, c_isect);
        CU(cudaGetLastError());
        return TAKE_OK;
    };
    if (int rc = read_counts()) return rc;
    if (int rc = run_small(h_cnt[3], cur_lo, cur_hi)) return rc;   // the whole scene is one small node
    int32_t n_huge = h_cnt[1], n_med = h_cnt[2], n_chunks = h_cnt[4];
    int levels = 0;
    while (n_huge > 0 || n_med > 0) {
        CU(cudaMemsetAsync(&g->huge_count, 0, 4 * sizeof(int32_t), st));
        if (n_huge > 0) {
            const int per = 3 * TAKE_SAH_BINS;
            k_huge_clear<<<(unsigned)((n_huge * per + 127) / 128), 128, 0, st>>>(scur, n_huge);
            k_huge_bin<<<(unsigned)n_chunks, TAKE_DB_BLOCK, 0, st>>>(hcur, ccur, cur_lo, cur_hi, scur, chunk_cnt.as<uint32_t>());
            k_huge_choose<<<(unsigned)n_huge, TAKE_DB_BLOCK, 0, st>>>(hcur, scur, chunk_cnt.as<uint32_t>(), chunk_off.as<int32_t>());
            k_huge_partition<<<(unsigned)n_chunks, TAKE_DB_BLOCK, 0, st>>>(hcur, ccur, cur_lo, cur_hi, nxt_lo, nxt_hi, scur, chunk_off.as<int32_t>());
            k_huge_emit<<<(unsigned)((n_huge + 63) / 64), 64, 0, st>>>(hcur, n_huge, scur, nodes.as<BNode>(), g, hnext, mnext, small.as<SItem>(), cnext, snext);
            CU(cudaGetLastError());
        }
        if (n_med > 0) {
            k_sah_split<<<(unsigned)n_med, TAKE_DB_BLOCK, 0, st>>>(mcur, cur_lo, cur_hi, nxt_lo, nxt_hi, nodes.as<BNode>(), g, hnext, mnext,
                                                                   small.as<SItem>(), cnext, snext);
            CU(cudaGetLastError());
        }
        if (int rc = read_counts()) return rc;
        if (int rc = run_small(h_cnt[3], nxt_lo, nxt_hi)) return rc;
        std::swap(cur_lo, nxt_lo); std::swap(cur_hi, nxt_hi);
        std::swap(hcur, hnext); std::swap(mcur, mnext); std::swap(ccur, cnext); std::swap(scur, snext);
        n_huge = h_cnt[1]; n_med = h_cnt[2]; n_chunks = h_cnt[4];
        if (++levels > 4096) return fail(TAKE_E_CUDA, "device BVH build: no progress");
    }
    k_leaf_prims<<<gb, TAKE_DB_BLOCK, 0, st>>>(n, lo_f.as<float4>(), s->leaf_prims.as<int32_t>());
    CU(cudaGetLastError());
    // wide collapse, breadth first
    const int32_t root = 0;
    BNode h_root;
    CU(cudaMemcpyAsync(&h_root, nodes.as<BNode>() + root, sizeof(BNode), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    int32_t n_wide = 1, depth = 1;
    if (h_root.leaf) {
        k_wide_wrap_root<<<1, 32, 0, st>>>(nodes.as<BNode>(), root, n, wide_tmp.as<WideNode>());
        CU(cudaGetLastError());
    } else {
        size_t tmp_bytes = 0;
        CU(icount.ensure((size_t)n * 4)); CU(iscan.ensure((size_t)n * 4));
        CU(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, icount.as<uint32_t>(), iscan.as<uint32_t>(), (int)n, st));
        tmp_bytes = std::max<size_t>(tmp_bytes, 256);
        CU(cubtmp.ensure(tmp_bytes));
        CU(items_a.ensure((size_t)n * sizeof(WorkItem))); CU(items_b.ensure((size_t)n * sizeof(WorkItem)));
        CU(kids.ensure((size_t)n * sizeof(Kids)));
        WorkItem first_item = {root, 0, 0, 0};
        CU(cudaMemcpyAsync(items_a.p, &first_item, sizeof(first_item), cudaMemcpyHostToDevice, st));
        WorkItem *ia = items_a.as<WorkItem>(), *ib = items_b.as<WorkItem>();
        int32_t n_items = 1;
        uint32_t *h32 = reinterpret_cast<uint32_t *>(h_cnt) + 8;
        while (n_items > 0) {
            const unsigned gi = (unsigned)((n_items + TAKE_DB_BLOCK - 1) / TAKE_DB_BLOCK);
            k_wide_kids<<<gi, TAKE_DB_BLOCK, 0, st>>>(n_items, ia, nodes.as<BNode>(), kids.as<Kids>(), icount.as<uint32_t>());
            CU(cub::DeviceScan::ExclusiveSum(cubtmp.p, tmp_bytes, icount.as<uint32_t>(), iscan.as<uint32_t>(), n_items, st));
            CU(cudaMemcpyAsync(h32, iscan.as<uint32_t>() + (n_items - 1), 4, cudaMemcpyDeviceToHost, st));
            CU(cudaMemcpyAsync(h32 + 1, icount.as<uint32_t>() + (n_items - 1), 4, cudaMemcpyDeviceToHost, st));
            k_wide_emit<<<gi, TAKE_DB_BLOCK, 0, st>>>(n_items, ia, kids.as<Kids>(), iscan.as<uint32_t>(), nodes.as<BNode>(), wide_tmp.as<WideNode>(),
                                                     n_wide, ib);
            CU(cudaGetLastError());
            CU(cudaStreamSynchronize(st));
            const int32_t inner = (int32_t)(h32[0] + h32[1]);
            if (inner < 0 || (int64_t)n_wide + inner > n) return fail(TAKE_E_CUDA, "device BVH build: wide node count out of range");
            n_wide += inner;
            n_items = inner;
            std::swap(ia, ib);
            if (inner > 0) ++depth;
        }
    }
    CU(s->wide_nodes.ensure((size_t)n_wide * sizeof(WideNode)));
    CU(cudaMemcpyAsync(s->wide_nodes.p, wide_tmp.p, (size_t)n_wide * sizeof(WideNode), cudaMemcpyDeviceToDevice, st));
    k_leaf_records<<<gb, TAKE_DB_BLOCK, 0, st>>>(s->dev, n, s->leaf_prims.as<int32_t>(), s->tris.as<double>());
    CU(cudaGetLastError());
    Globals hg;
    CU(cudaMemcpyAsync(&hg, glob.p, sizeof(hg), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    memcpy(&abs_max, &hg.abs_max_bits, 8);
    s->fast_depth = s->wide_depth = depth;
    s->num_fast_nodes = n_wide;
    return TAKE_OK;
}

// Upload a scene to `device`.  hb != nullptr: acceleration structures built on the host (shared by all replicas of a
// multi-GPU render, or loaded from a file); hb == nullptr: the fast tree is built on the device and the reference-order
// tree on a background host thread.
static int scene_create_from(int device, const TakeSceneDesc *d, HostBuild *hb, TakeScene **out, std::shared_ptr<RefJob> job) {
    *out = nullptr;
    // TAKE_TIMING=1: every step of the call on stderr (development aid)
    const bool timing = env_int("TAKE_TIMING", 0) != 0;
    const double t_begin = now_ms();
    double t_last = t_begin;
    auto mark = [&](const char *what) {
        if (!timing) return;
        const double t = now_ms();
        fprintf(stderr, "[take_gpu] scene_create %-28s %8.2f ms (at %8.2f)\n", what, t - t_last, t - t_begin);
        t_last = t;
    };
    CU(cudaSetDevice(device));
    TakeScene *s = new TakeScene;
    struct Guard { TakeScene *&p; bool ok = false; ~Guard() { if (!ok) { delete p; p = nullptr; } } } guard{s};
    s->device = device;
    mark("set device");
    int sm_count = 0;   // (cudaGetDeviceProperties fills ~1 KB of fields through dozens of driver queries: one attribute is enough)
    CU(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, device));
    s->sm_count = sm_count;
    mark("device attribute");
    CU(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    mark("stream");
    cudaStream_t st = s->stream;
    const int64_t n = d->num_prims;
    const int threads = host_threads();
    double abs_max = hb ? hb->abs_max : 0.0;
    double t_mark = now_ms();
    std::vector<uint8_t> mtype((size_t)n);
    std::atomic<int> any_uv_flag(0);
    parallel_chunks(n, threads, [&](int64_t a0, int64_t a1) {
        bool uv = false;
        for (int64_t i = a0; i < a1; ++i) {
            mtype[i] = (uint8_t)d->materials[d->prim_material[i]].type;
            uv = uv || (d->prim_flags[i] & TAKE_PRIM_HAS_UVS);
        }
        if (uv) any_uv_flag.store(1);
    });

    int rc;
    if (hb) {
        s->build_ms_ref = hb->ms_ref;
        s->build_ms_fast = hb->ms_fast;
        s->fast_depth = hb->fast.depth;
        s->wide_depth = hb->fast.wide_depth;
        s->sah_cost = hb->fast.sah_cost;
        s->num_fast_nodes = (int64_t)hb->fast.nodes.size();
#if TAKE_EXPERIMENTAL
        if ((rc = upload(s->nodes, hb->fast.nodes.data(), hb->fast.nodes.size(), st))) return rc;  // binary image: A/B runs only
#endif
        if ((rc = upload(s->wide_nodes, hb->fast.wide.data(), hb->fast.wide.size(), st))) return rc;
        if ((rc = upload(s->tris, hb->tris.data(), hb->tris.size(), st))) return rc;
        if ((rc = upload(s->ref_nodes, hb->ref.nodes.data(), hb->ref.nodes.size(), st))) return rc;
        if ((rc = upload(s->dfs_rank, hb->ref.dfs_rank.data(), hb->ref.dfs_rank.size(), st))) return rc;
    } else {
        // the reference-order tree goes on in the background (one job for all replicas of a multi-GPU handle)
        mark("material types");
        s->ref_job = job ? job : start_reference_tree(d, threads);
        s->ref_pending = true;
        s->create_ms[2] = now_ms() - t_mark;
        mark("boxes + reference thread");
        CU(s->ref_nodes.ensure(64));
        CU(s->dfs_rank.ensure(16));
    }
    t_mark = now_ms();
    if ((rc = upload(s->positions, d->positions, (size_t)d->num_vertices * 3, st))) return rc;
    if ((rc = upload(s->normals, d->normals, (size_t)d->num_vertices * 3, st))) return rc;
    if ((rc = upload(s->uvs, d->uvs, (size_t)d->num_vertices * 2, st))) return rc;
    if ((rc = upload(s->indices, d->indices, (size_t)n * 3, st))) return rc;
    if ((rc = upload(s->prim_material, d->prim_material, (size_t)n, st))) return rc;
    if ((rc = upload(s->prim_light, d->prim_light, (size_t)n, st))) return rc;
    if ((rc = upload(s->prim_flags, d->prim_flags, (size_t)n, st))) return rc;
    if ((rc = upload(s->prim_mtype, mtype.data(), mtype.size(), st))) return rc;
    if ((rc = upload(s->spheres, d->spheres, (size_t)d->num_spheres * 4, st))) return rc;
    if ((rc = upload(s->materials, d->materials, (size_t)d->num_materials, st))) return rc;
    if ((rc = upload(s->lights, d->lights, (size_t)d->num_lights, st))) return rc;
    {   // tables of the power-proportional light pick: light_power (src/light.cpp:25-30) = luminance x get_area x pi per light,
        // summed in light order; pmf = power / total; cdf = the N + 1 running sums from 0 to exactly 1 that sample_light_power
        // walks (light.cpp:9-17).  The reference never fills these (dead code there); oracle/ref_harness.cpp builds them the
        // same way for the pin.
        const size_t nl = (size_t)d->num_lights;
        std::vector<double> pmf(nl, 0.0), cdf(nl + 1, 0.0), power(nl, 0.0);
        double total = 0;
        for (size_t i = 0; i < nl; ++i) {
            const TakeLightDesc &l = d->lights[i];
            if (l.kind == TAKE_LIGHT_AREA) {
                const int32_t *id = d->indices + 3 * (int64_t)l.prim_id;
                double area;
                if (d->prim_flags[l.prim_id] & TAKE_PRIM_SPHERE) {   // get_area_op, src/shape.cpp:171-184
                    const double r = d->spheres[4 * (int64_t)id[0] + 3];
                    area = 4 * TAKE_PI * r * r;
                } else {
                    const double *p0 = d->positions + 3 * (int64_t)id[0], *p1 = d->positions + 3 * (int64_t)id[1], *p2 = d->positions + 3 * (int64_t)id[2];
                    const double a[3] = {p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2]}, b[3] = {p2[0] - p0[0], p2[1] - p0[1], p2[2] - p0[2]};
                    const double c[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
                    area = sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]) / 2;
                }
                power[i] = (l.intensity[0] * 0.212671 + l.intensity[1] * 0.715160 + l.intensity[2] * 0.072169) * area * TAKE_PI;
            }
            total += power[i];
        }
        for (size_t i = 0; i < nl; ++i) { pmf[i] = power[i] / total; cdf[i + 1] = cdf[i] + pmf[i]; }
        if (nl) cdf[nl] = 1;
        if ((rc = upload(s->light_pmf, pmf.data(), pmf.size(), st))) return rc;
        if ((rc = upload(s->light_cdf, cdf.data(), cdf.size(), st))) return rc;
        s->power_ok = nl > 0 && total > 0 && std::isfinite(total);
    }
    std::vector<DevTexture> tex((size_t)d->num_textures);
    for (int i = 0; i < d->num_textures; ++i) {
        const TakeTextureDesc &t = d->textures[i];
        if (t.width <= 0 || t.height <= 0 || !t.rgb) return fail(TAKE_E_INVALID, "bad texture");
        DeviceBuffer *b = new DeviceBuffer;
        s->tex_data.push_back(b);
        if ((rc = upload(*b, t.rgb, (size_t)t.width * t.height * 3, st))) return rc;
        tex[i].w = t.width; tex[i].h = t.height; tex[i].rgb = b->as<double>();
    }
    if ((rc = upload(s->textures, tex.data(), tex.size(), st))) return rc;
    // EXTENSION: environment map + marginal / conditional CDF tables of luminance x sin(theta at the row centre).
    // (The CPU checker used by the tests builds its tables with the same loops and operation order, so sampling decisions agree.)
    std::vector<double> env_marg, env_cond;
    double env_total = 0;
    const bool has_env = d->env_rgb && d->env_width > 0 && d->env_height > 0;
    if (has_env) {
        const int W = d->env_width, H = d->env_height;
        env_marg.assign((size_t)H + 1, 0.0);
        env_cond.assign((size_t)H * (W + 1), 0.0);
        for (int j = 0; j < H; ++j) {
            double *c = &env_cond[(size_t)j * (W + 1)];
            const double sj = sin(TAKE_PI * (j + 0.5) / H);
            for (int i = 0; i < W; ++i) {
                const double *p = d->env_rgb + 3 * ((size_t)j * W + i);
                const double lum = p[0] * 0.212671 + p[1] * 0.715160 + p[2] * 0.072169;
                c[i + 1] = c[i] + lum * sj;
            }
            env_marg[j + 1] = env_marg[j] + c[W];
        }
        env_total = env_marg[H];
        if ((rc = upload(s->env_rgb, d->env_rgb, (size_t)W * H * 3, st))) return rc;
        if ((rc = upload(s->env_marg, env_marg.data(), env_marg.size(), st))) return rc;
        if ((rc = upload(s->env_cond, env_cond.data(), env_cond.size(), st))) return rc;
    }
    CU(cudaStreamSynchronize(st));

    DevScene &v = s->dev;
    v.nodes = s->nodes.as<float4>();
    v.wide_nodes = s->wide_nodes.as<float4>();
    v.tris = s->tris.as<double2>();
    v.ref_nodes = s->ref_nodes.as<RefNode>();
    v.ref_root = hb ? hb->ref.root : -1;
    s->create_ms[1] = now_ms() - t_mark;
    mark("uploads");
    v.positions = s->positions.as<double>(); v.normals = s->normals.as<double>(); v.uvs = s->uvs.as<double>();
    v.indices = s->indices.as<int32_t>(); v.prim_material = s->prim_material.as<int32_t>();
    v.prim_light = s->prim_light.as<int32_t>(); v.dfs_rank = s->dfs_rank.as<int32_t>();
    v.prim_flags = s->prim_flags.as<uint8_t>(); v.prim_mtype = s->prim_mtype.as<uint8_t>();
    v.spheres = s->spheres.as<double>();
    v.materials = s->materials.as<TakeMaterialDesc>();
    v.lights = s->lights.as<TakeLightDesc>();
    v.textures = s->textures.as<DevTexture>();
    CU(s->tie_count.ensure(16));
    CU(cudaMemsetAsync(s->tie_count.p, 0, 16, st));
    v.tie_count = s->tie_count.as<unsigned long long>();
    v.num_lights = d->num_lights; v.num_materials = d->num_materials;
    v.light_pmf = s->light_pmf.as<double>(); v.light_cdf = s->light_cdf.as<double>();
    v.env_rgb = has_env ? s->env_rgb.as<double>() : nullptr;
    v.env_marg = has_env ? s->env_marg.as<double>() : nullptr;
    v.env_cond = has_env ? s->env_cond.as<double>() : nullptr;
    v.env_total = env_total;
    v.env_w = has_env ? d->env_width : 0;
    v.env_h = has_env ? d->env_height : 0;
    v.env_light = (has_env && d->env_sample) ? 1 : 0;
    v.pick_count = d->num_lights + v.env_light;
    v.num_prims = n;
    // camera basis: src/render.cpp:37-44, same operations on the host's libm
    const TakeCamera &c = d->camera;
    s->width = v.width = c.width; s->height = v.height = c.height;
    {
        auto sub3 = [](const double *a, const double *b, double *o) { for (int i = 0; i < 3; ++i) o[i] = a[i] - b[i]; };
        auto norm3 = [](double *a) {
            double l = sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
            if (l <= 0) { a[0] = a[1] = a[2] = 0; return; }
            double inv = 1.0 / l;
            for (int i = 0; i < 3; ++i) a[i] *= inv;
        };
        auto cross3 = [](const double *a, const double *b, double *o) {
            o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
        };
        double theta = c.vfov / 180 * TAKE_PI;
        double h = tan(theta / 2);
        v.viewport_h = 2 * h;
        v.viewport_w = v.viewport_h / c.height * c.width;
        double w[3], u[3], vv[3];
        sub3(c.lookfrom, c.lookat, w); norm3(w);
        cross3(c.up, w, u); norm3(u);
        cross3(w, u, vv);
        v.lookfrom = {c.lookfrom[0], c.lookfrom[1], c.lookfrom[2]};
        v.cam_u = {u[0], u[1], u[2]}; v.cam_v = {vv[0], vv[1], vv[2]}; v.cam_w = {w[0], w[1], w[2]};
    }
    v.background = {d->background[0], d->background[1], d->background[2]};
    if (!hb) {
        t_mark = now_ms();
        const int max_leaf = std::min(8, std::max(1, env_int("TAKE_BVH_MAX_LEAF", 4)));
        if ((rc = device_build_fast_tree(s, max_leaf, abs_max))) return rc;
        if (3 * s->fast_depth + 1 > TAKE_STACK_SMEM + TAKE_STACK_LOCAL)
            return fail(TAKE_E_INVALID, "acceleration tree too deep (" + std::to_string(s->fast_depth) + " wide levels)");
        if (!std::isfinite(abs_max)) return fail(TAKE_E_INVALID, "non-finite scene extent");
        v.wide_nodes = s->wide_nodes.as<float4>();
        v.tris = s->tris.as<double2>();
        s->device_built = true;
        s->build_ms_fast = s->create_ms[3] = now_ms() - t_mark;
        mark("device build");
    }
    v.fast_depth = s->fast_depth;
    v.abs_max = (float)abs_max;

    // per-primitive shading records, derived on the device from the arrays uploaded above (shading.cuh: ShadeRec)
    t_mark = now_ms();
    {
        const bool any_uv = any_uv_flag.load() != 0;
        v.shade_stride = any_uv ? 20 : 16;
        CU(s->shade_recs.ensure(std::max<size_t>((size_t)n * v.shade_stride * sizeof(double), 32)));
        v.shade_recs = s->shade_recs.as<double>();
        if (n > 0) {
            k_build_shade_recs<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(v, s->shade_recs.as<double>());
            CU(cudaGetLastError());
        }
        // per-light records (shading.cuh: LightRec)
        CU(s->light_recs.ensure(std::max<size_t>((size_t)d->num_lights * TAKE_LIGHT_REC_STRIDE * sizeof(double), 256)));
        v.light_recs = s->light_recs.as<double>();
        if (d->num_lights > 0) {
            k_build_light_recs<<<(unsigned)((d->num_lights + 127) / 128), 128, 0, st>>>(v, s->light_recs.as<double>());
            CU(cudaGetLastError());
        }
        CU(cudaStreamSynchronize(st));
    }
    s->create_ms[4] = now_ms() - t_mark;
    mark("records");

    // persistent-kernel launch widths: every SM filled to the occupancy the kernel allows
    auto blocks_for = [&](const void *fn) {
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, 128, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
        return per_sm * s->sm_count;
    };
    s->blocks_extend = blocks_for((const void *)k_extend<false, true>);
    s->blocks_extend_primary = blocks_for((const void *)k_extend_primary<false>);
    s->blocks_shadow = blocks_for((const void *)k_shadow<false, true>);
    s->blocks_isect = blocks_for((const void *)k_intersect_fast<false, true>);
    s->blocks_occl = blocks_for((const void *)k_intersect_fast<true, true>);
#if TAKE_EXPERIMENTAL
    s->wide = env_int("TAKE_BVH_WIDTH", 4) == 4;
    if (!s->wide) {
        s->blocks_extend = blocks_for((const void *)k_extend<false, false>);
        s->blocks_shadow = blocks_for((const void *)k_shadow<false, false>);
        s->blocks_isect = blocks_for((const void *)k_intersect_fast<false, false>);
        s->blocks_occl = blocks_for((const void *)k_intersect_fast<true, false>);
    }
#endif
    s->blocks_extend_refill = blocks_for((const void *)k_extend_refill<false>);
    s->blocks_shadow_refill = blocks_for((const void *)k_shadow_refill<false>);
    mark("occupancy queries");
    CU(s->fetch.ensure(256));
    apply_l2_policy(s, s->stream);
    mark("tail");
    guard.ok = true;
    *out = s;
    return TAKE_OK;
}

int take_gpu_scene_destroy(TakeScene *s) {
    if (!s) return TAKE_OK;
    cudaSetDevice(s->device);
    delete s;
    return TAKE_OK;
}

void *take_gpu_scene_stream(TakeScene *s) { return s ? (void *)s->stream : nullptr; }

// Hands the device memory this library keeps cached in the device's pool (freed scenes, builder scratch, wave buffers of
// earlier sizes) back to the driver.
int take_gpu_release_cached_memory(int device) {
    cudaMemPool_t pool;
    CU(cudaSetDevice(device));
    CU(cudaDeviceSynchronize());
    CU(cudaDeviceGetDefaultMemPool(&pool, device));
    CU(cudaMemPoolTrimTo(pool, 0));
    return TAKE_OK;
}

// diagnostics: out[0..5] = reference-tree build ms, fast-tree build ms, fast-tree depth, SAH cost, #fast nodes, #SMs
int take_gpu_scene_info(TakeScene *s, double *out) try {
    if (!s || !out) return fail(TAKE_E_INVALID, "null argument");
    if (int rc = finish_reference_tree(s)) return rc;   // (its build time is one of the figures)
    out[0] = s->build_ms_ref; out[1] = s->build_ms_fast; out[2] = s->fast_depth; out[3] = s->sah_cost;
    out[4] = (double)s->num_fast_nodes; out[5] = s->sm_count;
    return TAKE_OK;
} catch (...) { return translate_exception(); }

// out[0..7] = ms of take_gpu_scene_create: validate, scene upload, primitive boxes for the reference-order tree (host), fast
// tree (device build; 0 for host-built scenes), shading / light records, whole call, 1 if the fast tree was built on the
// device, 1 if the reference-order tree is still being built in the background
int take_gpu_scene_create_timings(TakeScene *s, double *out) {
    if (!s || !out) return fail(TAKE_E_INVALID, "null argument");
    for (int i = 0; i < 6; ++i) out[i] = s->create_ms[i];
    out[6] = s->device_built ? 1 : 0;
    out[7] = s->ref_pending ? 1 : 0;
    return TAKE_OK;
}

// out[0] = renders that ran before the reference-order tree had arrived, out[1] = how many of them had to be repeated because
// a tie-break rank mattered
int take_gpu_scene_provisional_stats(TakeScene *s, int64_t *out) {
    if (!s || !out) return fail(TAKE_E_INVALID, "null argument");
    out[0] = s->provisional_renders; out[1] = s->provisional_reruns;
    return TAKE_OK;
}

// Copies the fast tree out of the device (128-byte WideNodes, and the primitive id of every leaf slot) for inspection;
// returns the number of wide nodes.  leaf_prims is only kept for device-built scenes (NULL otherwise is fine).
int64_t take_gpu_scene_debug_tree(TakeScene *s, void *wide_nodes, int32_t *leaf_prims) {
    if (!s) return fail(TAKE_E_INVALID, "null argument");
    if (cudaSetDevice(s->device) != cudaSuccess) return fail(TAKE_E_CUDA, "cudaSetDevice");
    const int64_t nw = s->device_built ? s->num_fast_nodes : (int64_t)(s->wide_nodes.bytes / sizeof(WideNode));
    if (wide_nodes && cudaMemcpy(wide_nodes, s->wide_nodes.p, (size_t)nw * sizeof(WideNode), cudaMemcpyDeviceToHost) != cudaSuccess)
        return fail(TAKE_E_CUDA, "copying the tree failed");
    if (leaf_prims) {
        if (!s->device_built) return fail(TAKE_E_INVALID, "leaf slots are only kept for device-built scenes");
        if (cudaMemcpy(leaf_prims, s->leaf_prims.p, (size_t)s->dev.num_prims * 4, cudaMemcpyDeviceToHost) != cudaSuccess)
            return fail(TAKE_E_CUDA, "copying the leaf slots failed");
    }
    return nw;
}

// ---- host-only diagnostics: the acceleration structures scene_create would upload, without touching CUDA ----
struct TakeHostBuild {
    HostBuild hb;
};

int take_gpu_host_build(const TakeSceneDesc *d, TakeHostBuild **out) try {
    if (!out) return fail(TAKE_E_INVALID, "null argument");
    *out = nullptr;
    if (int rc = validate(d)) return rc;
    TakeHostBuild *h = new TakeHostBuild;
    if (int rc = host_build(d, host_threads(), h->hb)) { delete h; return rc; }
    *out = h;
    return TAKE_OK;
} catch (...) { return translate_exception(); }
// out[0..7] = #reference nodes, reference root, #fast nodes, #prims, fast depth, abs_max, build ms (ref), build ms (fast)
int take_gpu_host_build_info(TakeHostBuild *h, double *out) {
    if (!h || !out) return fail(TAKE_E_INVALID, "null argument");
    out[0] = (double)h->hb.ref.nodes.size(); out[1] = h->hb.ref.root; out[2] = (double)h->hb.fast.nodes.size();
    out[3] = (double)h->hb.fast.leaf_prims.size(); out[4] = h->hb.fast.depth; out[5] = h->hb.abs_max;
    out[6] = h->hb.ms_ref; out[7] = h->hb.ms_fast;
    return TAKE_OK;
}
// any pointer may be NULL.  ref_nodes: 64 B each (6 doubles box, 4 int32 left,right,prim,pad); fast_nodes: 64 B each
// (12 floats, 4 int32); leaf_records: 12 doubles per slot.
int take_gpu_host_build_copy(TakeHostBuild *h, void *ref_nodes, int32_t *dfs_rank, void *fast_nodes, int32_t *leaf_prims,
                             double *leaf_records) {
    if (!h) return fail(TAKE_E_INVALID, "null argument");
    const HostBuild &b = h->hb;
    if (ref_nodes) memcpy(ref_nodes, b.ref.nodes.data(), b.ref.nodes.size() * sizeof(RefNode));
    if (dfs_rank) memcpy(dfs_rank, b.ref.dfs_rank.data(), b.ref.dfs_rank.size() * 4);
    if (fast_nodes) memcpy(fast_nodes, b.fast.nodes.data(), b.fast.nodes.size() * sizeof(FastNode));
    if (leaf_prims) memcpy(leaf_prims, b.fast.leaf_prims.data(), b.fast.leaf_prims.size() * 4);
    if (leaf_records) memcpy(leaf_records, b.tris.data(), b.tris.size() * 8);
    return TAKE_OK;
}
// 4-wide nodes (128 B each: 24 floats, 4 int32 links, 4 int32 counts); returns their number, copies if `wide_nodes` != NULL
int64_t take_gpu_host_build_wide(TakeHostBuild *h, void *wide_nodes) {
    if (!h) return fail(TAKE_E_INVALID, "null argument");
    if (wide_nodes) memcpy(wide_nodes, h->hb.fast.wide.data(), h->hb.fast.wide.size() * sizeof(WideNode));
    return (int64_t)h->hb.fast.wide.size();
}
int take_gpu_host_build_free(TakeHostBuild *h) {
    delete h;
    return TAKE_OK;
}

}  // extern "C"

// ---- build once, create many (one process per GPU: a single rank runs the host builders) ---------------------------
namespace {
struct HbHeader {
    char magic[8];           // "TAKEHB04"
    uint64_t geom_hash;
    int64_t num_prims, n_ref, n_rank, n_fast, n_wide, n_leaf, n_tris;
    int32_t ref_root, depth, wide_depth, pad;
    double sah_cost, abs_max, ms_ref, ms_fast;
    uint64_t payload_hash;   // of the arrays that follow (the links in them are trusted by the device code)
    uint64_t header_hash;    // of this header with both hash fields zeroed
};
uint64_t payload_hash(const HostBuild &b) {
    const int threads = (int)std::max(1u, std::thread::hardware_concurrency());
    uint64_t h[6] = {hash_array(b.ref.nodes.data(), b.ref.nodes.size() * sizeof(RefNode), threads),
                     hash_array(b.ref.dfs_rank.data(), b.ref.dfs_rank.size() * 4, threads),
                     hash_array(b.fast.nodes.data(), b.fast.nodes.size() * sizeof(FastNode), threads),
                     hash_array(b.fast.wide.data(), b.fast.wide.size() * sizeof(WideNode), threads),
                     hash_array(b.fast.leaf_prims.data(), b.fast.leaf_prims.size() * 4, threads),
                     hash_array(b.tris.data(), b.tris.size() * 8, threads)};
    return fnv1a(h, sizeof(h));
}
template <typename T>
bool put(FILE *f, const std::vector<T> &v) { return v.empty() || fwrite(v.data(), sizeof(T), v.size(), f) == v.size(); }
template <typename T>
bool get(FILE *f, std::vector<T> &v, int64_t n) {
    if (n < 0) return false;
    v.resize((size_t)n);
    return n == 0 || fread(v.data(), sizeof(T), (size_t)n, f) == (size_t)n;
}
}  // namespace

extern "C" {

int64_t take_gpu_selftest_sort(int64_t n, int64_t distinct, int32_t pattern, int32_t threads, uint64_t seed) {
    return sort_selftest(n, distinct, pattern, threads, seed);
}

// The header is hashed too (with its own hash field zeroed): the scalars in it are trusted by the device code as much as
// the arrays are -- ref_root is dereferenced by the exact traversal, abs_max sizes the conservative padding of every box
// test, the depths bound the traversal stack.
static uint64_t header_hash(HbHeader hd) {
    hd.payload_hash = 0;
    hd.header_hash = 0;
    return fnv1a(&hd, sizeof(hd));
}

int take_gpu_host_build_save(TakeHostBuild *h, const char *path) {
    if (!h || !path) return fail(TAKE_E_INVALID, "null argument");
    const HostBuild &b = h->hb;
    HbHeader hd;
    memset(&hd, 0, sizeof(hd));
    memcpy(hd.magic, "TAKEHB04", 8);
    hd.geom_hash = b.geom_hash; hd.num_prims = b.num_prims;
    hd.n_ref = (int64_t)b.ref.nodes.size(); hd.n_rank = (int64_t)b.ref.dfs_rank.size(); hd.n_fast = (int64_t)b.fast.nodes.size();
    hd.n_wide = (int64_t)b.fast.wide.size(); hd.n_leaf = (int64_t)b.fast.leaf_prims.size(); hd.n_tris = (int64_t)b.tris.size();
    hd.ref_root = b.ref.root; hd.depth = b.fast.depth; hd.wide_depth = b.fast.wide_depth;
    hd.sah_cost = b.fast.sah_cost; hd.abs_max = b.abs_max; hd.ms_ref = b.ms_ref; hd.ms_fast = b.ms_fast;
    hd.header_hash = header_hash(hd);
    hd.payload_hash = payload_hash(b);
    const std::string tmp = std::string(path) + ".part";  // readers never see a half-written file
    FILE *f = fopen(tmp.c_str(), "wbx");                  // "x": never follow or clobber something already there
    if (!f) { remove(tmp.c_str()); f = fopen(tmp.c_str(), "wbx"); }
    if (!f) return fail(TAKE_E_INVALID, std::string("cannot write ") + tmp);
    bool ok = fwrite(&hd, sizeof(hd), 1, f) == 1 && put(f, b.ref.nodes) && put(f, b.ref.dfs_rank) && put(f, b.fast.nodes) &&
              put(f, b.fast.wide) && put(f, b.fast.leaf_prims) && put(f, b.tris);
    ok = (fclose(f) == 0) && ok;
    if (!ok || rename(tmp.c_str(), path) != 0) { remove(tmp.c_str()); return fail(TAKE_E_INVALID, std::string("write failed: ") + path); }
    return TAKE_OK;
}

int take_gpu_host_build_load(const char *path, TakeHostBuild **out) try {
    if (!path || !out) return fail(TAKE_E_INVALID, "null argument");
    *out = nullptr;
    FILE *f = fopen(path, "rb");
    if (!f) return fail(TAKE_E_INVALID, std::string("cannot read ") + path);
    const std::string bad = std::string("not a host-build file of this library: ") + path;
    HbHeader hd;
    bool ok = fread(&hd, sizeof(hd), 1, f) == 1 && memcmp(hd.magic, "TAKEHB04", 8) == 0 && header_hash(hd) == hd.header_hash;
    // sizes are fixed by the primitive count: anything else is not a file this library wrote
    ok = ok && hd.num_prims >= 0 && hd.num_prims < (1 << 28) && hd.n_leaf == hd.num_prims && hd.n_rank == hd.num_prims &&
         hd.n_tris == 12 * hd.num_prims && hd.n_ref == (hd.num_prims > 0 ? 2 * hd.num_prims - 1 : 0) && hd.n_fast >= 1 &&
         hd.n_fast <= std::max<int64_t>(hd.num_prims, 1) && hd.n_wide >= 1 && hd.n_wide <= hd.n_fast;
    // the scalars the device code trusts
    ok = ok && (hd.num_prims > 0 ? (hd.ref_root >= 0 && hd.ref_root < hd.n_ref) : hd.ref_root == -1) && std::isfinite(hd.abs_max) &&
         hd.abs_max >= 0 && hd.depth >= 0 && hd.depth <= TAKE_STACK_SMEM + TAKE_STACK_LOCAL && hd.wide_depth >= 0 &&
         hd.wide_depth <= hd.depth + 1;
    if (ok) {  // the file must be exactly as long as the header says BEFORE anything is allocated from its counts
        const int64_t want = (int64_t)sizeof(hd) + hd.n_ref * (int64_t)sizeof(RefNode) + hd.n_rank * 4 + hd.n_fast * (int64_t)sizeof(FastNode) +
                             hd.n_wide * (int64_t)sizeof(WideNode) + hd.n_leaf * 4 + hd.n_tris * 8;
        ok = fseek(f, 0, SEEK_END) == 0 && (int64_t)ftell(f) == want && fseek(f, (long)sizeof(hd), SEEK_SET) == 0;
    }
    if (!ok) { fclose(f); return fail(TAKE_E_INVALID, bad); }
    TakeHostBuild *h = nullptr;
    try {
        h = new TakeHostBuild;
        HostBuild &b = h->hb;
        ok = get(f, b.ref.nodes, hd.n_ref) && get(f, b.ref.dfs_rank, hd.n_rank) && get(f, b.fast.nodes, hd.n_fast) &&
             get(f, b.fast.wide, hd.n_wide) && get(f, b.fast.leaf_prims, hd.n_leaf) && get(f, b.tris, hd.n_tris);
        ok = ok && fgetc(f) == EOF;
        ok = ok && payload_hash(b) == hd.payload_hash;
    } catch (const std::bad_alloc &) {
        fclose(f);
        delete h;
        return fail(TAKE_E_NOMEM, std::string("out of host memory loading ") + path);
    }
    fclose(f);
    if (!ok) { delete h; return fail(TAKE_E_INVALID, bad); }
    HostBuild &b = h->hb;
    b.geom_hash = hd.geom_hash; b.num_prims = hd.num_prims; b.ref.root = hd.ref_root; b.fast.depth = hd.depth;
    b.fast.wide_depth = hd.wide_depth; b.fast.sah_cost = hd.sah_cost; b.abs_max = hd.abs_max; b.ms_ref = hd.ms_ref; b.ms_fast = hd.ms_fast;
    *out = h;
    return TAKE_OK;
} catch (...) { return translate_exception(); }

int take_gpu_scene_create_prebuilt(int device, const TakeSceneDesc *d, TakeHostBuild *h, TakeScene **out) try {
    if (!out || !h) return fail(TAKE_E_INVALID, "null argument");
    *out = nullptr;
    if (int rc = validate(d)) return rc;
    const HostBuild &b = h->hb;
    if (b.num_prims != d->num_prims || (int64_t)b.fast.leaf_prims.size() != d->num_prims ||
        b.geom_hash != geometry_hash(d, (int)std::max(1u, std::thread::hardware_concurrency())))
        return fail(TAKE_E_INVALID, "the prebuilt acceleration structures do not belong to this scene");
    if (b.fast.depth > TAKE_STACK_SMEM + TAKE_STACK_LOCAL) return fail(TAKE_E_INVALID, "acceleration tree too deep");
    return scene_create_from(device, d, &h->hb, out);
} catch (...) { return translate_exception(); }

int take_gpu_intersect_device(TakeScene *s, const TakeRay *d_rays, int64_t n, TakeHit *d_hits, int flags) try {
    if (!s || (n > 0 && (!d_rays || !d_hits))) return fail(TAKE_E_INVALID, "null argument");
    if (n < 0) return fail(TAKE_E_INVALID, "negative ray count");
    if (n == 0) return TAKE_OK;
    if (int rc = finish_reference_tree(s)) return rc;
    // the persistent kernels end on a 32-bit work cursor that advances in steps of 32: keep every launch far below 2^32
    const int64_t kChunk = (int64_t)1 << 30;
    if (n > kChunk) {
        for (int64_t off = 0; off < n; off += kChunk)
            if (int rc = take_gpu_intersect_device(s, d_rays + off, std::min(kChunk, n - off), d_hits + off, flags)) return rc;
        return TAKE_OK;
    }
    CU(cudaSetDevice(s->device));
    if (flags == TAKE_ISECT_EXACT) {
        k_intersect_exact<<<(unsigned)((n + 127) / 128), 128, 0, s->stream>>>(s->dev, d_rays, n, d_hits);
    } else if (flags == TAKE_ISECT_FAST) {
        CU(cudaMemsetAsync(s->fetch.p, 0, 4, s->stream));
#if TAKE_EXPERIMENTAL
        if (!s->wide)
            k_intersect_fast<false, false><<<s->blocks_isect, 128, 0, s->stream>>>(s->dev, d_rays, n, d_hits, nullptr, s->fetch.as<uint32_t>());
        else
#endif
            k_intersect_fast<false, true><<<s->blocks_isect, 128, 0, s->stream>>>(s->dev, d_rays, n, d_hits, nullptr, s->fetch.as<uint32_t>());
    } else {
        return fail(TAKE_E_INVALID, "unknown intersect flags");
    }
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(s->stream));
    return TAKE_OK;
} catch (...) { return translate_exception(); }

int take_gpu_intersect(TakeScene *s, const TakeRay *rays, int64_t n, TakeHit *hits, int flags) try {
    if (!s || (n > 0 && (!rays || !hits))) return fail(TAKE_E_INVALID, "null argument");
    if (n < 0) return fail(TAKE_E_INVALID, "negative ray count");
    CU(cudaSetDevice(s->device));
    const int64_t chunk = 1 << 24;
    for (int64_t off = 0; off < n; off += chunk) {
        const int64_t m = std::min(chunk, n - off);
        CU(s->scratch_a.ensure(m * sizeof(TakeRay)));
        CU(s->scratch_b.ensure(m * sizeof(TakeHit)));
        CU(cudaMemcpyAsync(s->scratch_a.p, rays + off, m * sizeof(TakeRay), cudaMemcpyHostToDevice, s->stream));
        if (int rc = take_gpu_intersect_device(s, s->scratch_a.as<TakeRay>(), m, s->scratch_b.as<TakeHit>(), flags)) return rc;
        CU(cudaMemcpyAsync(hits + off, s->scratch_b.p, m * sizeof(TakeHit), cudaMemcpyDeviceToHost, s->stream));
        CU(cudaStreamSynchronize(s->stream));
    }
    return TAKE_OK;
} catch (...) { return translate_exception(); }

int take_gpu_occluded(TakeScene *s, const TakeRay *rays, int64_t n, uint8_t *occluded) try {
    if (!s || (n > 0 && (!rays || !occluded))) return fail(TAKE_E_INVALID, "null argument");
    if (n < 0) return fail(TAKE_E_INVALID, "negative ray count");
    CU(cudaSetDevice(s->device));
    if (int rc = finish_reference_tree(s)) return rc;
    const int64_t chunk = 1 << 24;
    for (int64_t off = 0; off < n; off += chunk) {
        const int64_t m = std::min(chunk, n - off);
        CU(s->scratch_a.ensure(m * sizeof(TakeRay)));
        CU(s->scratch_b.ensure(m));
        CU(cudaMemcpyAsync(s->scratch_a.p, rays + off, m * sizeof(TakeRay), cudaMemcpyHostToDevice, s->stream));
        CU(cudaMemsetAsync(s->fetch.p, 0, 4, s->stream));
#if TAKE_EXPERIMENTAL
        if (!s->wide)
            k_intersect_fast<true, false><<<s->blocks_occl, 128, 0, s->stream>>>(s->dev, s->scratch_a.as<TakeRay>(), m, nullptr,
                                                                                 s->scratch_b.as<uint8_t>(), s->fetch.as<uint32_t>());
        else
#endif
            k_intersect_fast<true, true><<<s->blocks_occl, 128, 0, s->stream>>>(s->dev, s->scratch_a.as<TakeRay>(), m, nullptr,
                                                                                s->scratch_b.as<uint8_t>(), s->fetch.as<uint32_t>());
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(occluded + off, s->scratch_b.p, m, cudaMemcpyDeviceToHost, s->stream));
        CU(cudaStreamSynchronize(s->stream));
    }
    return TAKE_OK;
} catch (...) { return translate_exception(); }

}  // extern "C"

namespace {
// Queue one render call on the scene's streams: everything up to (and including) an event `e1` on s->stream that
// orders all of its work.  No host synchronisation; `d_totals` receives the counters.
// `provisional`: the caller can cope with a render whose tie-break ranks are not in place yet (it checks the tie counter
// afterwards and repeats the call if a rank ever mattered); otherwise the reference-order tree is waited for here.
int render_enqueue(TakeScene *s, const TakeRenderOpts *o, double *d_sum, double *d_sumsq, Totals *d_totals, StageTimer &tm,
                   cudaEvent_t e0, cudaEvent_t e1, int64_t &launches, int64_t &waves, bool provisional = false) {
    if (!provisional) {
        if (int rc = finish_reference_tree(s)) return rc;   // tie-break ranks must be in the leaf records before the first ray
    } else {
        CU(cudaMemsetAsync(s->tie_count.p, 0, 8, s->stream));
    }
    const int64_t npix = (int64_t)s->width * s->height;
    const int64_t spp = o->spp_end - o->spp_begin;
    const int64_t cap_env = std::max<int64_t>(1024, (int64_t)env_int("TAKE_WAVE_SLOTS", 1 << 25));
    const int64_t capacity = std::min<int64_t>(cap_env, std::max<int64_t>(npix * std::max<int64_t>(spp, 1), 1024));
    tm.stream = s->stream;
    const bool count = (o->flags & TAKE_RENDER_COUNT_TESTS) || env_int("TAKE_COUNT_TESTS", 0) != 0;
    // Two waves in flight on two streams: the thin late passes of one wave (few, long rays) overlap the fat early passes
    // of the next, and the latency-bound shade kernels share the SMs with the issue-bound traversal kernels.  Per-kernel
    // timing (TAKE_RENDER_STAGE_TIMES) needs kernels that run alone, so it serialises on one stream.
    // Wave shape: a wave covers `chunk_pixels` pixels x `per_wave` sample indices.  The samples of a pixel that sit in
    // one wave can share warps (TAKE_WARP_SAMPLES, wavefront.cuh: slot_to_local), which makes the camera rays of a warp
    // nearly identical -- worth more than covering the whole frame per wave -- so when the capacity holds fewer than
    // that many samples of every pixel the wave covers fewer pixels instead.
    const int64_t warp_samples = std::max(1, std::min(32, env_int("TAKE_WARP_SAMPLES", TAKE_WARP_SAMPLES_DEFAULT)));
    int64_t group = 1;  // largest power of two <= min(warp_samples, spp)
    while (group * 2 <= std::min<int64_t>(warp_samples, std::max<int64_t>(spp, 1))) group *= 2;
    int64_t chunk_pixels = std::min(npix, capacity & ~int64_t(31));
    if (capacity / chunk_pixels < group) chunk_pixels = std::max<int64_t>(32, (capacity / group) & ~int64_t(31));
    const int64_t per_wave_full = std::max<int64_t>(1, capacity / chunk_pixels);
    const int64_t n_waves_est = ((npix + chunk_pixels - 1) / chunk_pixels) * ((std::max<int64_t>(spp, 1) + per_wave_full - 1) / per_wave_full);
    const int sets = (!tm.on && n_waves_est > 1 && env_int("TAKE_OVERLAP", 1)) ? 2 : 1;
    if (int rc = ensure_wave(s, capacity, sets, o->max_depth + 2)) return rc;
    if (sets == 2 && !s->stream2) {
        CU(cudaStreamCreateWithFlags(&s->stream2, cudaStreamNonBlocking));
        apply_l2_policy(s, s->stream2);
        CU(cudaEventCreateWithFlags(&s->ev_acc[0], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&s->ev_acc[1], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&s->ev_begin, cudaEventDisableTiming));
    }
    cudaStream_t streams[2] = {s->stream, sets == 2 ? s->stream2 : s->stream};
    CU(cudaMemsetAsync(d_totals, 0, sizeof(Totals), s->stream));
    CU(cudaEventRecord(e0, s->stream));
    if (sets == 2) {  // the second stream starts after everything already queued on the first (output buffers, totals)
        CU(cudaEventRecord(s->ev_begin, s->stream));
        CU(cudaStreamWaitEvent(s->stream2, s->ev_begin, 0));
    }
    launches = 0; waves = 0;
    Wave w[2];
    fill_wave_ptrs(s, w[0], o, 0);
    if (sets == 2) fill_wave_ptrs(s, w[1], o, 1);
    w[0].totals = w[1].totals = d_totals;
    w[0].count_ties = w[1].count_ties = provisional ? 1 : 0;
    cudaEvent_t prev_acc = nullptr;
    for (int64_t base = 0; base < npix; base += chunk_pixels) {
        const int64_t cp = std::min(chunk_pixels, npix - base);
        int64_t per_wave = std::max<int64_t>(1, capacity / cp);
        if (per_wave >= group) per_wave -= per_wave % group;  // whole sample groups per wave
        for (int64_t s0 = o->spp_begin; s0 < o->spp_end; s0 += per_wave) {
            const int64_t ns = std::min(per_wave, o->spp_end - s0);
            const int k = sets == 2 ? (int)(waves & 1) : 0;
            Wave &wv = w[k];
            wv.chunk_pixels = (int32_t)cp;
            wv.chunk_base = (int32_t)base;
            wv.sample0 = s0;
            wv.samples_in_wave = (int32_t)ns;
            {   // samples of one pixel that share a warp (wavefront.cuh: slot_to_local): largest power of two <= the
                // request that divides the wave's sample count while 32 / G_s divides its pixel count
                int g = 0;
                while ((2 << g) <= warp_samples && ns % (2 << g) == 0) ++g;
                while (g > 0 && cp % (32 >> g) != 0) --g;
                wv.gs_log2 = g;
            }
            wv.n_slots = (int32_t)(cp * ns);
            cudaEvent_t done = sets == 2 ? s->ev_acc[k] : nullptr;
            if (int rc = launch_wave(s, wv, o, d_sum, d_sumsq, nullptr, tm, count, launches, streams[k], prev_acc, done)) return rc;
            prev_acc = done;
            waves++;
        }
    }
    // join: accumulations are chained and are the last kernel of their wave, so the last one orders everything
    if (sets == 2 && prev_acc) CU(cudaStreamWaitEvent(s->stream, prev_acc, 0));
    CU(cudaEventRecord(e1, s->stream));
    return TAKE_OK;
}
}  // namespace

extern "C" {

int take_gpu_render_device(TakeScene *s, const TakeRenderOpts *o, double *d_sum, double *d_sumsq, TakeStats *stats) try {
    if (int rc = check_opts(s, o)) return rc;
    if (!d_sum) return fail(TAKE_E_INVALID, "null output buffer");
    CU(cudaSetDevice(s->device));
    StageTimer tm;
    tm.on = stats && ((o->flags & TAKE_RENDER_STAGE_TIMES) || env_int("TAKE_STAGE_TIMES", 0));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    int64_t launches = 0, waves = 0;
    CU(s->totals.ensure(sizeof(Totals)));
    // The reference-order tree (the source of the equal-t tie-break ranks) may still be under construction on its host thread
    // (device-built scenes).  Instead of waiting, render now and count the leaf tests in which a rank decided anything: with
    // jittered rays that is almost always zero, and then the image is exactly what it would have been with the ranks in
    // place.  If the count is not zero the outputs are restored from a snapshot and the call is repeated after the join.
    bool provisional = false;
    const size_t out_bytes = (size_t)s->width * s->height * 3 * sizeof(double);
    if (s->ref_pending) {
        if (s->ref_job->done.load() || !env_int("TAKE_PROVISIONAL", 1) || tm.on) {
            if (int rc0 = finish_reference_tree(s)) { cudaEventDestroy(e0); cudaEventDestroy(e1); return rc0; }
        } else {
            provisional = true;
            CU(s->snap_sum.ensure(out_bytes));
            CU(cudaMemcpyAsync(s->snap_sum.p, d_sum, out_bytes, cudaMemcpyDeviceToDevice, s->stream));
            if (d_sumsq) {
                CU(s->snap_sq.ensure(out_bytes));
                CU(cudaMemcpyAsync(s->snap_sq.p, d_sumsq, out_bytes, cudaMemcpyDeviceToDevice, s->stream));
            }
        }
    }
    int rc = render_enqueue(s, o, d_sum, d_sumsq, s->totals.as<Totals>(), tm, e0, e1, launches, waves, provisional);
    if (rc == TAKE_OK) {
        cudaError_t e = cudaStreamSynchronize(s->stream);
        if (e != cudaSuccess) rc = fail(TAKE_E_CUDA, std::string("render: ") + cudaGetErrorString(e));
    }
    if (rc == TAKE_OK && provisional) {
        s->provisional_renders++;
        unsigned long long ties = 0;
        CU(cudaMemcpy(&ties, s->tie_count.p, 8, cudaMemcpyDeviceToHost));
        if (ties != 0) {   // a rank mattered: take the outputs back and do it again with the ranks in place
            s->provisional_reruns++;
            CU(cudaMemcpyAsync(d_sum, s->snap_sum.p, out_bytes, cudaMemcpyDeviceToDevice, s->stream));
            if (d_sumsq) CU(cudaMemcpyAsync(d_sumsq, s->snap_sq.p, out_bytes, cudaMemcpyDeviceToDevice, s->stream));
            rc = finish_reference_tree(s);
            if (rc == TAKE_OK) rc = render_enqueue(s, o, d_sum, d_sumsq, s->totals.as<Totals>(), tm, e0, e1, launches, waves, false);
            if (rc == TAKE_OK) {
                cudaError_t e = cudaStreamSynchronize(s->stream);
                if (e != cudaSuccess) rc = fail(TAKE_E_CUDA, std::string("render: ") + cudaGetErrorString(e));
            }
        }
    }
    float ms = 0;
    if (rc == TAKE_OK) cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (rc != TAKE_OK) return rc;
    tm.stream = s->stream;
    tm.collect();
    read_totals(s, stats, tm, ms, launches, waves);
    return TAKE_OK;
} catch (...) { return translate_exception(); }

// ---- asynchronous render: up to two calls in flight; the device->host copy of call k overlaps the kernels of call k+1 --
int take_gpu_render_async(TakeScene *s, const TakeRenderOpts *o, double *sum_rgb, double *sumsq_rgb, int64_t *ticket) try {
    if (int rc = check_opts(s, o)) return rc;
    if (!sum_rgb || !ticket) return fail(TAKE_E_INVALID, "null argument");
    if ((o->flags & TAKE_RENDER_STAGE_TIMES) || env_int("TAKE_STAGE_TIMES", 0))
        return fail(TAKE_E_INVALID, "per-stage timing serialises the kernels: use take_gpu_render for it");
    CU(cudaSetDevice(s->device));
    TakeScene::AsyncSlot &a = s->async_slot[s->next_ticket & 1];
    if (a.busy) return fail(TAKE_E_INVALID, "two renders already in flight: take_gpu_render_wait the older ticket first");
    const size_t bytes = (size_t)s->width * s->height * 3 * sizeof(double);
    if (!s->copy_stream) CU(cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking));
    if (!a.e0) {
        CU(cudaEventCreate(&a.e0));
        CU(cudaEventCreate(&a.e1));
        CU(cudaEventCreateWithFlags(&a.done, cudaEventDisableTiming));
        CU(cudaMallocHost((void **)&a.h_totals, sizeof(Totals)));
        CU(a.totals.ensure(sizeof(Totals)));
    }
    CU(a.sum.ensure(bytes));
    CU(cudaMemsetAsync(a.sum.p, 0, bytes, s->stream));
    if (sumsq_rgb) {
        CU(a.sq.ensure(bytes));
        CU(cudaMemsetAsync(a.sq.p, 0, bytes, s->stream));
    }
    StageTimer tm;
    if (int rc = render_enqueue(s, o, a.sum.as<double>(), sumsq_rgb ? a.sq.as<double>() : nullptr, a.totals.as<Totals>(), tm, a.e0, a.e1,
                                a.launches, a.waves))
        return rc;
    CU(cudaStreamWaitEvent(s->copy_stream, a.e1, 0));
    CU(cudaMemcpyAsync(sum_rgb, a.sum.p, bytes, cudaMemcpyDeviceToHost, s->copy_stream));
    if (sumsq_rgb) CU(cudaMemcpyAsync(sumsq_rgb, a.sq.p, bytes, cudaMemcpyDeviceToHost, s->copy_stream));
    CU(cudaMemcpyAsync(a.h_totals, a.totals.p, sizeof(Totals), cudaMemcpyDeviceToHost, s->copy_stream));
    CU(cudaEventRecord(a.done, s->copy_stream));
    a.busy = true;
    a.ticket = s->next_ticket;
    *ticket = s->next_ticket++;
    return TAKE_OK;
} catch (...) { return translate_exception(); }

int take_gpu_render_wait(TakeScene *s, int64_t ticket, TakeStats *stats) try {
    if (!s) return fail(TAKE_E_INVALID, "null scene");
    TakeScene::AsyncSlot &a = s->async_slot[ticket & 1];
    if (!a.busy || a.ticket != ticket) return fail(TAKE_E_INVALID, "unknown or already collected ticket");
    CU(cudaSetDevice(s->device));
    a.busy = false;
    CU(cudaEventSynchronize(a.done));
    float ms = 0;
    cudaEventElapsedTime(&ms, a.e0, a.e1);
    StageTimer tm;
    fill_stats(*a.h_totals, stats, tm, ms, a.launches, a.waves);
    return TAKE_OK;
} catch (...) { return translate_exception(); }

int take_gpu_render(TakeScene *s, const TakeRenderOpts *o, double *sum_rgb, double *sumsq_rgb, TakeStats *stats) try {
    if (int rc = check_opts(s, o)) return rc;
    if (!sum_rgb) return fail(TAKE_E_INVALID, "null output buffer");
    CU(cudaSetDevice(s->device));
    const size_t bytes = (size_t)s->width * s->height * 3 * sizeof(double);
    CU(s->scratch_a.ensure(bytes));
    CU(cudaMemsetAsync(s->scratch_a.p, 0, bytes, s->stream));
    if (sumsq_rgb) {
        CU(s->scratch_b.ensure(bytes));
        CU(cudaMemsetAsync(s->scratch_b.p, 0, bytes, s->stream));
    }
    if (int rc = take_gpu_render_device(s, o, s->scratch_a.as<double>(), sumsq_rgb ? s->scratch_b.as<double>() : nullptr, stats))
        return rc;
    CU(cudaMemcpyAsync(sum_rgb, s->scratch_a.p, bytes, cudaMemcpyDeviceToHost, s->stream));
    if (sumsq_rgb) CU(cudaMemcpyAsync(sumsq_rgb, s->scratch_b.p, bytes, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return TAKE_OK;
} catch (...) { return translate_exception(); }

// ---- single-process multi-GPU render: scene replicated, sample ranges sharded, one NCCL sum-reduce ------------------
}  // extern "C"

namespace {
struct Nccl {
    typedef void *comm_t;
    int (*CommInitAll)(comm_t *, int, const int *) = nullptr;
    int (*CommDestroy)(comm_t) = nullptr;
    int (*Reduce)(const void *, void *, size_t, int, int, int, comm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool ok = false;
    Nccl() {
        void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return;
        CommInitAll = (decltype(CommInitAll))dlsym(h, "ncclCommInitAll");
        CommDestroy = (decltype(CommDestroy))dlsym(h, "ncclCommDestroy");
        Reduce = (decltype(Reduce))dlsym(h, "ncclReduce");
        GroupStart = (decltype(GroupStart))dlsym(h, "ncclGroupStart");
        GroupEnd = (decltype(GroupEnd))dlsym(h, "ncclGroupEnd");
        GetErrorString = (decltype(GetErrorString))dlsym(h, "ncclGetErrorString");
        ok = CommInitAll && CommDestroy && Reduce && GroupStart && GroupEnd && GetErrorString;
    }
};
const int kNcclFloat64 = 8, kNcclSum = 0;  // ncclDataType_t / ncclRedOp_t values (nccl.h)
Nccl &g_nccl() {
    static Nccl n;  // libnccl.so.2 is loaded on first use
    return n;
}
}  // namespace

// Persistent multi-GPU handle: everything that does not depend on the render call is done once -- host-side trees built
// once, one replica per device uploaded from its own host thread, ONE communicator over all devices, per-device partial
// accumulation buffers -- so that a render call is only: enqueue every device's share (no host synchronisation in
// between), one grouped ncclReduce onto devices[0], one device->host copy there.
struct TakeMulti {
    int ndev = 0;
    std::vector<int> devices;
    std::vector<TakeScene *> scenes;
    std::vector<double *> d_sum, d_sq;
    std::vector<cudaEvent_t> e0, e1, e_end;
    std::vector<Nccl::comm_t> comms;
    size_t count = 0;
};

extern "C" {

int take_gpu_multi_destroy(TakeMulti *m) {
    if (!m) return TAKE_OK;
    for (int i = 0; i < m->ndev; ++i) {
        cudaSetDevice(m->devices[i]);
        if (i < (int)m->comms.size() && m->comms[i]) g_nccl().CommDestroy(m->comms[i]);
        if (i < (int)m->d_sum.size() && m->d_sum[i]) cudaFree(m->d_sum[i]);
        if (i < (int)m->d_sq.size() && m->d_sq[i]) cudaFree(m->d_sq[i]);
        if (i < (int)m->e0.size() && m->e0[i]) cudaEventDestroy(m->e0[i]);
        if (i < (int)m->e1.size() && m->e1[i]) cudaEventDestroy(m->e1[i]);
        if (i < (int)m->e_end.size() && m->e_end[i]) cudaEventDestroy(m->e_end[i]);
        if (i < (int)m->scenes.size() && m->scenes[i]) take_gpu_scene_destroy(m->scenes[i]);
    }
    delete m;
    return TAKE_OK;
}

int take_gpu_multi_create(int ndev, const int *devices, const TakeSceneDesc *d, TakeMulti **out) try {
    if (!out) return fail(TAKE_E_INVALID, "null argument");
    *out = nullptr;
    if (ndev < 1 || ndev > 64 || !devices) return fail(TAKE_E_INVALID, "bad device list");
    for (int i = 0; i < ndev; ++i)
        for (int j = 0; j < i; ++j)
            if (devices[i] == devices[j]) return fail(TAKE_E_INVALID, "a device is listed twice");
    if (int rc = validate(d)) return rc;
    if (ndev > 1 && !g_nccl().ok) return fail(TAKE_E_CUDA, "libnccl.so.2 could not be loaded (needed to combine the partial images)");
    // every replica builds its fast tree on its own device; the reference-order tree is built ONCE, in the background, and
    // shared (TAKE_DEVICE_BUILD=0: both trees on the host, once, as before)
    const bool device_build = env_int("TAKE_DEVICE_BUILD", 1) && !TAKE_EXPERIMENTAL;
    HostBuild hb;
    std::shared_ptr<RefJob> job;
    if (device_build) job = start_reference_tree(d, host_threads());
    else if (int rc = host_build(d, host_threads(), hb)) return rc;
    TakeMulti *m = new TakeMulti;
    m->ndev = ndev;
    m->devices.assign(devices, devices + ndev);
    m->scenes.assign(ndev, nullptr);
    m->d_sum.assign(ndev, nullptr);
    m->d_sq.assign(ndev, nullptr);
    m->e0.assign(ndev, nullptr);
    m->e1.assign(ndev, nullptr);
    m->e_end.assign(ndev, nullptr);
    m->count = (size_t)d->camera.width * d->camera.height * 3;
    const size_t bytes = m->count * sizeof(double);
    std::vector<int> rcs(ndev, TAKE_OK);
    std::vector<std::string> errs(ndev);
    std::vector<std::thread> pool;
    for (int i = 0; i < ndev; ++i) {
        pool.emplace_back([&, i]() {
            auto body = [&]() -> int {
                if (int rc = scene_create_from(devices[i], d, device_build ? nullptr : &hb, &m->scenes[i], job)) return rc;
                CU(cudaMalloc((void **)&m->d_sum[i], bytes));
                CU(cudaMalloc((void **)&m->d_sq[i], bytes));
                CU(cudaEventCreate(&m->e0[i]));
                CU(cudaEventCreate(&m->e1[i]));
                CU(cudaEventCreate(&m->e_end[i]));
                return TAKE_OK;
            };
            rcs[i] = body();
            if (rcs[i]) errs[i] = g_error;  // g_error is thread-local
        });
    }
    for (auto &t : pool) t.join();
    for (int i = 0; i < ndev; ++i)
        if (rcs[i]) { int rc = rcs[i]; std::string e = errs[i]; take_gpu_multi_destroy(m); return fail(rc, e); }
    if (ndev > 1) {
        m->comms.assign(ndev, nullptr);
        int r = g_nccl().CommInitAll(m->comms.data(), ndev, devices);
        if (r) {
            std::string e = std::string("ncclCommInitAll: ") + g_nccl().GetErrorString(r);
            take_gpu_multi_destroy(m);
            return fail(TAKE_E_CUDA, e);
        }
    }
    *out = m;
    return TAKE_OK;
} catch (...) { return translate_exception(); }

// Device i renders a contiguous share of [spp_begin, spp_end) of every pixel; the partial sums are combined on
// devices[0] with one grouped ncclReduce (the path's one exchange step) and copied to the host from there.
int take_gpu_multi_render(TakeMulti *m, const TakeRenderOpts *o, double *sum_rgb, double *sumsq_rgb, TakeStats *stats) try {
    if (!m || !o || !sum_rgb) return fail(TAKE_E_INVALID, "null argument");
    if (int rc = check_opts(m->scenes[0], o)) return rc;
    const int ndev = m->ndev;
    const size_t bytes = m->count * sizeof(double);
    const int64_t n = std::max<int64_t>(0, o->spp_end - o->spp_begin), base = n / ndev, rem = n % ndev;
    std::vector<int64_t> launches(ndev, 0), waves(ndev, 0);
    for (int i = 0; i < ndev; ++i) {   // enqueue only: nothing here waits for a device
        TakeScene *s = m->scenes[i];
        CU(cudaSetDevice(m->devices[i]));
        CU(cudaMemsetAsync(m->d_sum[i], 0, bytes, s->stream));
        if (sumsq_rgb) CU(cudaMemsetAsync(m->d_sq[i], 0, bytes, s->stream));
        TakeRenderOpts mine = *o;
        mine.spp_begin = o->spp_begin + i * base + std::min<int64_t>(i, rem);
        mine.spp_end = mine.spp_begin + base + (i < rem ? 1 : 0);
        mine.flags &= ~TAKE_RENDER_STAGE_TIMES;
        StageTimer tm;
        CU(s->totals.ensure(sizeof(Totals)));
        if (int rc = render_enqueue(s, &mine, m->d_sum[i], sumsq_rgb ? m->d_sq[i] : nullptr, s->totals.as<Totals>(), tm, m->e0[i],
                                    m->e1[i], launches[i], waves[i]))
            return rc;
    }
    if (ndev > 1) {
        int r = 0;
        for (int pass = 0; pass < (sumsq_rgb ? 2 : 1) && !r; ++pass) {
            g_nccl().GroupStart();
            for (int i = 0; i < ndev && !r; ++i) {
                double *buf = pass == 0 ? m->d_sum[i] : m->d_sq[i];
                r = g_nccl().Reduce(buf, buf, m->count, kNcclFloat64, kNcclSum, 0, m->comms[i], m->scenes[i]->stream);
            }
            const int e = g_nccl().GroupEnd();
            if (!r) r = e;
        }
        if (r) return fail(TAKE_E_CUDA, std::string("ncclReduce: ") + g_nccl().GetErrorString(r));
    }
    for (int i = 0; i < ndev; ++i) {
        CU(cudaSetDevice(m->devices[i]));
        CU(cudaEventRecord(m->e_end[i], m->scenes[i]->stream));
    }
    CU(cudaSetDevice(m->devices[0]));
    CU(cudaMemcpyAsync(sum_rgb, m->d_sum[0], bytes, cudaMemcpyDeviceToHost, m->scenes[0]->stream));
    if (sumsq_rgb) CU(cudaMemcpyAsync(sumsq_rgb, m->d_sq[0], bytes, cudaMemcpyDeviceToHost, m->scenes[0]->stream));
    if (stats) memset(stats, 0, sizeof(*stats));
    for (int i = 0; i < ndev; ++i) {
        CU(cudaSetDevice(m->devices[i]));
        CU(cudaStreamSynchronize(m->scenes[i]->stream));
        if (stats) {
            TakeStats st;
            StageTimer tm;
            float ms = 0;
            cudaEventElapsedTime(&ms, m->e0[i], m->e_end[i]);   // this device's share AND its part in the reduce
            read_totals(m->scenes[i], &st, tm, ms, launches[i], waves[i]);
            stats->samples += st.samples; stats->extend_rays += st.extend_rays; stats->shadow_rays += st.shadow_rays;
            stats->shaded += st.shaded; stats->kernel_launches += st.kernel_launches; stats->waves += st.waves;
            stats->miss_after_light_sample += st.miss_after_light_sample;
            stats->ms_total = std::max(stats->ms_total, st.ms_total);
        }
    }
    return TAKE_OK;
} catch (...) { return translate_exception(); }

// One-shot form: create, render once, destroy.
int take_gpu_render_multi(int ndev, const int *devices, const TakeSceneDesc *d, const TakeRenderOpts *o, double *sum_rgb,
                          double *sumsq_rgb, TakeStats *stats) try {
    if (ndev < 1 || !devices || !o || !sum_rgb) return fail(TAKE_E_INVALID, "bad arguments");
    TakeMulti *m = nullptr;
    if (int rc = take_gpu_multi_create(ndev, devices, d, &m)) return rc;
    const int rc = take_gpu_multi_render(m, o, sum_rgb, sumsq_rgb, stats);
    const std::string e = g_error;
    take_gpu_multi_destroy(m);
    if (rc) g_error = e;
    return rc;
} catch (...) { return translate_exception(); }

// ---- output step: mean, double -> float -> half, B/G/R planes, ZIP pre-filter on the device (exr_out.cuh) -----------
int take_gpu_exr_pack_device(TakeScene *s, const double *d_sum_rgb, int64_t spp, uint8_t *packed) try {
    if (!s || !d_sum_rgb || !packed || spp <= 0) return fail(TAKE_E_INVALID, "take_gpu_exr_pack_device: bad arguments");
    CU(cudaSetDevice(s->device));
    const int64_t bytes = take_gpu_exr_packed_size(s->width, s->height);
    CU(s->exr_packed.ensure((size_t)bytes));
    ExrGeom g;
    g.width = s->width; g.height = s->height;
    g.line_bytes = (int64_t)s->width * 6;
    g.block_bytes = TAKE_EXR_BLOCK_LINES * g.line_bytes;
    const double inv = 1.0 / (double)spp;  // vector.h:194-197: color / Real(spp) multiplies by the reciprocal
    const int blocks = (int)std::min<int64_t>((bytes + 255) / 256, (int64_t)s->sm_count * 16);
    k_exr_pack<<<blocks, 256, 0, s->stream>>>(g, d_sum_rgb, inv, s->exr_packed.as<uint8_t>());
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(packed, s->exr_packed.p, (size_t)bytes, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return TAKE_OK;
} catch (...) { return translate_exception(); }

int take_gpu_exr_pack(TakeScene *s, const double *sum_rgb, int64_t spp, uint8_t *packed) try {
    if (!s || !sum_rgb) return fail(TAKE_E_INVALID, "take_gpu_exr_pack: bad arguments");
    CU(cudaSetDevice(s->device));
    const size_t bytes = (size_t)s->width * s->height * 3 * sizeof(double);
    CU(s->scratch_a.ensure(bytes));
    CU(cudaMemcpyAsync(s->scratch_a.p, sum_rgb, bytes, cudaMemcpyHostToDevice, s->stream));
    return take_gpu_exr_pack_device(s, s->scratch_a.as<double>(), spp, packed);
} catch (...) { return translate_exception(); }

int take_gpu_render_to_exr(TakeScene *s, const TakeRenderOpts *o, const char *path, TakeStats *stats) try {
    if (int rc = check_opts(s, o)) return rc;
    if (!path) return fail(TAKE_E_INVALID, "null path");
    if (o->spp_end <= o->spp_begin) return fail(TAKE_E_INVALID, "empty sample range");
    CU(cudaSetDevice(s->device));
    const size_t bytes = (size_t)s->width * s->height * 3 * sizeof(double);
    CU(s->scratch_a.ensure(bytes));
    CU(cudaMemsetAsync(s->scratch_a.p, 0, bytes, s->stream));
    if (int rc = take_gpu_render_device(s, o, s->scratch_a.as<double>(), nullptr, stats)) return rc;
    std::vector<uint8_t> packed((size_t)take_gpu_exr_packed_size(s->width, s->height));
    if (int rc = take_gpu_exr_pack_device(s, s->scratch_a.as<double>(), o->spp_end - o->spp_begin, packed.data())) return rc;
    return take_gpu_exr_write_packed(path, s->width, s->height, packed.data(), 0);
} catch (...) { return translate_exception(); }

int take_gpu_radiance_samples(TakeScene *s, const TakeRenderOpts *o, int64_t n, const int32_t *px, const int32_t *py,
                              const int64_t *smp, double *rgb) try {
    if (int rc = check_opts(s, o)) return rc;
    if (n < 0 || (n > 0 && (!px || !py || !smp || !rgb))) return fail(TAKE_E_INVALID, "bad sample list");
    if (n == 0) return TAKE_OK;
    CU(cudaSetDevice(s->device));
    if (int rc = finish_reference_tree(s)) return rc;
    const int64_t cap = 1 << 20;
    if (int rc = ensure_wave(s, std::min(n, cap), 1, o->max_depth + 2)) return rc;
    std::vector<int32_t> pixel((size_t)std::min(n, cap));
    StageTimer tm;
    tm.stream = s->stream;
    int64_t launches = 0;
    CU(cudaMemsetAsync(s->totals.p, 0, sizeof(Totals), s->stream));
    for (int64_t off = 0; off < n; off += cap) {
        const int64_t m = std::min(cap, n - off);
        for (int64_t i = 0; i < m; ++i) {
            if (px[off + i] < 0 || px[off + i] >= s->width || py[off + i] < 0 || py[off + i] >= s->height)
                return fail(TAKE_E_INVALID, "pixel out of range");
            pixel[i] = py[off + i] * s->width + px[off + i];
        }
        CU(s->scratch_a.ensure(m * 4));
        CU(s->scratch_b.ensure(m * 8));
        CU(s->scratch_c.ensure(m * 24));
        CU(cudaMemcpyAsync(s->scratch_a.p, pixel.data(), m * 4, cudaMemcpyHostToDevice, s->stream));
        CU(cudaMemcpyAsync(s->scratch_b.p, smp + off, m * 8, cudaMemcpyHostToDevice, s->stream));
        Wave w;
        fill_wave_ptrs(s, w, o);
        w.chunk_pixels = (int32_t)m;
        w.n_slots = (int32_t)m;
        w.samples_in_wave = 1;
        w.list_pixel = s->scratch_a.as<int32_t>();
        w.list_sample = s->scratch_b.as<int64_t>();
        if (int rc = launch_wave(s, w, o, nullptr, nullptr, s->scratch_c.as<double>(), tm, false, launches, s->stream)) return rc;
        CU(cudaMemcpyAsync(rgb + 3 * off, s->scratch_c.p, m * 24, cudaMemcpyDeviceToHost, s->stream));
        CU(cudaStreamSynchronize(s->stream));
    }
    return TAKE_OK;
} catch (...) { return translate_exception(); }

}  // extern "C"
