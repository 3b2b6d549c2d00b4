/* take_gpu.h -- C ABI of the B200 rendering core that replaces TaKe's CPU hot path.
 *
 * The reference (TaKeTube/TaKe) has no plugin / FFI seam: `render()` (src/render.cpp:9-87) parses the
 * scene, builds the BVH and runs the per-pixel path-tracing loop in one statically linked C++ function.
 * This library replaces everything `render()` does AFTER `parse_scene` (src/render.cpp:28): BVH
 * construction, ray-scene intersection and the Monte-Carlo path-integration loop.  A maintainer keeps
 * the reference's front end (Mitsuba XML / PLY / OBJ / textures), flattens the parsed `Scene`
 * (src/scene.h:13-33) into a `TakeSceneDesc`, and calls the functions below (INTEGRATION.md shows the
 * ~80-line adapter).
 *
 * Conventions: plain C, no C++ types, no exceptions across the boundary.  Every function returns
 * 0 on success or a negative TAKE_E_* code, with a human-readable message available from
 * take_gpu_last_error() (thread-local).  The caller owns every host buffer; the library owns all
 * device memory behind the opaque TakeScene handle.  A handle is bound to one CUDA device and is
 * not thread-safe; different handles may be used from different host threads.  All calls are
 * complete when they return.  There is NO CPU fallback: without a usable CUDA device every call
 * fails with TAKE_E_CUDA.
 *
 * All reference arithmetic is IEEE double (src/take.h:27) and so are the data here.
 */
#ifndef TAKE_GPU_H
#define TAKE_GPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TAKE_OK 0
#define TAKE_E_INVALID (-1) /* bad argument / inconsistent scene description            */
#define TAKE_E_CUDA (-2)    /* CUDA runtime error or no device (there is no CPU path)    */
#define TAKE_E_NOMEM (-3)   /* host or device allocation failed                          */

/* Material::index() of the reference's std::variant (src/material.h:82-93). */
enum {
    TAKE_MAT_DIFFUSE = 0,
    TAKE_MAT_MIRROR = 1,
    TAKE_MAT_PLASTIC = 2,
    TAKE_MAT_PHONG = 3,
    TAKE_MAT_BLINN_PHONG = 4,
    TAKE_MAT_BLINN_MICROFACET = 5,
    TAKE_MAT_DISNEY_DIFFUSE = 6,
    TAKE_MAT_DISNEY_METAL = 7,     /* 7,8,10,11 evaluate as Lambertian in the reference   */
    TAKE_MAT_DISNEY_GLASS = 8,     /* (src/materials/disney_metal.inl:1-28 and siblings)  */
    TAKE_MAT_DISNEY_CLEARCOAT = 9, /* cosine sampling, eval == 0 (disney_clearcoat.inl:22-27) */
    TAKE_MAT_DISNEY_SHEEN = 10,
    TAKE_MAT_DISNEY_BSDF = 11,
    /* EXTENSION (no counterpart in the reference, whose only microfacet model is BlinnPhongMicrofacet): GGX /
     * Trowbridge-Reitz distribution with the separable Smith shadowing term and Schlick Fresnel, written in the
     * structure of src/materials/blinn_phong_microfacet.inl.  p[0] = roughness alpha.  Parity unpinned. */
    TAKE_MAT_GGX = 12
};

enum { TAKE_LIGHT_POINT = 0, TAKE_LIGHT_AREA = 1 }; /* src/light.h:9-19 */

/* prim_flags bits */
enum { TAKE_PRIM_HAS_NORMALS = 1, TAKE_PRIM_HAS_UVS = 2, TAKE_PRIM_SPHERE = 4 };

/* Which of the reference's integrators to run (src/integrator/path_tracing.h). */
enum {
    TAKE_INTEGRATOR_MIS = 0,           /* path_tracing                :5-111  (what render.cpp:76 calls) */
    TAKE_INTEGRATOR_RAW = 1,           /* path_tracing_raw            :114-157 */
    TAKE_INTEGRATOR_ONE_SAMPLE_MIS = 2, /* path_tracing_one_sample_MIS :161-271 */
    /* path_tracing_one_sample_MIS_power :274-380: one-sample MIS with the light picked in proportion to its power
     * (luminance x area x pi, src/light.cpp:9-30) instead of uniformly -- what a scene with hundreds of unequal emitters
     * wants (BASELINE config 4).  Dead code in the reference as shipped: nothing fills Scene::lights_power_cdf / _pmf.
     * The tables are built here from the reference's own light_power() formula in the layout its readers expect (N + 1
     * running sums from 0 to 1), and the integrator is pinned against the reference's own function driven with tables built
     * the same way (oracle/ref_harness.cpp).  Not available on scenes with a sampled environment map. */
    TAKE_INTEGRATOR_ONE_SAMPLE_MIS_POWER = 3
};

/* take_gpu_intersect flags */
enum {
    TAKE_ISECT_FAST = 0, /* SAH tree, conservative FP32 boxes, FP64 leaf test; ties -> larger reference DFS rank */
    TAKE_ISECT_EXACT = 1 /* the reference's own tree topology, FP64 slab test and tmax updates (bvh.cpp:86-109) */
};

typedef struct TakeCamera { /* src/camera.h:5-11 */
    int32_t width, height;
    double lookfrom[3], lookat[3], up[3];
    double vfov; /* vertical field of view in degrees (already converted, parse_scene.cpp:366-377) */
} TakeCamera;

typedef struct TakeTextureDesc { /* one Image3 of the TexturePool (src/texture.h:8-14, src/image.h:13-39) */
    int32_t width, height;
    const double *rgb; /* height*width*3, row-major, row 0 first */
} TakeTextureDesc;

typedef struct TakeMaterialDesc { /* POD image of one `Material` alternative (src/material.h:7-80) */
    int32_t type;                 /* TAKE_MAT_*                                                          */
    int32_t tex_id;               /* >= 0: ImageTexture into textures[]; -1: ConstTexture `color`       */
    double color[3];              /* ConstTexture::value (src/texture.h:22-25)                          */
    double uscale, vscale, uoffset, voffset; /* ImageTexture (src/texture.h:16-20)                      */
    double p[2];                  /* eta (mirror, plastic) | exponent (phong, blinn*) | roughness, subsurface (disney diffuse) */
} TakeMaterialDesc;

typedef struct TakeLightDesc { /* src/light.h:9-19 */
    int32_t kind;              /* TAKE_LIGHT_*                                             */
    int32_t prim_id;           /* DiffuseAreaLight::shape_id (primitive id), -1 for points */
    double intensity[3];
    double position[3];        /* PointLight only */
} TakeLightDesc;

/* Flattened `Scene` (src/scene.h:13-33).  Primitive id == index into the reference's scene.shapes:
 * one entry per mesh face in parse order (src/parse/parse_scene.cpp:937-945) or per sphere. */
typedef struct TakeSceneDesc {
    TakeCamera camera;
    double background[3];          /* Scene::background_color */
    int64_t num_vertices;          /* all meshes' vertex arrays concatenated (src/shape.h:13-18) */
    const double *positions;       /* 3*num_vertices */
    const double *normals;         /* 3*num_vertices (zeros where the mesh has none) */
    const double *uvs;             /* 2*num_vertices (zeros where the mesh has none) */
    int64_t num_prims;
    const int32_t *indices;        /* 3*num_prims global vertex indices; for a sphere: {sphere index, 0, 0} */
    const int32_t *prim_material;  /* material id the hit reports (mesh.material_id, src/shape.cpp:85)     */
    const int32_t *prim_light;     /* area_light_id or -1 (src/shape.h:8-11)                               */
    const uint8_t *prim_flags;     /* TAKE_PRIM_* */
    int64_t num_spheres;
    const double *spheres;         /* 4*num_spheres: center xyz, radius (src/shape.h:20-23) */
    int32_t num_materials, num_textures, num_lights, reserved;
    const TakeMaterialDesc *materials;
    const TakeTextureDesc *textures;
    const TakeLightDesc *lights;
    /* EXTENSION (no counterpart in the reference, whose only "environment" is the constant background_color):
     * a lat-long environment map.  Row 0 is +y (up); a direction d maps to u = (atan2(-d.z, d.x) + pi) / 2pi (the phi of
     * get_sphere_uv, src/shape.cpp:3-11) and v = acos(d.y) / pi; lookups are piecewise constant.  With env_rgb == NULL
     * nothing changes.  env_sample == 0: the map only replaces background_color on misses (a constant map then gives
     * exactly the reference's <background> render).  env_sample == 1: the map additionally takes part in light
     * sampling as light number num_lights (uniform pick over num_lights + 1), importance sampled through marginal /
     * conditional CDF tables of luminance x sin(theta), with MIS weights on BSDF-sampled misses. */
    int32_t env_width, env_height;
    int32_t env_sample, reserved2;
    const double *env_rgb; /* env_height * env_width * 3 */
} TakeSceneDesc;

/* A ray exactly as the reference's `Ray` (src/ray.h:4-9): 8 doubles. */
typedef struct TakeRay {
    double origin[3];
    double dir[3];
    double tmin, tmax;
} TakeRay;

/* Closest hit: what scene_intersect() (src/scene.cpp:25-47) determines, plus the primitive id that
 * `Intersection` lacks (src/intersection.h:4-12). */
typedef struct TakeHit {
    int32_t prim_id; /* -1 = miss */
    int32_t pad;
    double t;        /* bit-identical to the reference's Intersection::t */
    double u, v;     /* Moller-Trumbore barycentrics (src/shape.cpp:63,69); 0 for spheres */
} TakeHit;

typedef struct TakeRenderOpts {
    int32_t integrator;   /* TAKE_INTEGRATOR_* */
    int32_t max_depth;    /* `-max_depth` (src/render.cpp:14-23); the loop runs max_depth+1 bounces */
    int64_t spp_begin;    /* sample-index range [spp_begin, spp_end) of EVERY pixel: the unit of   */
    int64_t spp_end;      /* multi-GPU sharding.  Results depend only on (seed, pixel, index).     */
    uint64_t seed;
    int32_t flags;        /* TAKE_RENDER_* */
    int32_t reserved;     /* with TAKE_RENDER_RUSSIAN_ROULETTE: first loop iteration that plays (0 = the default, 3); else 0 */
} TakeRenderOpts;

enum {
    TAKE_RENDER_DEFAULT = 0,
    TAKE_RENDER_NO_SORT = 1,     /* disable the per-bounce material sort (A/B measurements)                    */
    TAKE_RENDER_STAGE_TIMES = 2, /* bracket every kernel with CUDA events and fill TakeStats::ms_<stage>       */
    TAKE_RENDER_COUNT_TESTS = 4, /* run the instrumented traversal kernels and fill box_tests / tri_tests      */
    /* EXTENSION (the reference's README.md:19-24 lists Russian roulette as a goal; its integrators have none): from loop
     * iteration `reserved` (default 3) on, a path survives with probability q = min(max component of its throughput, 0.95)
     * and its throughput is divided by q -- unbiased, one extra random draw per iteration.  Off by default: without the flag
     * every result is the reference's.  Parity unpinned (checked against our own CPU restatement and for unbiasedness). */
    TAKE_RENDER_RUSSIAN_ROULETTE = 8
};

typedef struct TakeStats {
    int64_t samples;        /* path samples started                               */
    int64_t extend_rays;    /* closest-hit rays traced (primary + bounce)         */
    int64_t shadow_rays;    /* any-hit rays traced                                */
    int64_t shaded;         /* path vertices shaded                               */
    int64_t box_tests;      /* extend kernel: ray-box tests   (only with TAKE_RENDER_COUNT_TESTS) */
    int64_t tri_tests;      /* extend kernel: leaf tests      (only with TAKE_RENDER_COUNT_TESTS) */
    int64_t kernel_launches;
    double ms_total;        /* device time of the whole call (CUDA events)        */
    double ms_generate, ms_extend, ms_shade, ms_shadow, ms_sort, ms_other; /* only with TAKE_RENDER_STAGE_TIMES */
    int64_t shadow_box_tests, shadow_tri_tests; /* shadow kernel (only with TAKE_RENDER_COUNT_TESTS) */
    int64_t miss_after_light_sample; /* one-sample MIS: light-aimed rays that missed everything (reference: UB) */
    int64_t waves;                   /* generate..accumulate rounds; every pass of a wave launches one extend kernel */
} TakeStats;

typedef struct TakeScene TakeScene; /* opaque */

/* Number of usable CUDA devices. */
int take_gpu_device_count(int *count);

/* Build the acceleration structures (replaces build_bvh / construct_bvh, src/scene.cpp:4-23, src/bvh.cpp:8-45) and upload
 * the scene to `device` once.  The traversal tree is built ON THE DEVICE (PLOC over Morton-sorted primitives, collapsed to
 * 4-wide nodes: take_b200/csrc/bvh_device.cuh); the reference's own tree -- needed for its equal-t tie order and for
 * TAKE_ISECT_EXACT -- is built by a background host thread from a private copy of the primitive boxes, so this call returns
 * without waiting for it and the first call that traces rays joins it.  The caller's arrays are not read after the call
 * returns.  TAKE_DEVICE_BUILD=0 (environment) selects the host's binned-SAH builder instead. */
int take_gpu_scene_create(int device, const TakeSceneDesc *desc, TakeScene **out);
int take_gpu_scene_destroy(TakeScene *scene);
/* Device memory freed by this library (destroyed scenes, builder scratch, wave buffers) stays cached in the device's
 * stream-ordered memory pool, so that the next scene of the process does not pay the driver's allocation cost again (about
 * 0.1 s per GB).  This call returns the cached memory to the driver.  TAKE_MEMPOOL=0 (environment) disables the caching. */
int take_gpu_release_cached_memory(int device);

/* Closest hit for `n` host rays (replaces scene_intersect, src/scene.cpp:25-47).  `flags` = TAKE_ISECT_*. */
int take_gpu_intersect(TakeScene *scene, const TakeRay *rays, int64_t n, TakeHit *hits, int flags);
/* Boolean occlusion for `n` host rays (replaces scene_occluded, src/scene.cpp:49-64). */
int take_gpu_occluded(TakeScene *scene, const TakeRay *rays, int64_t n, uint8_t *occluded);
/* Same, with rays / results already resident on the scene's device (no copies; async on the scene's stream
 * followed by a stream synchronise). */
int take_gpu_intersect_device(TakeScene *scene, const TakeRay *d_rays, int64_t n, TakeHit *d_hits, int flags);

/* Schedule switches (environment variables, read per call; for A/B measurements and for the parity tests, which compare the
 * schedules against each other -- NONE of them changes a result):
 *   TAKE_PACKET=0        camera rays one per thread instead of one packet per warp (trace_packet4)
 *   TAKE_REFILL=0..3     bit 0: bounce passes, bit 1: shadow passes traced with lane refill (trace_refill4); default 3
 *   TAKE_ORDERED_SORT=0  the material sort keeps the order in which rays finished instead of queue order
 *   TAKE_NO_MISS_FAST=1  camera rays that leave the scene go through sort and shade like any other
 *   TAKE_OVERLAP=0       one wave at a time instead of two in flight;  TAKE_WAVE_SLOTS=n  path samples per wave
 *   TAKE_PROVISIONAL=0   renders wait for the reference-order tree instead of running ahead of it
 *   TAKE_DEVICE_BUILD=0  host tree builders;  TAKE_HOST_THREADS=n;  TAKE_MEMPOOL=0;  TAKE_TIMING=1 (creation phases to stderr) */

/* The render loop (replaces the tile lambda of src/render.cpp:59-82 and the integrators of
 * src/integrator/path_tracing.h).  Adds, for every pixel, the radiance of samples
 * [spp_begin, spp_end) to sum_rgb and its square to sumsq_rgb (may be NULL): height*width*3 doubles in image
 * layout (row 0 = top, the layout src/render.cpp:78 writes).  Host buffers are overwritten, not accumulated. */
int take_gpu_render(TakeScene *scene, const TakeRenderOpts *opts, double *sum_rgb, double *sumsq_rgb, TakeStats *stats);
/* Same with DEVICE output buffers, which are accumulated into (zero them first). */
int take_gpu_render_device(TakeScene *scene, const TakeRenderOpts *opts, double *d_sum_rgb, double *d_sumsq_rgb,
                           TakeStats *stats);

/* Asynchronous take_gpu_render: queues the render and the device->host copies of its results and returns a ticket;
 * take_gpu_render_wait blocks until the host buffers of that ticket are complete and fills `stats`.  Up to TWO tickets
 * may be in flight per scene, and they complete in order: the copy of call k then overlaps the kernels of call k+1
 * (a host that consumes per-range results -- progressive display, checkpoints, the bench's per-step read-back -- hides
 * the PCIe transfer completely).  The host buffers must stay valid until the wait returns and should be page-locked
 * (cudaHostAlloc / cudaHostRegister); with pageable memory the call still works but the copy is staged synchronously.
 * TAKE_RENDER_STAGE_TIMES is not available here (it serialises the kernels). */
int take_gpu_render_async(TakeScene *scene, const TakeRenderOpts *opts, double *sum_rgb, double *sumsq_rgb, int64_t *ticket);
int take_gpu_render_wait(TakeScene *scene, int64_t ticket, TakeStats *stats);

/* Single-process multi-GPU render (what parallel_for over tiles, src/parallel.cpp:183-237, becomes across GPUs): the
 * host-side acceleration structures are built once, the scene is replicated on devices[0..ndev), device i renders a
 * contiguous share of the sample-index range [spp_begin, spp_end) of every pixel, and the partial sums are combined on
 * devices[0] with one NCCL sum-reduce (libnccl.so.2 is loaded on first use).  Because a sample's random stream depends
 * only on (seed, pixel, sample index) the result equals the single-GPU render up to summation order.  Host outputs as
 * take_gpu_render. */
int take_gpu_render_multi(int ndev, const int *devices, const TakeSceneDesc *desc, const TakeRenderOpts *opts, double *sum_rgb,
                          double *sumsq_rgb, TakeStats *stats);
/* The same as a persistent handle, for hosts that render more than once (progressive refinement, spp ranges, animation
 * of the camera-independent parts): take_gpu_multi_create builds the host-side structures ONCE, uploads one replica per
 * device (one host thread per device) and creates ONE communicator over all of them; every take_gpu_multi_render then
 * only enqueues each device's share of the sample range (no host synchronisation in between), issues one grouped
 * ncclReduce onto devices[0] and one device->host copy from there.  stats->ms_total is the slowest device's time from
 * its first kernel to the end of its part in the reduce.  Not thread-safe; one handle per set of devices. */
typedef struct TakeMulti TakeMulti; /* opaque */
int take_gpu_multi_create(int ndev, const int *devices, const TakeSceneDesc *desc, TakeMulti **out);
int take_gpu_multi_render(TakeMulti *multi, const TakeRenderOpts *opts, double *sum_rgb, double *sumsq_rgb, TakeStats *stats);
int take_gpu_multi_destroy(TakeMulti *multi);

/* Output step (replaces the .exr branch of imwrite, src/image.cpp:157-175, and the tinyexr code under it).
 * Device half: mean = sum * (1/spp) (src/render.cpp:78 via vector.h:194-197), double -> float (image.cpp:159-161),
 * float -> half with tinyexr's float_to_half_full rounding, channel-planar B,G,R scanlines in 16-line blocks, and the
 * ZIP pre-filter of CompressZip (byte de-interleave + delta predictor): `packed` (HOST, take_gpu_exr_packed_size
 * bytes) receives every block exactly as the reference hands it to deflate.  `d_sum_rgb` is a DEVICE buffer in the
 * layout take_gpu_render_device accumulates into; take_gpu_exr_pack takes the same sums from the HOST. */
int64_t take_gpu_exr_packed_size(int32_t width, int32_t height);
int take_gpu_exr_pack_device(TakeScene *scene, const double *d_sum_rgb, int64_t spp, uint8_t *packed);
int take_gpu_exr_pack(TakeScene *scene, const double *sum_rgb, int64_t spp, uint8_t *packed);
/* Host half: deflate the blocks on `threads` host threads (0 = all) and write a scanline OpenEXR file with the
 * reference writer's attributes (HALF B,G,R, ZIP, increasing Y).  Needs no CUDA device. */
int take_gpu_exr_write_packed(const char *path, int32_t width, int32_t height, const uint8_t *packed, int32_t threads);
/* render() + imwrite("image.exr") in one call (src/main.cpp:21-24): the FP64 sums never leave the device, the
 * device->host transfer is 6 bytes per pixel. */
int take_gpu_render_to_exr(TakeScene *scene, const TakeRenderOpts *opts, const char *path, TakeStats *stats);

/* Radiance of `n` individual path samples (pixel x, image row y from the top, sample index s): 3 doubles each. */
int take_gpu_radiance_samples(TakeScene *scene, const TakeRenderOpts *opts, int64_t n, const int32_t *px,
                              const int32_t *py, const int64_t *s, double *rgb);

/* Raw CUDA stream the scene's work is issued on (a cudaStream_t), for callers that time with events. */
void *take_gpu_scene_stream(TakeScene *scene);

/* Diagnostics: out[0..5] = reference-order tree build ms, fast tree build ms, fast tree depth, SAH cost,
 * number of fast-tree nodes, SM count of the device. */
int take_gpu_scene_info(TakeScene *scene, double *out);

/* out[0..7] = milliseconds of take_gpu_scene_create: validation, scene upload, primitive boxes for the reference-order tree
 * (host), fast-tree build on the device (0 for host-built scenes), shading / light records, the whole call; then 1 if the
 * fast tree was built on the device, 1 if the reference-order tree is still being built in the background (it is joined by
 * the first call that traces rays: take_gpu_scene_create itself does not wait for it). */
int take_gpu_scene_create_timings(TakeScene *scene, double *out8);
/* take_gpu_render / take_gpu_render_device do not wait for that background tree either: its only product the render needs
 * are the ranks that decide EQUAL-t ties, so they render at once, count the leaf tests in which a rank decided anything
 * (practically none with jittered rays), and only if that count is not zero restore the outputs and repeat the call after
 * joining the tree -- the result is always the one with the ranks in place.  out[0] = renders that ran ahead of the tree,
 * out[1] = how many had to be repeated.  (TAKE_PROVISIONAL=0 in the environment makes every render wait instead.) */
int take_gpu_scene_provisional_stats(TakeScene *scene, int64_t *out2);
/* Copies the fast tree out of the device for inspection: 128-byte 4-wide nodes (take_b200/csrc/bvh_build.h: WideNode) and
 * the primitive id of every leaf slot (device-built scenes only; pass NULL otherwise).  Returns the number of wide nodes
 * (either pointer may be NULL). */
int64_t take_gpu_scene_debug_tree(TakeScene *scene, void *wide_nodes, int32_t *leaf_prims);

/* Host-only diagnostics: build the acceleration structures of the HOST builders (the ones take_gpu_scene_create_prebuilt
 * and TAKE_DEVICE_BUILD=0 upload; by default take_gpu_scene_create builds its fast tree on the device), without
 * touching CUDA, and copy them out for inspection (tests/test_bvh_host.py).  Layouts are documented in
 * take_b200/csrc/bvh_build.h. */
typedef struct TakeHostBuild TakeHostBuild;
int take_gpu_host_build(const TakeSceneDesc *desc, TakeHostBuild **out);
int take_gpu_host_build_info(TakeHostBuild *h, double *out8);
int take_gpu_host_build_copy(TakeHostBuild *h, void *ref_nodes, int32_t *dfs_rank, void *fast_nodes, int32_t *leaf_prims,
                             double *leaf_records);
int64_t take_gpu_host_build_wide(TakeHostBuild *h, void *wide_nodes); /* returns the node count; copies if non-NULL */
int take_gpu_host_build_free(TakeHostBuild *h);

/* Build once, create many: with one process per GPU (torchrun, MPI) every rank would otherwise run the host builders
 * on the same cores at the same time.  One rank builds (take_gpu_host_build) and saves; the others load the file --
 * /dev/shm is the natural place on one node -- and all create their scene from the prebuilt structures.  A build
 * carries a hash of the primitive data it was made from; take_gpu_scene_create_prebuilt refuses a build that does not
 * belong to `desc` (TAKE_E_INVALID).  The handle stays the caller's (free it with take_gpu_host_build_free).  The file
 * is a private cache format tied to this library version, not an exchange format. */
int take_gpu_host_build_save(TakeHostBuild *h, const char *path);
int take_gpu_host_build_load(const char *path, TakeHostBuild **out);
int take_gpu_scene_create_prebuilt(int device, const TakeSceneDesc *desc, TakeHostBuild *h, TakeScene **out);

/* Scene-description builder: big meshes straight into the flat arrays of TakeSceneDesc, on all host threads.
 * Replaces, for the shapes routed through it, parse_ply (src/parse/parse_ply.cpp:16-120: tinyply buffers copied element by
 * element), compute_normals (src/compute_normals.cpp:12-47) and the per-face expansion into one std::variant<Sphere,
 * Triangle> plus one light per emissive face (src/parse/parse_scene.cpp:934-945).  Shapes must be added in the order the
 * scene file lists them: primitive ids, light ids and vertex offsets are assigned in call order, exactly as the reference's
 * parser assigns them, and every value is computed with the reference's FP64 operations in its order -- the arrays are
 * bit-identical to flattening the reference parser's Scene (tests/test_mesh_load.py).  Host-only: needs no CUDA device.
 * Matrices are 16 doubles, row-major (Matrix4x4::data, src/matrix.h); NULL = identity.  `inv_to_world` is the reference's
 * own inverse(to_world) (parse_ply.cpp:72), passed in so that the host's matrix code stays the single source of it.
 * `radiance` = the shape's <emitter type="area"> radiance (3 doubles) or NULL.  Only triangle faces are accepted (the
 * reference reads any face list as if it had 3 entries, parse_ply.cpp:83-120; a quad there is silently garbage). */
typedef struct TakeDescBuilder TakeDescBuilder; /* opaque */
int take_gpu_builder_create(TakeDescBuilder **out);
int take_gpu_builder_destroy(TakeDescBuilder *b);
/* <shape type="ply">: ascii, binary_little_endian or binary_big_endian; x,y,z (float / double), optional nx,ny,nz and u,v,
 * `list <any int> <any int> vertex_indices`.  face_normals: parse_scene.cpp:826-834 (drop the normals; else compute
 * angle-weighted vertex normals when the file has none). */
int take_gpu_builder_add_ply(TakeDescBuilder *b, const char *path, const double *to_world, const double *inv_to_world,
                             int32_t material_id, int32_t face_normals, const double *radiance);
/* An already parsed TriangleMesh in world space (rectangle / obj / serialized shapes parsed by the reference's own code):
 * positions 3*nv, normals 3*nv or NULL, uvs 2*nv or NULL, indices 3*nf mesh-local. */
int take_gpu_builder_add_mesh(TakeDescBuilder *b, int64_t num_vertices, const double *positions, const double *normals,
                              const double *uvs, int64_t num_faces, const int32_t *indices, int32_t material_id,
                              int32_t compute_missing_normals, const double *radiance);
int take_gpu_builder_add_sphere(TakeDescBuilder *b, const double *center, double radius, int32_t material_id, const double *radiance);
int take_gpu_builder_add_point_light(TakeDescBuilder *b, const double *intensity, const double *position);
/* Fills the geometry and light fields of `desc` (counts + pointers into the builder: valid until the next add / destroy);
 * camera, background, materials, textures and environment fields stay the caller's. */
int take_gpu_builder_finish(TakeDescBuilder *b, TakeSceneDesc *desc);
/* out[0..3] = file read, conversion, vertex normals, append: milliseconds of the last take_gpu_builder_add_ply */
int take_gpu_builder_timings(TakeDescBuilder *b, double *out4);
/* Writes `desc` as a TAKESCN1 file (take_b200/sceneio.py: the flat arrays, little endian); `spp` = Scene::options.spp.
 * Host-only. */
int take_gpu_scene_desc_save(const TakeSceneDesc *desc, int64_t spp, const char *path);

/* Self-test: the multi-threaded twin of std::sort used by the reference-order builder must return std::sort's exact
 * permutation, ties included.  Returns the number of differing positions (0 = identical).  pattern: 0 random, 1 ascending,
 * 2 descending, 3 all equal, 4 organ pipe; distinct: number of different key values (0 = all different). */
int64_t take_gpu_selftest_sort(int64_t n, int64_t distinct, int32_t pattern, int32_t threads, uint64_t seed);

const char *take_gpu_last_error(void);
const char *take_gpu_version(void);

#ifdef __cplusplus
}
#endif
#endif /* TAKE_GPU_H */
