// TEST INFRASTRUCTURE -- not part of the shipped product.  Only tests/,
// __graft_entry__.smoke() and bench.py's CPU-baseline legs may load the library
// built from this file.
//
// C-ABI harness around the UNMODIFIED reference (TaKeTube/TaKe), compiled by
// oracle/Makefile from the sources where they lie under /root/reference/src into
// oracle/_ref/libtake_ref.so.  Everything here is a thin driver: the arithmetic
// is the reference's own (parse_scene, build_bvh, intersect(BBox,Ray),
// intersect_shape, scene_intersect, scene_occluded, path_tracing*).
//
// What it adds, and why:
//  * ref_intersect: `Intersection` carries no primitive id (src/intersection.h:4-12),
//    so the recursion of src/bvh.cpp:86-109 is restated around the reference's own
//    box / shape tests carrying node.primitive_id, and cross-checked bitwise
//    against scene_intersect().
//  * ref_render / ref_radiance_samples: render() seeds from std::random_device
//    (src/render.cpp:60), so the pixel loop of src/render.cpp:65-78 is restated
//    with a reproducible counter-based stream (oracle/take_rng.h) pre-loaded
//    into a real std::mt19937 that is then passed to the reference integrators.
//  * ref_scene_dump: flattens `Scene` (src/scene.h:13-33) to the TAKESCN1 layout
//    the product's C-ABI consumes (include/take_gpu.h).
#include "scene.h"
#include "parse/parse_scene.h"
#include "take_rng.h"

#include <atomic>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

// Non-inline functions defined in src/integrator/path_tracing.h and emitted in
// render.o; declared (not re-included) to avoid duplicate definitions.
Vector3 path_tracing(const Scene &scene, const Ray &ray, std::mt19937 &rng);
Vector3 path_tracing_raw(const Scene &scene, const Ray &ray, std::mt19937 &rng);
Vector3 path_tracing_one_sample_MIS(const Scene &scene, const Ray &ray, std::mt19937 &rng);
Vector3 path_tracing_one_sample_MIS_power(const Scene &scene, const Ray &ray, std::mt19937 &rng);

namespace {

thread_local std::string g_error;

struct RefScene {
    Scene scene;
};

struct Hit {
    int prim = -1;
    Real t = 0;
    std::optional<Intersection> isect;
};

// src/bvh.cpp:86-109 with the primitive id carried along.
Hit shim(const Scene &sc, int node_id, Ray ray) {
    const BVHNode &node = sc.bvh_nodes[node_id];
    if (node.primitive_id != -1) {
        Hit h;
        h.isect = intersect_shape(sc.shapes[node.primitive_id], sc.meshes, ray);
        if (h.isect) {
            h.prim = node.primitive_id;
            h.t = h.isect->t;
        }
        return h;
    }
    Hit left;
    if (intersect(sc.bvh_nodes[node.left_node_id].box, ray)) {
        left = shim(sc, node.left_node_id, ray);
        if (left.prim != -1) ray.tmax = left.t;
    }
    if (intersect(sc.bvh_nodes[node.right_node_id].box, ray)) {
        Hit right = shim(sc, node.right_node_id, ray);
        if (right.prim != -1) return right;
    }
    return left;
}

inline Ray make_ray(const double *r) {
    return Ray{Vector3{r[0], r[1], r[2]}, Vector3{r[3], r[4], r[5]}, r[6], r[7]};
}

inline uint32_t untemper(uint32_t y) {
    // inverse of mt19937's output tempering (u=11,d=ffffffff,s=7,b=9d2c5680,t=15,c=efc60000,l=18)
    y ^= y >> 18;
    y ^= (y << 15) & 0xEFC60000u;
    uint32_t x = y;
    for (int i = 0; i < 5; ++i) x = y ^ ((x << 7) & 0x9D2C5680u);
    y = x;
    x = y;
    for (int i = 0; i < 3; ++i) x = y ^ (x >> 11);
    return x;
}

// libstdc++ layout: result_type _M_x[624]; size_t _M_p;  (checked at load time)
struct MtImage {
    std::mt19937::result_type x[624];
    size_t p;
};
static_assert(sizeof(MtImage) == sizeof(std::mt19937), "unexpected std::mt19937 layout");

// An mt19937 whose next `nwords` outputs are the sample's stream words.
inline void load_stream(std::mt19937 &rng, uint64_t seed, uint32_t pixel, uint64_t sample, int nwords) {
    MtImage img;
    std::memset(&img, 0, sizeof(img));
    for (int b = 0; 4 * b < nwords; ++b) {
        uint32_t w[4];
        take_stream_block(seed, pixel, sample, (uint32_t)b, w);
        for (int j = 0; j < 4 && 4 * b + j < 624; ++j) img.x[4 * b + j] = untemper(w[j]);
    }
    img.p = 0;
    std::memcpy((void *)&rng, &img, sizeof(img));
}

inline int stream_words_needed(int max_depth) {
    // 2 pixel reals + per bounce at most 1 (strategy) + 1 (light pick) + 2 (light point) + 3 (plastic)
    long n = 4 + (long)(max_depth + 2) * 14;
    n = (n + 3) & ~3L;
    return (int)std::min<long>(n, 624);
}

struct CameraBasis {
    Vector3 u, v, w;
    Real vw, vh;
};

// src/render.cpp:37-44
CameraBasis camera_basis(const Camera &cam) {
    CameraBasis b;
    Real theta = cam.vfov / 180 * c_PI;
    Real h = tan(theta / 2);
    b.vh = 2 * h;
    b.vw = b.vh / cam.height * cam.width;
    b.w = normalize(cam.lookfrom - cam.lookat);
    b.u = normalize(cross(cam.up, b.w));
    b.v = cross(b.w, b.u);
    return b;
}

typedef Vector3 (*Integrator)(const Scene &, const Ray &, std::mt19937 &);

Integrator pick_integrator(int id) {
    switch (id) {
        case 0: return path_tracing;
        case 1: return path_tracing_raw;
        case 2: return path_tracing_one_sample_MIS;
        case 3: return path_tracing_one_sample_MIS_power;
    }
    return nullptr;
}

// path_tracing_one_sample_MIS_power (path_tracing.h:274-380) picks lights through scene.lights_power_cdf / _pmf, which
// nothing in the reference ever fills (sample_light_power asserts on the empty table, light.cpp:13): dead code as shipped.
// The tables are built here, from the reference's own light_power() (light.cpp:25-30), in the only layout its readers accept
// -- pmf[i] = power_i / total; cdf = the N + 1 running sums from 0 to exactly 1 that sample_light_power's upper_bound +
// clamp walks (light.cpp:9-17).  Summation in light order, plain doubles: take_b200 and oracle/take_oracle.cpp build the same.
void build_power_tables(Scene &sc) {
    const size_t n = sc.lights.size();
    if (sc.lights_power_cdf.size() == n + 1 || n == 0) return;
    std::vector<Real> power(n);
    Real total = 0;
    for (size_t i = 0; i < n; ++i) { power[i] = light_power(sc, sc.lights[i]); total += power[i]; }
    sc.lights_power_pmf.assign(n, Real(0));
    sc.lights_power_cdf.assign(n + 1, Real(0));
    for (size_t i = 0; i < n; ++i) {
        sc.lights_power_pmf[i] = power[i] / total;
        sc.lights_power_cdf[i + 1] = sc.lights_power_cdf[i] + sc.lights_power_pmf[i];
    }
    sc.lights_power_cdf[n] = 1;
}

// One path sample: pixel (x, y) in the reference's y-up loop coordinates (src/render.cpp:65-77).
inline Vector3 one_sample(const Scene &sc, const CameraBasis &b, Integrator f, int x, int y, int64_t s,
                          uint64_t seed, int nwords) {
    const Camera &cam = sc.camera;
    std::mt19937 rng;
    uint32_t pixel = (uint32_t)((cam.height - y - 1) * cam.width + x);  // image-space index (row 0 = top)
    load_stream(rng, seed, pixel, (uint64_t)s, nwords);
    // src/render.cpp:69-75 draws both jitters inside one expression, whose evaluation order C++ leaves
    // unspecified; the seeded restatement fixes it: first draw -> x, second draw -> y.
    Real jx = random_real(rng);
    Real jy = random_real(rng);
    Ray r = {cam.lookfrom,
             normalize(b.u * ((x + jx) / cam.width - Real(0.5)) * b.vw +
                       b.v * ((y + jy) / cam.height - Real(0.5)) * b.vh - b.w),
             c_EPSILON, infinity<Real>()};
    return f(sc, r, rng);
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

template <typename T>
void put(FILE *f, const T *p, size_t n) {
    if (n) fwrite(p, sizeof(T), n, f);
}

struct MatRec {
    int32_t type, tex_id;
    double color[3];
    double uv[4];
    double p[2];
};
struct LightRec {
    int32_t kind, prim_id;
    double intensity[3];
    double position[3];
};

void fill_tex(MatRec &m, const Texture &t) {
    m.tex_id = -1;
    m.color[0] = m.color[1] = m.color[2] = 0;
    m.uv[0] = m.uv[1] = 1;
    m.uv[2] = m.uv[3] = 0;
    if (auto *c = std::get_if<ConstTexture>(&t)) {
        m.color[0] = c->value.x; m.color[1] = c->value.y; m.color[2] = c->value.z;
    } else if (auto *i = std::get_if<ImageTexture>(&t)) {
        m.tex_id = i->texture_id;
        m.uv[0] = i->uscale; m.uv[1] = i->vscale; m.uv[2] = i->uoffset; m.uv[3] = i->voffset;
    }
}

}  // namespace

extern "C" {

const char *ref_last_error() { return g_error.c_str(); }

void *ref_scene_load(const char *xml_path) {
    try {
        auto *rs = new RefScene{parse_scene(fs::path(xml_path))};
        build_bvh(rs->scene);
        return rs;
    } catch (const std::exception &e) {
        g_error = e.what();
        return nullptr;
    }
}

void ref_scene_free(void *h) { delete (RefScene *)h; }

// out[0..7] = #shapes, #bvh nodes, bvh root, #lights, width, height, spp, #materials
void ref_scene_info(void *h, int64_t *out) {
    const Scene &sc = ((RefScene *)h)->scene;
    out[0] = (int64_t)sc.shapes.size();
    out[1] = (int64_t)sc.bvh_nodes.size();
    out[2] = sc.bvh_root_id;
    out[3] = (int64_t)sc.lights.size();
    out[4] = sc.camera.width;
    out[5] = sc.camera.height;
    out[6] = sc.options.spp;
    out[7] = (int64_t)sc.materials.size();
}

// Flatten Scene -> TAKESCN1 file (layout documented in take_b200/sceneio.py).
int ref_scene_dump(void *h, const char *path) {
    const Scene &sc = ((RefScene *)h)->scene;
    std::vector<double> pos, nrm, uv, sph;
    std::vector<int32_t> idx, pmat, plight;
    std::vector<uint8_t> pflags;
    std::vector<int64_t> mesh_base(sc.meshes.size());
    for (size_t m = 0; m < sc.meshes.size(); ++m) {
        const TriangleMesh &mesh = sc.meshes[m];
        mesh_base[m] = (int64_t)pos.size() / 3;
        bool hn = !mesh.normals.empty(), hu = !mesh.uvs.empty();
        for (size_t i = 0; i < mesh.positions.size(); ++i) {
            const Vector3 &p = mesh.positions[i];
            pos.insert(pos.end(), {p.x, p.y, p.z});
            Vector3 n = hn ? mesh.normals[i] : Vector3{0, 0, 0};
            nrm.insert(nrm.end(), {n.x, n.y, n.z});
            Vector2 t = hu ? mesh.uvs[i] : Vector2{0, 0};
            uv.insert(uv.end(), {t.x, t.y});
        }
    }
    for (const Shape &s : sc.shapes) {
        if (auto *tri = std::get_if<Triangle>(&s)) {
            const TriangleMesh &mesh = sc.meshes[tri->mesh_id];
            Vector3i id = mesh.indices[tri->face_id];
            int32_t base = (int32_t)mesh_base[tri->mesh_id];
            idx.insert(idx.end(), {base + id.x, base + id.y, base + id.z});
            pmat.push_back(mesh.material_id);  // what intersect_op uses (src/shape.cpp:85)
            plight.push_back(tri->area_light_id);
            pflags.push_back((uint8_t)((mesh.normals.empty() ? 0 : 1) | (mesh.uvs.empty() ? 0 : 2)));
        } else if (auto *sp = std::get_if<Sphere>(&s)) {
            idx.insert(idx.end(), {(int32_t)(sph.size() / 4), 0, 0});
            sph.insert(sph.end(), {sp->center.x, sp->center.y, sp->center.z, sp->radius});
            pmat.push_back(sp->material_id);
            plight.push_back(sp->area_light_id);
            pflags.push_back(4);
        }
    }
    std::vector<MatRec> mats(sc.materials.size());
    for (size_t i = 0; i < mats.size(); ++i) {
        MatRec &m = mats[i];
        std::memset(&m, 0, sizeof(m));
        const Material &mat = sc.materials[i];
        m.type = (int32_t)mat.index();
        m.tex_id = -1;
        m.uv[0] = m.uv[1] = 1;
        if (auto *d = std::get_if<Diffuse>(&mat)) { fill_tex(m, d->reflectance); }
        else if (auto *d = std::get_if<Mirror>(&mat)) { fill_tex(m, d->reflectance); m.p[0] = d->eta; }
        else if (auto *d = std::get_if<Plastic>(&mat)) { fill_tex(m, d->reflectance); m.p[0] = d->eta; }
        else if (auto *d = std::get_if<Phong>(&mat)) { fill_tex(m, d->reflectance); m.p[0] = d->exponent; }
        else if (auto *d = std::get_if<BlinnPhong>(&mat)) { fill_tex(m, d->reflectance); m.p[0] = d->exponent; }
        else if (auto *d = std::get_if<BlinnPhongMicrofacet>(&mat)) { fill_tex(m, d->reflectance); m.p[0] = d->exponent; }
        else if (auto *d = std::get_if<DisneyDiffuse>(&mat)) { fill_tex(m, d->reflectance); m.p[0] = d->roughness; m.p[1] = d->subsurface; }
        else if (auto *d = std::get_if<DisneyMetal>(&mat)) { fill_tex(m, d->reflectance); }
        else if (auto *d = std::get_if<DisneyGlass>(&mat)) { fill_tex(m, d->reflectance); }
        else if (auto *d = std::get_if<DisneySheen>(&mat)) { fill_tex(m, d->reflectance); }
        else if (auto *d = std::get_if<DisneyBSDF>(&mat)) { fill_tex(m, d->reflectance); }
    }
    std::vector<LightRec> lights(sc.lights.size());
    for (size_t i = 0; i < lights.size(); ++i) {
        LightRec &l = lights[i];
        std::memset(&l, 0, sizeof(l));
        if (auto *a = std::get_if<DiffuseAreaLight>(&sc.lights[i])) {
            l.kind = 1; l.prim_id = a->shape_id;
            l.intensity[0] = a->intensity.x; l.intensity[1] = a->intensity.y; l.intensity[2] = a->intensity.z;
        } else if (auto *p = std::get_if<PointLight>(&sc.lights[i])) {
            l.kind = 0; l.prim_id = -1;
            l.intensity[0] = p->intensity.x; l.intensity[1] = p->intensity.y; l.intensity[2] = p->intensity.z;
            l.position[0] = p->position.x; l.position[1] = p->position.y; l.position[2] = p->position.z;
        }
    }
    FILE *f = fopen(path, "wb");
    if (!f) { g_error = std::string("cannot open ") + path; return -1; }
    fwrite("TAKESCN1", 1, 8, f);
    int64_t hdr[8] = {(int64_t)pos.size() / 3, (int64_t)pmat.size(), (int64_t)mats.size(),
                      (int64_t)sc.textures.image3s.size(), (int64_t)lights.size(), (int64_t)sph.size() / 4,
                      sc.options.spp, 0};
    put(f, hdr, 8);
    const Camera &c = sc.camera;
    int64_t wh[2] = {c.width, c.height};
    put(f, wh, 2);
    double cam[10] = {c.lookfrom.x, c.lookfrom.y, c.lookfrom.z, c.lookat.x, c.lookat.y, c.lookat.z,
                      c.up.x, c.up.y, c.up.z, c.vfov};
    put(f, cam, 10);
    double bg[3] = {sc.background_color.x, sc.background_color.y, sc.background_color.z};
    put(f, bg, 3);
    put(f, pos.data(), pos.size());
    put(f, nrm.data(), nrm.size());
    put(f, uv.data(), uv.size());
    put(f, idx.data(), idx.size());
    put(f, pmat.data(), pmat.size());
    put(f, plight.data(), plight.size());
    if ((idx.size() + pmat.size() + plight.size()) % 2) { int32_t pad = 0; put(f, &pad, 1); }  // keep 8-byte alignment
    pflags.resize((pflags.size() + 7) & ~size_t(7), 0);
    put(f, pflags.data(), pflags.size());
    put(f, sph.data(), sph.size());
    put(f, mats.data(), mats.size());
    put(f, lights.data(), lights.size());
    for (const Image3 &img : sc.textures.image3s) {
        int64_t d[2] = {img.width, img.height};
        put(f, d, 2);
        put(f, (const double *)img.data.data(), img.data.size() * 3);
    }
    fclose(f);
    return 0;
}

// BVH topology exactly as construct_bvh (src/bvh.cpp:8-45) left it.
// box: 6 doubles per node (min xyz, max xyz); links: left,right,prim per node.
void ref_bvh_dump(void *h, double *box, int32_t *links) {
    const Scene &sc = ((RefScene *)h)->scene;
    for (size_t i = 0; i < sc.bvh_nodes.size(); ++i) {
        const BVHNode &n = sc.bvh_nodes[i];
        box[6 * i + 0] = n.box.p_min.x; box[6 * i + 1] = n.box.p_min.y; box[6 * i + 2] = n.box.p_min.z;
        box[6 * i + 3] = n.box.p_max.x; box[6 * i + 4] = n.box.p_max.y; box[6 * i + 5] = n.box.p_max.z;
        links[3 * i + 0] = n.left_node_id; links[3 * i + 1] = n.right_node_id; links[3 * i + 2] = n.primitive_id;
    }
}

// rays: n x {ox,oy,oz,dx,dy,dz,tmin,tmax}.  prim[i] = -1 on miss.  rec (optional): n x 16 doubles
// {pos3, geo_normal3, shading_normal3, uv2, t, material_id, area_light_id, 0, 0}.
// Returns the number of rays whose shim result disagrees bitwise with scene_intersect (expected 0).
int64_t ref_intersect(void *h, const double *rays, int64_t n, int32_t *prim, double *t, double *rec, int nthreads) {
    const Scene &sc = ((RefScene *)h)->scene;
    std::atomic<int64_t> bad{0};
    const int64_t chunk = 4096;
    run_threads(nthreads, (n + chunk - 1) / chunk, [&](int64_t c) {
        for (int64_t i = c * chunk; i < std::min(n, (c + 1) * chunk); ++i) {
            Ray r = make_ray(rays + 8 * i);
            Hit hit = shim(sc, sc.bvh_root_id, r);
            std::optional<Intersection> ref = scene_intersect(sc, r);
            if ((bool)ref != (hit.prim != -1) || (ref && std::memcmp(&ref->t, &hit.t, sizeof(Real)) != 0)) bad++;
            prim[i] = hit.prim;
            t[i] = hit.prim != -1 ? hit.t : 0.0;
            if (rec) {
                double *o = rec + 16 * i;
                std::memset(o, 0, 16 * sizeof(double));
                if (ref) {
                    const Intersection &v = *ref;
                    o[0] = v.pos.x; o[1] = v.pos.y; o[2] = v.pos.z;
                    o[3] = v.geo_normal.x; o[4] = v.geo_normal.y; o[5] = v.geo_normal.z;
                    o[6] = v.shading_normal.x; o[7] = v.shading_normal.y; o[8] = v.shading_normal.z;
                    o[9] = v.uv.x; o[10] = v.uv.y; o[11] = v.t;
                    o[12] = v.material_id; o[13] = v.area_light_id;
                }
            }
        }
    });
    return bad.load();
}

void ref_occluded(void *h, const double *rays, int64_t n, uint8_t *occ, int nthreads) {
    const Scene &sc = ((RefScene *)h)->scene;
    const int64_t chunk = 4096;
    run_threads(nthreads, (n + chunk - 1) / chunk, [&](int64_t c) {
        for (int64_t i = c * chunk; i < std::min(n, (c + 1) * chunk); ++i)
            occ[i] = scene_occluded(sc, make_ray(rays + 8 * i)) ? 1 : 0;
    });
}

// Seeded render of sample indices [spp_begin, spp_end) of every pixel with the reference integrator
// `integrator` (0 = path_tracing, 1 = path_tracing_raw, 2 = path_tracing_one_sample_MIS).
// sum / sumsq: W*H*3 doubles in image layout (row 0 = top, as src/render.cpp:78 writes), accumulated
// in ascending sample order starting from 0.  sumsq may be NULL.
// Only image rows r (from the top) with r % row_step == row_begin are rendered (row_step = 1: all rows), which
// gives bench.py a bounded, unbiased subsample of a frame.
int ref_render(void *h, int integrator, int max_depth, int64_t spp_begin, int64_t spp_end, uint64_t seed,
               int nthreads, double *sum, double *sumsq, int row_begin, int row_step) {
    Scene &sc = ((RefScene *)h)->scene;
    Integrator f = pick_integrator(integrator);
    if (!f) { g_error = "unknown integrator"; return -1; }
    if (integrator == 3) build_power_tables(sc);
    sc.options.max_depth = max_depth;
    const Camera &cam = sc.camera;
    CameraBasis b = camera_basis(cam);
    int nwords = stream_words_needed(max_depth);
    try {
        run_threads(nthreads, cam.height, [&](int64_t y) {
            if (row_step > 1 && (cam.height - y - 1) % row_step != row_begin) return;
            for (int x = 0; x < cam.width; ++x) {
                Vector3 acc{0, 0, 0}, acc2{0, 0, 0};
                for (int64_t s = spp_begin; s < spp_end; ++s) {
                    Vector3 c = one_sample(sc, b, f, x, (int)y, s, seed, nwords);
                    acc += c;
                    acc2 += c * c;
                }
                size_t o = 3 * ((size_t)(cam.height - y - 1) * cam.width + x);
                sum[o] = acc.x; sum[o + 1] = acc.y; sum[o + 2] = acc.z;
                if (sumsq) { sumsq[o] = acc2.x; sumsq[o + 1] = acc2.y; sumsq[o + 2] = acc2.z; }
            }
        });
    } catch (const std::exception &e) {
        g_error = e.what();
        return -1;
    }
    return 0;
}

// Radiance of individual samples: pixel given in image space (px, row py from the top).
int ref_radiance_samples(void *h, int integrator, int max_depth, uint64_t seed, int64_t n, const int32_t *px,
                         const int32_t *py, const int64_t *s, double *out, int nthreads) {
    Scene &sc = ((RefScene *)h)->scene;
    Integrator f = pick_integrator(integrator);
    if (!f) { g_error = "unknown integrator"; return -1; }
    if (integrator == 3) build_power_tables(sc);
    sc.options.max_depth = max_depth;
    CameraBasis b = camera_basis(sc.camera);
    int nwords = stream_words_needed(max_depth);
    const int64_t chunk = 256;
    run_threads(nthreads, (n + chunk - 1) / chunk, [&](int64_t c) {
        for (int64_t i = c * chunk; i < std::min(n, (c + 1) * chunk); ++i) {
            Vector3 v = one_sample(sc, b, f, px[i], sc.camera.height - 1 - py[i], s[i], seed, nwords);
            out[3 * i] = v.x; out[3 * i + 1] = v.y; out[3 * i + 2] = v.z;
        }
    });
    return 0;
}

// Self-check of the mt19937 pre-load trick: returns the number of the first `n` random_real draws of
// stream (seed, pixel, sample) that differ from take_stream_real (expected 0).
int ref_rng_selfcheck(uint64_t seed, uint32_t pixel, uint64_t sample, int n) {
    std::mt19937 rng;
    load_stream(rng, seed, pixel, sample, 624);
    int bad = 0;
    for (int k = 0; k < n && 2 * k + 1 < 624; ++k) {
        double a = random_real(rng);
        double b = take_stream_real(seed, pixel, sample, (uint32_t)k);
        if (std::memcmp(&a, &b, sizeof(double)) != 0) bad++;
    }
    return bad;
}

// The reference's own output step: Image3 of per-pixel means -> imwrite (src/image.cpp:135-175; ".exr" = half B,G,R ZIP
// through the vendored tinyexr).  `mean_rgb` is height*width*3 doubles in image layout.
int ref_imwrite(const char *path, int width, int height, const double *mean_rgb) {
    try {
        Image3 img(width, height);
        for (size_t i = 0; i < img.data.size(); ++i) img.data[i] = Vector3{mean_rgb[3 * i], mean_rgb[3 * i + 1], mean_rgb[3 * i + 2]};
        imwrite(fs::path(path), img);
        return 0;
    } catch (const std::exception &e) {
        g_error = e.what();
        return -1;
    }
}

// The reference's own reader (src/image.cpp imread3): returns 0 and fills out[height*width*3]; -2 if `cap` is too small.
int ref_imread3(const char *path, int *width, int *height, double *out, int64_t cap) {
    try {
        Image3 img = imread3(fs::path(path));
        *width = img.width; *height = img.height;
        if ((int64_t)img.data.size() * 3 > cap) return -2;
        for (size_t i = 0; i < img.data.size(); ++i) { out[3 * i] = img.data[i].x; out[3 * i + 1] = img.data[i].y; out[3 * i + 2] = img.data[i].z; }
        return 0;
    } catch (const std::exception &e) {
        g_error = e.what();
        return -1;
    }
}

}  // extern "C"
