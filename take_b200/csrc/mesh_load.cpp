// Scene-description builder: big meshes straight into the flat arrays of TakeSceneDesc.
//
// The reference turns a mesh into the renderer's input in three single-threaded steps: tinyply reads the file into
// temporary buffers and parse_ply copies them element by element into std::vector<Vector3> (src/parse/parse_ply.cpp:16-120);
// compute_normals walks the faces once more when the file has no normals (src/compute_normals.cpp:12-47); and parse_shape
// pushes one std::variant<Sphere, Triangle> -- plus, on an emitter, one std::variant<PointLight, DiffuseAreaLight> -- per
// face (src/parse/parse_scene.cpp:937-945), which a GPU host then has to take apart again.  Here the file goes through
// one read, the per-vertex / per-face conversions run in chunks on all host threads, and the per-face expansion is a
// fill of three plain arrays.  Every value is produced by the reference's own FP64 operations in its order (xform_point,
// xform_normal, normalize, the angle-weighted normal sum in FACE ORDER per vertex, asin from the host's libm), so the
// arrays are bit-identical to what `flatten(parse_scene(...))` yields -- tests/test_mesh_load.py checks that against the
// unmodified reference parser.
//
// Host-only code: no CUDA here.
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/take_gpu.h"

extern "C" int take_exr_fail(int code, const std::string &msg);  // sets take_gpu_last_error (take_gpu.cu)

namespace {

int fail(int code, const std::string &msg) { return take_exr_fail(code, msg); }

template <typename F>
void chunks(int64_t n, F f, int64_t min_parallel = 1 << 15) {
    const int hw = (int)std::max(1u, std::thread::hardware_concurrency());
    const int parts = n >= min_parallel ? (int)std::min<int64_t>(hw, n / (min_parallel / 2)) : 1;
    if (parts <= 1) { f((int64_t)0, n, 0); return; }
    std::vector<std::thread> pool;
    for (int t = 1; t < parts; ++t) pool.emplace_back(f, n * t / parts, n * (t + 1) / parts, t);
    f((int64_t)0, n / parts, 0);
    for (auto &t : pool) t.join();
}

struct V3 { double x, y, z; };
inline V3 sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 add(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 mul(V3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }                    // vector.h:222-225
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline double length(V3 v) { return sqrt(dot(v, v)); }                                          // vector.h:240-247
inline V3 divs(V3 v, double s) { const double inv = 1.0 / s; return {v.x * inv, v.y * inv, v.z * inv}; }  // vector.h:193-197
inline V3 normalize(V3 v) { const double l = length(v); return l <= 0 ? V3{0, 0, 0} : divs(v, l); }  // vector.h:249-257

// xform_point, src/transform.cpp:79-87 (m row-major)
inline V3 xform_point(const double *m, V3 p) {
    const double x = m[0] * p.x + m[1] * p.y + m[2] * p.z + m[3];
    const double y = m[4] * p.x + m[5] * p.y + m[6] * p.z + m[7];
    const double z = m[8] * p.x + m[9] * p.y + m[10] * p.z + m[11];
    const double w = m[12] * p.x + m[13] * p.y + m[14] * p.z + m[15];
    const double inv_w = 1.0 / w;
    return {x * inv_w, y * inv_w, z * inv_w};
}
// xform_normal, src/transform.cpp:95-100: normalize(transpose(inv) * n)
inline V3 xform_normal(const double *inv, V3 n) {
    return normalize(V3{inv[0] * n.x + inv[4] * n.y + inv[8] * n.z, inv[1] * n.x + inv[5] * n.y + inv[9] * n.z,
                        inv[2] * n.x + inv[6] * n.y + inv[10] * n.z});
}
// unit_angle, src/compute_normals.cpp:4-10
inline double unit_angle(V3 u, V3 v) {
    const double kPi = 3.14159265358979323846;
    if (dot(u, v) < 0) return (kPi - 2) * asin(0.5 * length(add(v, u)));
    return 2 * asin(0.5 * length(sub(v, u)));
}

const double kIdentity[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};

// ---- PLY ---------------------------------------------------------------------------------------------------------
enum PlyType { T_I8, T_U8, T_I16, T_U16, T_I32, T_U32, T_F32, T_F64, T_BAD };
const int kTypeSize[] = {1, 1, 2, 2, 4, 4, 4, 8, 0};
PlyType ply_type(const std::string &s) {
    if (s == "char" || s == "int8") return T_I8;
    if (s == "uchar" || s == "uint8") return T_U8;
    if (s == "short" || s == "int16") return T_I16;
    if (s == "ushort" || s == "uint16") return T_U16;
    if (s == "int" || s == "int32") return T_I32;
    if (s == "uint" || s == "uint32") return T_U32;
    if (s == "float" || s == "float32") return T_F32;
    if (s == "double" || s == "float64") return T_F64;
    return T_BAD;
}
struct PlyProp {
    std::string name;
    PlyType type = T_BAD, count_type = T_BAD;
    bool is_list = false;
    int offset = 0;  // byte offset inside a fixed-size row (elements without list properties)
};
struct PlyElement {
    std::string name;
    int64_t count = 0;
    std::vector<PlyProp> props;
    bool has_list = false;
    int row_bytes = 0;
    const PlyProp *find(const char *n) const {
        for (auto &p : props) if (p.name == n) return &p;
        return nullptr;
    }
};

template <bool SWAP>
inline double read_scalar(const unsigned char *p, PlyType t) {
    unsigned char b[8];
    const int n = kTypeSize[t];
    if (SWAP) for (int i = 0; i < n; ++i) b[i] = p[n - 1 - i];
    else memcpy(b, p, n);
    switch (t) {
        case T_I8: { int8_t v; memcpy(&v, b, 1); return v; }
        case T_U8: { uint8_t v; memcpy(&v, b, 1); return v; }
        case T_I16: { int16_t v; memcpy(&v, b, 2); return v; }
        case T_U16: { uint16_t v; memcpy(&v, b, 2); return v; }
        case T_I32: { int32_t v; memcpy(&v, b, 4); return v; }
        case T_U32: { uint32_t v; memcpy(&v, b, 4); return v; }
        case T_F32: { float v; memcpy(&v, b, 4); return v; }     // float -> double widening, as Vector3{data[3 * i], ...}
        case T_F64: { double v; memcpy(&v, b, 8); return v; }
        default: return 0;
    }
}
// the reference stores every index in an int (Vector3i): values wrap the way its conversion does
template <bool SWAP>
inline int32_t read_index(const unsigned char *p, PlyType t) {
    unsigned char b[4] = {0, 0, 0, 0};
    const int n = kTypeSize[t];
    if (SWAP) for (int i = 0; i < n; ++i) b[i] = p[n - 1 - i];
    else memcpy(b, p, n);
    switch (t) {
        case T_I8: { int8_t v; memcpy(&v, b, 1); return v; }
        case T_U8: { uint8_t v; memcpy(&v, b, 1); return v; }
        case T_I16: { int16_t v; memcpy(&v, b, 2); return v; }
        case T_U16: { uint16_t v; memcpy(&v, b, 2); return v; }
        case T_I32: { int32_t v; memcpy(&v, b, 4); return v; }
        case T_U32: { uint32_t v; memcpy(&v, b, 4); return (int32_t)v; }
        default: return 0;
    }
}

struct Mesh {  // one parsed mesh, world space (what TriangleMesh holds after parse_ply + compute_normals)
    std::vector<double> pos, nrm, uv;  // 3, 3, 2 per vertex; nrm / uv empty when absent
    std::vector<int32_t> idx;          // 3 per face, mesh-local
};

}  // namespace

struct TakeDescBuilder {
    std::vector<double> positions, normals, uvs, spheres;
    std::vector<int32_t> indices, prim_material, prim_light;
    std::vector<uint8_t> prim_flags;
    std::vector<TakeLightDesc> lights;
    double ms_read = 0, ms_convert = 0, ms_normals = 0, ms_append = 0;  // of the last add_ply
};

namespace {

double now_ms() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

// compute_normals, src/compute_normals.cpp:12-47, on all host threads with the reference's summation order: a vertex's
// normal is the sum of its incident faces' (normal x corner angle) IN FACE ORDER, so the incidences are bucketed by vertex
// with the face order kept (each thread owns a vertex range and scans the index array for it) and every vertex is then
// summed by one thread.
void compute_normals(const std::vector<double> &pos, const std::vector<int32_t> &idx, std::vector<double> &nrm) {
    const int64_t nv = (int64_t)pos.size() / 3, nf = (int64_t)idx.size() / 3;
    nrm.assign((size_t)nv * 3, 0.0);
    if (nv == 0) return;
    std::vector<uint32_t> start((size_t)nv + 1, 0);
    {
        std::vector<std::atomic<uint32_t>> cnt((size_t)nv);
        for (auto &c : cnt) c.store(0, std::memory_order_relaxed);
        chunks(3 * nf, [&](int64_t a, int64_t b, int) {
            for (int64_t i = a; i < b; ++i) cnt[idx[i]].fetch_add(1, std::memory_order_relaxed);
        });
        uint32_t acc = 0;
        for (int64_t v = 0; v < nv; ++v) { start[v] = acc; acc += cnt[v].load(std::memory_order_relaxed); }
        start[nv] = acc;
    }
    std::vector<int32_t> inc((size_t)3 * nf);  // incidences (3 * face + corner), grouped by vertex, ascending inside a group
    chunks(nv, [&](int64_t v0, int64_t v1, int) {
        if (v0 >= v1) return;
        std::vector<uint32_t> fill(start.begin() + v0, start.begin() + v1);
        for (int64_t i = 0; i < 3 * nf; ++i) {
            const int64_t v = idx[i];
            if (v >= v0 && v < v1) inc[fill[v - v0]++] = (int32_t)i;
        }
    }, 1 << 12);
    auto P = [&](int32_t v) { return V3{pos[3 * (int64_t)v], pos[3 * (int64_t)v + 1], pos[3 * (int64_t)v + 2]}; };
    chunks(nv, [&](int64_t v0, int64_t v1, int) {
        for (int64_t v = v0; v < v1; ++v) {
            V3 sum = {0, 0, 0};
            for (uint32_t k = start[v]; k < start[v + 1]; ++k) {
                const int64_t f = inc[k] / 3;
                const int c = inc[k] % 3;
                const int32_t *id = idx.data() + 3 * f;
                // the face normal comes from corner 0 (compute_normals.cpp:26-33); a zero-area face contributes nothing
                V3 n = cross(sub(P(id[1]), P(id[0])), sub(P(id[2]), P(id[0])));
                const double l = length(n);
                if (l == 0) continue;
                n = divs(n, l);
                const V3 p0 = P(id[c]), p1 = P(id[(c + 1) % 3]), p2 = P(id[(c + 2) % 3]);
                const double angle = unit_angle(normalize(sub(p1, p0)), normalize(sub(p2, p0)));
                sum = add(sum, mul(n, angle));
            }
            const double l = length(sum);
            const V3 r = l != 0 ? divs(sum, l) : V3{0, 0, 0};
            nrm[3 * v] = r.x; nrm[3 * v + 1] = r.y; nrm[3 * v + 2] = r.z;
        }
    }, 1 << 12);
}

int load_ply(const char *path, const double *to_world, const double *inv_to_world, Mesh &m, std::string &err, double *ms) {
    const double t0 = now_ms();
    FILE *f = fopen(path, "rb");
    if (!f) { err = std::string("cannot open ") + path; return TAKE_E_INVALID; }
    fseek(f, 0, SEEK_END);
    const int64_t size = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<unsigned char> buf;
    try { buf.resize((size_t)size); } catch (const std::bad_alloc &) { fclose(f); err = "out of memory"; return TAKE_E_NOMEM; }
    const bool ok = size == 0 || fread(buf.data(), 1, (size_t)size, f) == (size_t)size;
    fclose(f);
    if (!ok) { err = std::string("short read: ") + path; return TAKE_E_INVALID; }
    ms[0] = now_ms() - t0;

    // header
    int64_t p = 0;
    auto line = [&](std::string &out) {
        if (p >= size) return false;
        int64_t e = p;
        while (e < size && buf[e] != '\n') ++e;
        out.assign((const char *)buf.data() + p, (size_t)(e - p));
        while (!out.empty() && (out.back() == '\r' || out.back() == ' ')) out.pop_back();
        p = e + 1;
        return true;
    };
    std::string ln;
    if (!line(ln) || ln != "ply") { err = std::string("not a PLY file: ") + path; return TAKE_E_INVALID; }
    int format = -1;  // 0 ascii, 1 binary little endian, 2 binary big endian
    std::vector<PlyElement> elems;
    bool ended = false;
    while (line(ln)) {
        char a[64] = "", b[64] = "", c[64] = "", d[64] = "", e[64] = "";
        const int n = sscanf(ln.c_str(), "%63s %63s %63s %63s %63s", a, b, c, d, e);
        if (n <= 0) continue;
        const std::string k = a;
        if (k == "format") {
            format = !strcmp(b, "ascii") ? 0 : !strcmp(b, "binary_little_endian") ? 1 : !strcmp(b, "binary_big_endian") ? 2 : -1;
        } else if (k == "element" && n >= 3) {
            PlyElement el;
            el.name = b;
            el.count = atoll(c);
            elems.push_back(el);
        } else if (k == "property" && !elems.empty()) {
            PlyProp pr;
            if (!strcmp(b, "list") && n >= 5) { pr.is_list = true; pr.count_type = ply_type(c); pr.type = ply_type(d); pr.name = e; }
            else if (n >= 3) { pr.type = ply_type(b); pr.name = c; }
            if (pr.type == T_BAD || (pr.is_list && pr.count_type == T_BAD)) { err = "unknown PLY property type: " + ln; return TAKE_E_INVALID; }
            PlyElement &el = elems.back();
            pr.offset = el.row_bytes;
            if (pr.is_list) el.has_list = true; else el.row_bytes += kTypeSize[pr.type];
            el.props.push_back(pr);
        } else if (k == "end_header") {
            ended = true;
            break;
        }
    }
    if (!ended || format < 0) { err = std::string("bad PLY header: ") + path; return TAKE_E_INVALID; }
    const PlyElement *ve = nullptr, *fe = nullptr;
    for (auto &el : elems) { if (el.name == "vertex") ve = &el; if (el.name == "face") fe = &el; }
    // parse_ply.cpp:15-34: positions and vertex_indices are required; uvs / normals are used when all their properties
    // exist with one common type (tinyply throws otherwise and the reference carries on without them)
    auto group = [&](const PlyElement *el, std::initializer_list<const char *> names, const PlyProp **out) -> bool {
        int i = 0;
        PlyType t = T_BAD;
        for (const char *nm : names) {
            const PlyProp *pr = el ? el->find(nm) : nullptr;
            if (!pr || pr->is_list) return false;
            if (i && pr->type != t) return false;
            t = pr->type;
            out[i++] = pr;
        }
        return true;
    };
    const PlyProp *pp[3], *pn[3], *pu[2];
    if (!ve || !group(ve, {"x", "y", "z"}, pp)) { err = std::string("Vertex positions not found in ") + path; return TAKE_E_INVALID; }
    if (pp[0]->type != T_F32 && pp[0]->type != T_F64) { err = std::string("vertex positions must be float or double: ") + path; return TAKE_E_INVALID; }
    bool has_n = group(ve, {"nx", "ny", "nz"}, pn) && (pn[0]->type == T_F32 || pn[0]->type == T_F64);
    bool has_uv = group(ve, {"u", "v"}, pu) && (pu[0]->type == T_F32 || pu[0]->type == T_F64);
    const PlyProp *pf = fe ? fe->find("vertex_indices") : nullptr;
    if (!pf || !pf->is_list || pf->type > T_U32) { err = std::string("Vertex indices not found in ") + path; return TAKE_E_INVALID; }
    if (ve->has_list) { err = std::string("list properties on vertices are not supported: ") + path; return TAKE_E_INVALID; }
    const int64_t nv = ve->count, nf = fe->count;
    if (nv < 0 || nf < 0 || nv >= ((int64_t)1 << 31)) { err = "bad PLY element counts"; return TAKE_E_INVALID; }
    // (every vertex and every face takes at least one byte of the file, in every format: a header that promises more is
    //  refused before anything is sized by it)
    if (nv > size || nf > size) { err = std::string("PLY element counts exceed the file size: ") + path; return TAKE_E_INVALID; }
    if (!to_world) to_world = kIdentity;
    if (!inv_to_world) inv_to_world = kIdentity;
    try {
        m.pos.resize((size_t)nv * 3);
        if (has_n) m.nrm.resize((size_t)nv * 3);
        if (has_uv) m.uv.resize((size_t)nv * 2);
        m.idx.resize((size_t)nf * 3);
    } catch (const std::exception &) { err = "out of memory"; return TAKE_E_NOMEM; }
    const double t1 = now_ms();
    std::atomic<int> bad(0);  // 1: face with != 3 indices, 2: truncated data
    auto put_vertex = [&](int64_t i, const double *xyz, const double *n3, const double *uv2) {
        const V3 w = xform_point(to_world, V3{xyz[0], xyz[1], xyz[2]});
        m.pos[3 * i] = w.x; m.pos[3 * i + 1] = w.y; m.pos[3 * i + 2] = w.z;
        if (has_n) {
            const V3 n = xform_normal(inv_to_world, V3{n3[0], n3[1], n3[2]});
            m.nrm[3 * i] = n.x; m.nrm[3 * i + 1] = n.y; m.nrm[3 * i + 2] = n.z;
        }
        if (has_uv) { m.uv[2 * i] = uv2[0]; m.uv[2 * i + 1] = uv2[1]; }
    };
    if (format == 0) {  // ascii: sequential tokens (small files; the big ones are binary)
        const char *s = (const char *)buf.data() + p, *end = (const char *)buf.data() + size;
        std::string text(s, (size_t)(end - s));  // NUL-terminated copy for strtod
        const char *q = text.c_str();
        // (a float property is read as a float, the way tinyply's operator>> into a float does: strtof, not strtod + rounding)
        auto next = [&](double &v, bool f32 = false) {
            char *e2;
            v = f32 ? (double)strtof(q, &e2) : strtod(q, &e2);
            if (e2 == q) return false;
            q = e2;
            return true;
        };
        for (auto &el : elems) {
            for (int64_t i = 0; i < el.count; ++i) {
                double xyz[3] = {0, 0, 0}, n3[3] = {0, 0, 0}, uv2[2] = {0, 0};
                for (auto &pr : el.props) {
                    if (pr.is_list) {
                        double c;
                        if (!next(c)) { err = "truncated PLY data"; return TAKE_E_INVALID; }
                        const int cn = (int)c;
                        if (&el == fe && &pr == pf && cn != 3) { err = std::string("only triangles are supported (a face has ") + std::to_string(cn) + " vertices): " + path; return TAKE_E_INVALID; }
                        for (int k = 0; k < cn; ++k) {
                            double v;
                            if (!next(v)) { err = "truncated PLY data"; return TAKE_E_INVALID; }
                            if (&el == fe && &pr == pf) m.idx[3 * i + k] = (int32_t)(int64_t)v;
                        }
                    } else {
                        double v;
                        if (!next(v, pr.type == T_F32)) { err = "truncated PLY data"; return TAKE_E_INVALID; }
                        if (&el == ve) {
                            const double vv = v;
                            for (int k = 0; k < 3; ++k) { if (&pr == pp[k]) xyz[k] = vv; if (has_n && &pr == pn[k]) n3[k] = vv; }
                            for (int k = 0; k < 2; ++k) if (has_uv && &pr == pu[k]) uv2[k] = vv;
                        }
                    }
                }
                if (&el == ve) put_vertex(i, xyz, n3, uv2);
            }
        }
    } else {
        const bool swap = format == 2;
        int64_t off = p;
        for (auto &el : elems) {
            if (&el == ve) {
                if (off + (int64_t)el.row_bytes * nv > size) { err = "truncated PLY data"; return TAKE_E_INVALID; }
                const unsigned char *base = buf.data() + off;
                const int rb = el.row_bytes;
                chunks(nv, [&](int64_t a, int64_t b, int) {
                    for (int64_t i = a; i < b; ++i) {
                        const unsigned char *r = base + i * rb;
                        double xyz[3], n3[3] = {0, 0, 0}, uv2[2] = {0, 0};
                        for (int k = 0; k < 3; ++k) xyz[k] = swap ? read_scalar<true>(r + pp[k]->offset, pp[k]->type) : read_scalar<false>(r + pp[k]->offset, pp[k]->type);
                        if (has_n) for (int k = 0; k < 3; ++k) n3[k] = swap ? read_scalar<true>(r + pn[k]->offset, pn[k]->type) : read_scalar<false>(r + pn[k]->offset, pn[k]->type);
                        if (has_uv) for (int k = 0; k < 2; ++k) uv2[k] = swap ? read_scalar<true>(r + pu[k]->offset, pu[k]->type) : read_scalar<false>(r + pu[k]->offset, pu[k]->type);
                        put_vertex(i, xyz, n3, uv2);
                    }
                });
                off += (int64_t)rb * nv;
            } else if (&el == fe && el.props.size() == 1) {
                // the common layout: each face is (count, 3 indices) -- fixed-size records as long as every face is a triangle
                const int cs = kTypeSize[pf->count_type], is = kTypeSize[pf->type], rb = cs + 3 * is;
                if (off + (int64_t)rb * nf > size) { err = "truncated PLY data"; return TAKE_E_INVALID; }
                const unsigned char *base = buf.data() + off;
                chunks(nf, [&](int64_t a, int64_t b, int) {
                    for (int64_t i = a; i < b; ++i) {
                        const unsigned char *r = base + i * rb;
                        const int cn = (int)(swap ? read_scalar<true>(r, pf->count_type) : read_scalar<false>(r, pf->count_type));
                        if (cn != 3) { bad.store(1); return; }
                        for (int k = 0; k < 3; ++k) m.idx[3 * i + k] = swap ? read_index<true>(r + cs + k * is, pf->type) : read_index<false>(r + cs + k * is, pf->type);
                    }
                });
                if (bad.load() == 1) { err = std::string("only triangles are supported (a face does not have 3 vertices): ") + path; return TAKE_E_INVALID; }
                off += (int64_t)rb * nf;
            } else {
                // any other element (or faces with extra properties): walk it row by row
                for (int64_t i = 0; i < el.count; ++i) {
                    for (auto &pr : el.props) {
                        if (!pr.is_list) { off += kTypeSize[pr.type]; continue; }
                        if (off + kTypeSize[pr.count_type] > size) { err = "truncated PLY data"; return TAKE_E_INVALID; }
                        const int cn = (int)(swap ? read_scalar<true>(buf.data() + off, pr.count_type) : read_scalar<false>(buf.data() + off, pr.count_type));
                        off += kTypeSize[pr.count_type];
                        if (cn < 0 || off + (int64_t)cn * kTypeSize[pr.type] > size) { err = "truncated PLY data"; return TAKE_E_INVALID; }
                        if (&el == fe && &pr == pf) {
                            if (cn != 3) { err = std::string("only triangles are supported (a face does not have 3 vertices): ") + path; return TAKE_E_INVALID; }
                            for (int k = 0; k < 3; ++k)
                                m.idx[3 * i + k] = swap ? read_index<true>(buf.data() + off + k * kTypeSize[pr.type], pr.type)
                                                        : read_index<false>(buf.data() + off + k * kTypeSize[pr.type], pr.type);
                        }
                        off += (int64_t)cn * kTypeSize[pr.type];
                    }
                    if (off > size) { err = "truncated PLY data"; return TAKE_E_INVALID; }
                }
            }
        }
    }
    // the reference indexes with .at() and would throw later (shape.cpp:46-50): refuse here
    std::atomic<int> oob(0);
    chunks(3 * nf, [&](int64_t a, int64_t b, int) {
        for (int64_t i = a; i < b; ++i) if (m.idx[i] < 0 || m.idx[i] >= nv) { oob.store(1); return; }
    });
    if (oob.load()) { err = std::string("vertex index out of range in ") + path; return TAKE_E_INVALID; }
    ms[1] = now_ms() - t1;
    return TAKE_OK;
}

// parse_scene.cpp:934-945: one Triangle shape per face in file order; on an emitter every face is its own DiffuseAreaLight
int append_mesh(TakeDescBuilder *b, const Mesh &m, int material_id, const double *radiance) {
    const int64_t nv = (int64_t)m.pos.size() / 3, nf = (int64_t)m.idx.size() / 3;
    const int64_t base_v = (int64_t)b->positions.size() / 3, base_p = (int64_t)b->prim_material.size();
    if (base_v + nv >= ((int64_t)1 << 31) || base_p + nf >= ((int64_t)1 << 28)) return fail(TAKE_E_INVALID, "scene too large");
    const bool has_n = !m.nrm.empty(), has_uv = !m.uv.empty();
    try {
        b->positions.insert(b->positions.end(), m.pos.begin(), m.pos.end());
        b->normals.resize((size_t)(base_v + nv) * 3, 0.0);
        b->uvs.resize((size_t)(base_v + nv) * 2, 0.0);
        if (has_n) std::copy(m.nrm.begin(), m.nrm.end(), b->normals.begin() + base_v * 3);
        if (has_uv) std::copy(m.uv.begin(), m.uv.end(), b->uvs.begin() + base_v * 2);
        b->indices.resize((size_t)(base_p + nf) * 3);
        b->prim_material.resize((size_t)(base_p + nf), material_id);
        b->prim_light.resize((size_t)(base_p + nf), -1);
        b->prim_flags.resize((size_t)(base_p + nf), (uint8_t)((has_n ? TAKE_PRIM_HAS_NORMALS : 0) | (has_uv ? TAKE_PRIM_HAS_UVS : 0)));
        const int64_t base_l = (int64_t)b->lights.size();
        if (radiance) b->lights.resize((size_t)(base_l + nf));
        chunks(nf, [&](int64_t a0, int64_t a1, int) {
            for (int64_t f = a0; f < a1; ++f) {
                for (int k = 0; k < 3; ++k) b->indices[3 * (base_p + f) + k] = (int32_t)(m.idx[3 * f + k] + base_v);
                if (radiance) {
                    b->prim_light[base_p + f] = (int32_t)(base_l + f);
                    TakeLightDesc &l = b->lights[base_l + f];
                    memset(&l, 0, sizeof(l));
                    l.kind = TAKE_LIGHT_AREA;
                    l.prim_id = (int32_t)(base_p + f);
                    l.intensity[0] = radiance[0]; l.intensity[1] = radiance[1]; l.intensity[2] = radiance[2];
                }
            }
        });
    } catch (const std::bad_alloc &) {
        return fail(TAKE_E_NOMEM, "out of host memory while appending a mesh");
    }
    return TAKE_OK;
}

}  // namespace

extern "C" {

int take_gpu_builder_create(TakeDescBuilder **out) {
    if (!out) return fail(TAKE_E_INVALID, "null argument");
    *out = new (std::nothrow) TakeDescBuilder;
    return *out ? TAKE_OK : fail(TAKE_E_NOMEM, "out of host memory");
}

int take_gpu_builder_destroy(TakeDescBuilder *b) {
    delete b;
    return TAKE_OK;
}

int take_gpu_builder_add_ply(TakeDescBuilder *b, const char *path, const double *to_world, const double *inv_to_world,
                             int32_t material_id, int32_t face_normals, const double *radiance) {
    if (!b || !path) return fail(TAKE_E_INVALID, "null argument");
    try {
    Mesh m;
    std::string err;
    double ms[2] = {0, 0};
    if (int rc = load_ply(path, to_world, inv_to_world, m, err, ms)) return fail(rc, err);
    b->ms_read = ms[0];
    b->ms_convert = ms[1];
    double t0 = now_ms();
    // parse_scene.cpp:826-834: face_normals drops the normals; otherwise a mesh without normals gets angle-weighted ones
    if (face_normals) m.nrm.clear();
    else if (m.nrm.empty()) compute_normals(m.pos, m.idx, m.nrm);
    b->ms_normals = now_ms() - t0;
    t0 = now_ms();
    const int rc = append_mesh(b, m, material_id, radiance);
    b->ms_append = now_ms() - t0;
    return rc;
    } catch (const std::bad_alloc &) {
        return fail(TAKE_E_NOMEM, "out of memory");
    } catch (const std::exception &e) {   // nothing may propagate through the C boundary
        return fail(TAKE_E_INVALID, std::string("PLY loader: ") + e.what());
    }
}

int take_gpu_builder_add_mesh(TakeDescBuilder *b, int64_t num_vertices, const double *positions, const double *normals, const double *uvs,
                              int64_t num_faces, const int32_t *indices, int32_t material_id, int32_t compute_missing_normals,
                              const double *radiance) {
    if (!b || num_vertices < 0 || num_faces < 0 || (num_vertices > 0 && !positions) || (num_faces > 0 && !indices))
        return fail(TAKE_E_INVALID, "bad mesh arguments");
    for (int64_t i = 0; i < 3 * num_faces; ++i)
        if (indices[i] < 0 || indices[i] >= num_vertices) return fail(TAKE_E_INVALID, "vertex index out of range");
    Mesh m;
    try {
        m.pos.assign(positions, positions + 3 * num_vertices);
        if (normals) m.nrm.assign(normals, normals + 3 * num_vertices);
        if (uvs) m.uv.assign(uvs, uvs + 2 * num_vertices);
        m.idx.assign(indices, indices + 3 * num_faces);
    } catch (const std::bad_alloc &) {
        return fail(TAKE_E_NOMEM, "out of host memory");
    }
    if (!normals && compute_missing_normals) compute_normals(m.pos, m.idx, m.nrm);
    return append_mesh(b, m, material_id, radiance);
}

// <shape type="sphere">, parse_scene.cpp:785-808
int take_gpu_builder_add_sphere(TakeDescBuilder *b, const double *center, double radius, int32_t material_id, const double *radiance) {
    if (!b || !center) return fail(TAKE_E_INVALID, "null argument");
    const int32_t prim = (int32_t)b->prim_material.size();
    const int32_t sph = (int32_t)(b->spheres.size() / 4);
    b->spheres.insert(b->spheres.end(), {center[0], center[1], center[2], radius});
    b->indices.insert(b->indices.end(), {sph, 0, 0});
    b->prim_material.push_back(material_id);
    b->prim_flags.push_back(TAKE_PRIM_SPHERE);
    if (radiance) {
        b->prim_light.push_back((int32_t)b->lights.size());
        TakeLightDesc l;
        memset(&l, 0, sizeof(l));
        l.kind = TAKE_LIGHT_AREA; l.prim_id = prim;
        l.intensity[0] = radiance[0]; l.intensity[1] = radiance[1]; l.intensity[2] = radiance[2];
        b->lights.push_back(l);
    } else {
        b->prim_light.push_back(-1);
    }
    return TAKE_OK;
}

// <emitter type="point"> at scene level, parse_scene.cpp:701-727 (takes the next light index, like any emitter)
int take_gpu_builder_add_point_light(TakeDescBuilder *b, const double *intensity, const double *position) {
    if (!b || !intensity || !position) return fail(TAKE_E_INVALID, "null argument");
    TakeLightDesc l;
    memset(&l, 0, sizeof(l));
    l.kind = TAKE_LIGHT_POINT; l.prim_id = -1;
    for (int k = 0; k < 3; ++k) { l.intensity[k] = intensity[k]; l.position[k] = position[k]; }
    b->lights.push_back(l);
    return TAKE_OK;
}

// Fills the geometry and light fields of `desc` (counts and pointers into the builder, valid until the next add / destroy);
// camera, background, materials, textures and the environment fields are the caller's.
int take_gpu_builder_finish(TakeDescBuilder *b, TakeSceneDesc *desc) {
    if (!b || !desc) return fail(TAKE_E_INVALID, "null argument");
    desc->num_vertices = (int64_t)b->positions.size() / 3;
    desc->positions = b->positions.data();
    desc->normals = b->normals.data();
    desc->uvs = b->uvs.data();
    desc->num_prims = (int64_t)b->prim_material.size();
    desc->indices = b->indices.data();
    desc->prim_material = b->prim_material.data();
    desc->prim_light = b->prim_light.data();
    desc->prim_flags = b->prim_flags.data();
    desc->num_spheres = (int64_t)b->spheres.size() / 4;
    desc->spheres = b->spheres.data();
    desc->num_lights = (int32_t)b->lights.size();
    desc->lights = b->lights.data();
    return TAKE_OK;
}

// TAKESCN1 writer (the exchange format of take_b200/sceneio.py: the flat arrays of a TakeSceneDesc): lets a C++ host hand a
// scene to the Python side -- tools, tests, api.render -- without going through XML again.
int take_gpu_scene_desc_save(const TakeSceneDesc *d, int64_t spp, const char *path) {
    if (!d || !path) return fail(TAKE_E_INVALID, "null argument");
    FILE *f = fopen(path, "wb");
    if (!f) return fail(TAKE_E_INVALID, std::string("cannot write ") + path);
    bool ok = true;
    auto put = [&](const void *p, size_t bytes) { if (bytes && fwrite(p, 1, bytes, f) != bytes) ok = false; };
    const bool env = d->env_rgb && d->env_width > 0 && d->env_height > 0;
    const int64_t np_ = d->num_prims, nv = d->num_vertices;
    const int64_t hdr[10] = {nv, np_, d->num_materials, d->num_textures, d->num_lights, d->num_spheres, spp,
                             (env ? 1 : 0) | (env && d->env_sample ? 2 : 0), d->camera.width, d->camera.height};
    put("TAKESCN1", 8);
    put(hdr, sizeof(hdr));
    put(d->camera.lookfrom, 24); put(d->camera.lookat, 24); put(d->camera.up, 24); put(&d->camera.vfov, 8);
    put(d->background, 24);
    put(d->positions, (size_t)nv * 24); put(d->normals, (size_t)nv * 24); put(d->uvs, (size_t)nv * 16);
    put(d->indices, (size_t)np_ * 12); put(d->prim_material, (size_t)np_ * 4); put(d->prim_light, (size_t)np_ * 4);
    const int32_t zero4 = 0;
    if ((5 * np_) % 2) put(&zero4, 4);
    put(d->prim_flags, (size_t)np_);
    const unsigned char zero8[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    put(zero8, (size_t)(((np_ + 7) & ~(int64_t)7) - np_));
    put(d->spheres, (size_t)d->num_spheres * 32);
    put(d->materials, (size_t)d->num_materials * sizeof(TakeMaterialDesc));
    put(d->lights, (size_t)d->num_lights * sizeof(TakeLightDesc));
    for (int i = 0; i < d->num_textures; ++i) {
        const int64_t wh[2] = {d->textures[i].width, d->textures[i].height};
        put(wh, 16);
        put(d->textures[i].rgb, (size_t)wh[0] * wh[1] * 24);
    }
    if (env) {
        const int64_t wh[2] = {d->env_width, d->env_height};
        put(wh, 16);
        put(d->env_rgb, (size_t)wh[0] * wh[1] * 24);
    }
    ok = (fclose(f) == 0) && ok;
    return ok ? TAKE_OK : fail(TAKE_E_INVALID, std::string("write failed: ") + path);
}

// out[0..3] = file read, conversion, vertex normals, append -- milliseconds of the last take_gpu_builder_add_ply
int take_gpu_builder_timings(TakeDescBuilder *b, double *out) {
    if (!b || !out) return fail(TAKE_E_INVALID, "null argument");
    out[0] = b->ms_read; out[1] = b->ms_convert; out[2] = b->ms_normals; out[3] = b->ms_append;
    return TAKE_OK;
}

}  // extern "C"
