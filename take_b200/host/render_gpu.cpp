// render_gpu.cpp -- drop-in replacement for the reference's src/render.cpp.
//
// A maintainer of TaKe adds this file to the tree IN PLACE OF src/render.cpp and links libtake_gpu.so
// (include/take_gpu.h).  Everything the reference does before the hot path is untouched: main() (src/main.cpp),
// the argument convention of render() (src/render.cpp:14-23), parse_scene() and the whole Mitsuba-XML / PLY / OBJ /
// texture front end, Image3 and imwrite().  What changes is everything after src/render.cpp:28: the `Scene`
// aggregate (src/scene.h:13-33) is flattened into a TakeSceneDesc -- this is the only code that looks inside the
// std::variant types -- and build_bvh + the tile loop + the integrators run on the GPU.
//
// Additive flags (the reference ignores unknown params): -integrator mis|raw|one_sample_mis   -seed N   -device D   -gpus N
//   -gpu_exr            the output step runs on the GPU too (take_gpu_render_to_exr writes ./image.exr)
//   -ref_parse          parse with the reference's parse_scene() and flatten its Scene (the default is load_scene_fast below)
//   -dump_scene FILE    write the flat scene as a TAKESCN1 file and return without rendering (needs no GPU)
// It is compiled against the reference's headers; it contains no reference code.
#include <cstring>
#include <filesystem>
#include <iostream>
#include <map>
#include <string>
#include <vector>

#include "3rdparty/pugixml.hpp"
#include "parse/parse_scene.h"
#include "render.h"
#include "scene.h"
#include "take_gpu.h"
#include "transform.h"
#include "utils/flexception.h"
#include "utils/timer.h"

// The reference parser's own building blocks (non-static functions of src/parse/parse_scene.cpp, declared here because
// parse_scene.h only exports the top-level entry): load_scene_fast() below walks the <scene> element the way
// parse_scene(pugi::xml_node) does (parse_scene.cpp:950-1025) and calls these for everything except the mesh payloads.
using DefaultMap = std::map<std::string, std::string>;
void parse_default_map(pugi::xml_node node, DefaultMap &default_map);
std::string parse_string(const std::string &value, const DefaultMap &default_map);
Real parse_float(const std::string &value, const DefaultMap &default_map);
bool parse_boolean(const std::string &value, const DefaultMap &default_map);
Matrix4x4 parse_transform(pugi::xml_node node, const DefaultMap &default_map);
Vector3 parse_intensity(pugi::xml_node node, const DefaultMap &default_map);
std::tuple<Camera, std::string, int> parse_sensor(pugi::xml_node node, const DefaultMap &default_map);
Texture parse_texture(pugi::xml_node node, const DefaultMap &default_map, TexturePool &texture_pool);
std::tuple<std::string, Material> parse_bsdf(pugi::xml_node node, std::map<std::string, Texture> &texture_map, TexturePool &texture_pool,
                                             const DefaultMap &default_map, const std::string &parent_id);
Light parse_emitter(pugi::xml_node node, const DefaultMap &default_map);
void parse_shape(pugi::xml_node node, std::vector<Material> &materials, std::map<std::string, int> &material_map,
                 std::map<std::string, Texture> &texture_map, TexturePool &texture_pool, std::vector<Light> &lights,
                 std::vector<Shape> &shapes, std::vector<TriangleMesh> &meshes, const DefaultMap &default_map);

namespace {

void fill_texture(TakeMaterialDesc &m, const Texture &t) {
    m.tex_id = -1;
    m.uscale = m.vscale = 1;
    m.uoffset = m.voffset = 0;
    if (auto *c = std::get_if<ConstTexture>(&t)) {
        m.color[0] = c->value.x; m.color[1] = c->value.y; m.color[2] = c->value.z;
    } else if (auto *i = std::get_if<ImageTexture>(&t)) {
        m.tex_id = i->texture_id;
        m.uscale = i->uscale; m.vscale = i->vscale; m.uoffset = i->uoffset; m.voffset = i->voffset;
    }
}

struct Flattened {  // owns the arrays a TakeSceneDesc points into
    std::vector<double> positions, normals, uvs, spheres;
    std::vector<int32_t> indices, prim_material, prim_light;
    std::vector<uint8_t> prim_flags;
    std::vector<TakeMaterialDesc> materials;
    std::vector<TakeLightDesc> lights;
    std::vector<TakeTextureDesc> textures;
    TakeSceneDesc desc{};
    // load_scene_fast: geometry and lights live in the library's builder, textures in this pool
    TakeDescBuilder *builder = nullptr;
    TexturePool pool;
    int spp = 16;
    ~Flattened() { if (builder) take_gpu_builder_destroy(builder); }
};

void fill_materials(const std::vector<Material> &mats, Flattened &f) {
    f.materials.resize(mats.size());
    for (size_t i = 0; i < mats.size(); ++i) {
        TakeMaterialDesc &m = f.materials[i];
        std::memset(&m, 0, sizeof(m));
        const Material &mat = mats[i];
        m.type = (int32_t)mat.index();  // TAKE_MAT_* follow the variant order (src/material.h:82-93)
        m.tex_id = -1;
        m.uscale = m.vscale = 1;
        std::visit([&](const auto &v) {
            using T = std::decay_t<decltype(v)>;
            if constexpr (!std::is_same_v<T, DisneyClearcoat>) fill_texture(m, v.reflectance);
            if constexpr (std::is_same_v<T, Mirror> || std::is_same_v<T, Plastic>) m.p[0] = v.eta;
            if constexpr (std::is_same_v<T, Phong> || std::is_same_v<T, BlinnPhong> || std::is_same_v<T, BlinnPhongMicrofacet>)
                m.p[0] = v.exponent;
            if constexpr (std::is_same_v<T, DisneyDiffuse>) { m.p[0] = v.roughness; m.p[1] = v.subsurface; }
        }, mat);
    }
}

void fill_camera(const Camera &c, const Vector3 &background, TakeSceneDesc &d) {
    d.camera.width = c.width; d.camera.height = c.height;
    for (int k = 0; k < 3; ++k) { d.camera.lookfrom[k] = c.lookfrom[k]; d.camera.lookat[k] = c.lookat[k]; d.camera.up[k] = c.up[k]; }
    d.camera.vfov = c.vfov;
    for (int k = 0; k < 3; ++k) d.background[k] = background[k];
}

// Scene loading without the per-triangle std::variant expansion.  The <scene> element is walked exactly
// like parse_scene(pugi::xml_node) (parse_scene.cpp:950-1025), with the reference's own functions for the sensor, BSDFs,
// textures, emitters and transforms; what differs is where the geometry goes: a "ply" shape is handed to
// take_gpu_builder_add_ply (file -> flat arrays on all host threads, vertex normals included), a sphere to
// take_gpu_builder_add_sphere, and any other shape type is parsed by the reference's parse_shape into a scratch mesh whose
// arrays are appended as they are.  Primitive, light and vertex numbering is the parser's (shapes in document order), and
// the arrays are bit-identical to flatten(parse_scene(file)) -- tests/test_adapter_fast_load.py compares the two dumps.
void load_scene_fast(const std::string &file, Flattened &f) {
    namespace fs = std::filesystem;
    pugi::xml_document doc;
    if (!doc.load_file(file.c_str())) Error("Parse error");
    const fs::path old_path = fs::current_path();
    fs::current_path(fs::path(file).parent_path());   // parse_scene.cpp:1036-1040: file names are relative to the scene
    if (take_gpu_builder_create(&f.builder) != TAKE_OK) Error(std::string("take_gpu: ") + take_gpu_last_error());
    auto check = [](int rc) { if (rc != TAKE_OK) Error(std::string("take_gpu: ") + take_gpu_last_error()); };
    Camera camera{256, 256, Vector3{0, 0, 0}, Vector3{0, 0, -1}, Vector3{0, 1, 0}, Real(45)};  // parse_scene.cpp:951-957, :19-21
    std::vector<Material> materials;
    DefaultMap default_map;
    std::map<std::string, Texture> texture_map;
    std::map<std::string, int> material_map;
    Vector3 background{0.5, 0.5, 0.5};
    std::string out_name;
    for (auto child : doc.child("scene").children()) {
        const std::string name = child.name();
        if (name == "default") {
            parse_default_map(child, default_map);
        } else if (name == "sensor") {
            std::tie(camera, out_name, f.spp) = parse_sensor(child, default_map);
        } else if (name == "bsdf") {
            auto [id, m] = parse_bsdf(child, texture_map, f.pool, default_map, "");
            if (!id.empty()) { material_map[id] = (int)materials.size(); materials.push_back(m); }
        } else if (name == "emitter") {
            const Light l = parse_emitter(child, default_map);
            const PointLight &p = std::get<PointLight>(l);   // the only kind parse_emitter returns
            check(take_gpu_builder_add_point_light(f.builder, &p.intensity.x, &p.position.x));
        } else if (name == "texture") {
            const std::string id = child.attribute("id").value();
            if (texture_map.find(id) != texture_map.end()) Error(std::string("Duplicated texture ID:") + id);
            texture_map[id] = parse_texture(child, default_map, f.pool);
        } else if (name == "background") {
            for (auto g : child.children())
                if (std::string(g.attribute("name").value()) == "radiance") background = parse_intensity(g, default_map);
        } else if (name == "shape") {
            const std::string type = child.attribute("type").value();
            bool is_emitter = false;
            Vector3 radiance{1, 1, 1};
            for (auto c : child.children())
                if (std::string(c.name()) == "emitter") {   // parse_scene.cpp:768-781
                    for (auto g : c.children())
                        if (std::string(g.attribute("name").value()) == "radiance") radiance = parse_intensity(g, default_map);
                    is_emitter = true;
                }
            const double *rad = is_emitter ? &radiance.x : nullptr;
            if (type == "ply" || type == "sphere") {
                int material_id = -1;   // parse_scene.cpp:739-765
                for (auto c : child.children()) {
                    const std::string cn = c.name();
                    if (cn == "ref") {
                        if (c.attribute("id").empty()) Error("Material reference id not specified.");
                        auto it = material_map.find(c.attribute("id").value());
                        if (it == material_map.end()) Error(std::string("Material reference ") + c.attribute("id").value() + " not found.");
                        material_id = it->second;
                    } else if (cn == "bsdf") {
                        auto [id, m] = parse_bsdf(c, texture_map, f.pool, default_map, "");
                        if (!id.empty()) material_map[id] = (int)materials.size();
                        material_id = (int)materials.size();
                        materials.push_back(m);
                    }
                }
                if (type == "sphere") {   // parse_scene.cpp:785-808
                    Vector3 center{0, 0, 0};
                    Real radius = 1;
                    for (auto c : child.children()) {
                        const std::string an = c.attribute("name").value();
                        if (an == "center")
                            center = Vector3{parse_float(c.attribute("x").value(), default_map), parse_float(c.attribute("y").value(), default_map),
                                             parse_float(c.attribute("z").value(), default_map)};
                        else if (an == "radius") radius = parse_float(c.attribute("value").value(), default_map);
                    }
                    check(take_gpu_builder_add_sphere(f.builder, &center.x, radius, material_id, rad));
                } else {                  // parse_scene.cpp:835-861
                    std::string filename;
                    Matrix4x4 to_world = Matrix4x4::identity();
                    bool face_normals = false;
                    for (auto c : child.children()) {
                        const std::string an = c.attribute("name").value();
                        if (an == "filename") filename = parse_string(c.attribute("value").value(), default_map);
                        else if ((an == "toWorld" || an == "to_world") && std::string(c.name()) == "transform") to_world = parse_transform(c, default_map);
                        else if (an == "faceNormals" || an == "face_normals") face_normals = parse_boolean(c.attribute("value").value(), default_map);
                    }
                    const Matrix4x4 inv = inverse(to_world);   // the reference's own inverse (parse_ply.cpp:72)
                    check(take_gpu_builder_add_ply(f.builder, filename.c_str(), &to_world.data[0][0], &inv.data[0][0], material_id,
                                                   face_normals ? 1 : 0, rad));
                }
            } else {
                // obj / serialized / rectangle: the reference's own parse_shape into scratch containers; its mesh goes in as it is
                std::vector<Light> l_scratch;
                std::vector<Shape> s_scratch;
                std::vector<TriangleMesh> m_scratch;
                parse_shape(child, materials, material_map, texture_map, f.pool, l_scratch, s_scratch, m_scratch, default_map);
                for (const TriangleMesh &mesh : m_scratch)
                    check(take_gpu_builder_add_mesh(f.builder, (int64_t)mesh.positions.size(), (const double *)mesh.positions.data(),
                                                    mesh.normals.empty() ? nullptr : (const double *)mesh.normals.data(),
                                                    mesh.uvs.empty() ? nullptr : (const double *)mesh.uvs.data(),
                                                    (int64_t)mesh.indices.size(), (const int32_t *)mesh.indices.data(), mesh.material_id, 0, rad));
            }
        }
    }
    fs::current_path(old_path);
    fill_materials(materials, f);
    for (const Image3 &img : f.pool.image3s)
        f.textures.push_back({img.width, img.height, (const double *)img.data.data()});  // Vector3 = 3 doubles
    TakeSceneDesc &d = f.desc;
    check(take_gpu_builder_finish(f.builder, &d));
    fill_camera(camera, background, d);
    d.num_materials = (int32_t)f.materials.size(); d.num_textures = (int32_t)f.textures.size();
    d.materials = f.materials.data(); d.textures = f.textures.data();
}

void flatten(const Scene &sc, Flattened &f) {
    std::vector<int64_t> base(sc.meshes.size());
    for (size_t m = 0; m < sc.meshes.size(); ++m) {
        const TriangleMesh &mesh = sc.meshes[m];
        base[m] = (int64_t)f.positions.size() / 3;
        const bool hn = !mesh.normals.empty(), hu = !mesh.uvs.empty();
        for (size_t i = 0; i < mesh.positions.size(); ++i) {
            const Vector3 &p = mesh.positions[i];
            f.positions.insert(f.positions.end(), {p.x, p.y, p.z});
            const Vector3 n = hn ? mesh.normals[i] : Vector3{0, 0, 0};
            f.normals.insert(f.normals.end(), {n.x, n.y, n.z});
            const Vector2 t = hu ? mesh.uvs[i] : Vector2{0, 0};
            f.uvs.insert(f.uvs.end(), {t.x, t.y});
        }
    }
    for (const Shape &s : sc.shapes) {  // primitive id == index into scene.shapes
        if (auto *tri = std::get_if<Triangle>(&s)) {
            const TriangleMesh &mesh = sc.meshes[tri->mesh_id];
            const Vector3i id = mesh.indices[tri->face_id];
            const int32_t b = (int32_t)base[tri->mesh_id];
            f.indices.insert(f.indices.end(), {b + id.x, b + id.y, b + id.z});
            f.prim_material.push_back(mesh.material_id);  // what intersect_op reports (src/shape.cpp:85)
            f.prim_light.push_back(tri->area_light_id);
            f.prim_flags.push_back((uint8_t)((mesh.normals.empty() ? 0 : TAKE_PRIM_HAS_NORMALS) |
                                             (mesh.uvs.empty() ? 0 : TAKE_PRIM_HAS_UVS)));
        } else if (auto *sp = std::get_if<Sphere>(&s)) {
            f.indices.insert(f.indices.end(), {(int32_t)(f.spheres.size() / 4), 0, 0});
            f.spheres.insert(f.spheres.end(), {sp->center.x, sp->center.y, sp->center.z, sp->radius});
            f.prim_material.push_back(sp->material_id);
            f.prim_light.push_back(sp->area_light_id);
            f.prim_flags.push_back(TAKE_PRIM_SPHERE);
        }
    }
    fill_materials(sc.materials, f);
    f.lights.resize(sc.lights.size());
    for (size_t i = 0; i < sc.lights.size(); ++i) {
        TakeLightDesc &l = f.lights[i];
        std::memset(&l, 0, sizeof(l));
        if (auto *a = std::get_if<DiffuseAreaLight>(&sc.lights[i])) {
            l.kind = TAKE_LIGHT_AREA; l.prim_id = a->shape_id;
            l.intensity[0] = a->intensity.x; l.intensity[1] = a->intensity.y; l.intensity[2] = a->intensity.z;
        } else if (auto *p = std::get_if<PointLight>(&sc.lights[i])) {
            l.kind = TAKE_LIGHT_POINT; l.prim_id = -1;
            l.intensity[0] = p->intensity.x; l.intensity[1] = p->intensity.y; l.intensity[2] = p->intensity.z;
            l.position[0] = p->position.x; l.position[1] = p->position.y; l.position[2] = p->position.z;
        }
    }
    for (const Image3 &img : sc.textures.image3s)
        f.textures.push_back({img.width, img.height, (const double *)img.data.data()});  // Vector3 = 3 doubles

    TakeSceneDesc &d = f.desc;
    fill_camera(sc.camera, sc.background_color, d);
    f.spp = sc.options.spp;
    d.num_vertices = (int64_t)f.positions.size() / 3;
    d.positions = f.positions.data(); d.normals = f.normals.data(); d.uvs = f.uvs.data();
    d.num_prims = (int64_t)f.prim_material.size();
    d.indices = f.indices.data(); d.prim_material = f.prim_material.data(); d.prim_light = f.prim_light.data();
    d.prim_flags = f.prim_flags.data();
    d.num_spheres = (int64_t)f.spheres.size() / 4; d.spheres = f.spheres.data();
    d.num_materials = (int32_t)f.materials.size(); d.num_textures = (int32_t)f.textures.size();
    d.num_lights = (int32_t)f.lights.size();
    d.materials = f.materials.data(); d.textures = f.textures.data(); d.lights = f.lights.data();
}

}  // namespace

Image3 render(const std::vector<std::string> &params) {
    if (params.size() < 1) return Image3(0, 0);
    int max_depth = 50, device = 0, gpus = 1, integrator = TAKE_INTEGRATOR_MIS;  // render.cpp:14,76
    uint64_t seed = 0;
    bool gpu_exr = false;  // -gpu_exr: the output step runs on the GPU too (take_gpu_render_to_exr writes ./image.exr)
    bool ref_parse = false;
    std::string filename, dump_scene;
    for (int i = 0; i < (int)params.size(); i++) {
        if (params[i] == "-max_depth") max_depth = std::stoi(params[++i]);
        else if (params[i] == "-device") device = std::stoi(params[++i]);
        else if (params[i] == "-gpus") gpus = std::stoi(params[++i]);
        else if (params[i] == "-seed") seed = std::stoull(params[++i]);
        else if (params[i] == "-gpu_exr") gpu_exr = true;
        else if (params[i] == "-ref_parse") ref_parse = true;
        else if (params[i] == "-dump_scene") dump_scene = params[++i];
        else if (params[i] == "-integrator") {
            const std::string v = params[++i];
            integrator = v == "raw" ? TAKE_INTEGRATOR_RAW : v == "one_sample_mis" ? TAKE_INTEGRATOR_ONE_SAMPLE_MIS : TAKE_INTEGRATOR_MIS;
        } else if (filename.empty()) filename = params[i];
    }
    Timer timer;
    std::cout << "Parsing and constructing scene " << params[0] << "." << std::endl;
    tick(timer);
    Flattened flat;
    if (ref_parse) {
        Scene scene = parse_scene(params[0]);
        flatten(scene, flat);   // (copies everything out of `scene` except the texels: keep those alive)
        flat.pool = std::move(scene.textures);
        flat.textures.clear();
        for (const Image3 &img : flat.pool.image3s) flat.textures.push_back({img.width, img.height, (const double *)img.data.data()});
        flat.desc.textures = flat.textures.data();
    } else {
        load_scene_fast(params[0], flat);
    }
    std::cout << "Scene parsing done. Took " << tick(timer) << " seconds." << std::endl;
    if (!dump_scene.empty()) {
        if (take_gpu_scene_desc_save(&flat.desc, flat.spp, dump_scene.c_str()) != TAKE_OK) Error(std::string("take_gpu: ") + take_gpu_last_error());
        return Image3(0, 0);   // imwrite returns at once for an empty image (src/image.cpp:136-138)
    }
    struct { Camera camera; struct { int spp; } options; } scene{};
    scene.camera.width = flat.desc.camera.width; scene.camera.height = flat.desc.camera.height;
    scene.options.spp = flat.spp;
    const Camera &cam0 = scene.camera;
    if (gpus > 1) {  // scene replicated on GPUs device .. device+gpus-1, samples sharded, one NCCL reduce
        std::vector<int> devs(gpus);
        for (int i = 0; i < gpus; ++i) devs[i] = device + i;
        std::vector<double> sum((size_t)cam0.width * cam0.height * 3);
        TakeRenderOpts opts{};
        opts.integrator = integrator; opts.max_depth = max_depth; opts.spp_begin = 0; opts.spp_end = scene.options.spp; opts.seed = seed;
        TakeStats stats{};
        std::cout << "Building BVH and rendering on " << gpus << " GPUs..." << std::endl;
        tick(timer);
        if (take_gpu_render_multi(gpus, devs.data(), &flat.desc, &opts, sum.data(), nullptr, &stats) != TAKE_OK)
            Error(std::string("take_gpu: ") + take_gpu_last_error());
        Image3 img(cam0.width, cam0.height);
        const Real inv = Real(1) / Real(scene.options.spp);
        for (size_t i = 0; i < img.data.size(); ++i) img.data[i] = Vector3{sum[3 * i], sum[3 * i + 1], sum[3 * i + 2]} * inv;
        std::cout << std::endl << "Finish building rendering. Took " << tick(timer) << " seconds." << std::endl;
        return img;
    }
    std::cout << "Building BVH..." << std::endl;
    tick(timer);
    TakeScene *gpu = nullptr;
    if (take_gpu_scene_create(device, &flat.desc, &gpu) != TAKE_OK) Error(std::string("take_gpu: ") + take_gpu_last_error());
    std::cout << "Finish building BVH. Took " << tick(timer) << " seconds." << std::endl;

    const Camera &cam = scene.camera;
    Image3 img(cam.width, cam.height);
    std::vector<double> sum((size_t)cam.width * cam.height * 3);
    TakeRenderOpts opts{};
    opts.integrator = integrator;
    opts.max_depth = max_depth;
    opts.spp_begin = 0;
    opts.spp_end = scene.options.spp;
    opts.seed = seed;
    std::cout << "Rendering..." << std::endl;
    tick(timer);
    TakeStats stats{};
    if (gpu_exr) {
        // render + `color / spp` + double->float->half + B,G,R planes + ZIP pre-filter on the device, deflate on all host
        // threads: the FP64 sums never cross PCIe.  main.cpp:23 then calls imwrite on the image returned here, which
        // returns at once for an empty image (src/image.cpp:136-138), so the stock main() needs no change.
        if (take_gpu_render_to_exr(gpu, &opts, "image.exr", &stats) != TAKE_OK) Error(std::string("take_gpu: ") + take_gpu_last_error());
        std::cout << std::endl << "Finish building rendering. Took " << tick(timer) << " seconds." << std::endl;
        take_gpu_scene_destroy(gpu);
        return Image3(0, 0);
    }
    if (take_gpu_render(gpu, &opts, sum.data(), nullptr, &stats) != TAKE_OK) Error(std::string("take_gpu: ") + take_gpu_last_error());
    const Real inv = Real(1) / Real(scene.options.spp);
    for (size_t i = 0; i < img.data.size(); ++i)  // img(x, H-y-1) = color / spp (render.cpp:78); sums are already in image layout
        img.data[i] = Vector3{sum[3 * i], sum[3 * i + 1], sum[3 * i + 2]} * inv;
    std::cout << std::endl << "Finish building rendering. Took " << tick(timer) << " seconds." << std::endl;
    std::cout << "take_gpu: " << (stats.extend_rays + stats.shadow_rays) << " rays, " << stats.samples << " samples, device "
              << stats.ms_total << " ms" << std::endl;
    take_gpu_scene_destroy(gpu);
    return img;
}
