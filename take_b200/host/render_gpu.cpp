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
// It is compiled against the reference's headers; it contains no reference code.
#include <cstring>
#include <iostream>
#include <string>
#include <vector>

#include "parse/parse_scene.h"
#include "render.h"
#include "scene.h"
#include "take_gpu.h"
#include "utils/flexception.h"
#include "utils/timer.h"

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
};

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
    f.materials.resize(sc.materials.size());
    for (size_t i = 0; i < sc.materials.size(); ++i) {
        TakeMaterialDesc &m = f.materials[i];
        std::memset(&m, 0, sizeof(m));
        const Material &mat = sc.materials[i];
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
    const Camera &c = sc.camera;
    d.camera.width = c.width; d.camera.height = c.height;
    for (int k = 0; k < 3; ++k) { d.camera.lookfrom[k] = c.lookfrom[k]; d.camera.lookat[k] = c.lookat[k]; d.camera.up[k] = c.up[k]; }
    d.camera.vfov = c.vfov;
    for (int k = 0; k < 3; ++k) d.background[k] = sc.background_color[k];
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
    std::string filename;
    for (int i = 0; i < (int)params.size(); i++) {
        if (params[i] == "-max_depth") max_depth = std::stoi(params[++i]);
        else if (params[i] == "-device") device = std::stoi(params[++i]);
        else if (params[i] == "-gpus") gpus = std::stoi(params[++i]);
        else if (params[i] == "-seed") seed = std::stoull(params[++i]);
        else if (params[i] == "-gpu_exr") gpu_exr = true;
        else if (params[i] == "-integrator") {
            const std::string v = params[++i];
            integrator = v == "raw" ? TAKE_INTEGRATOR_RAW : v == "one_sample_mis" ? TAKE_INTEGRATOR_ONE_SAMPLE_MIS : TAKE_INTEGRATOR_MIS;
        } else if (filename.empty()) filename = params[i];
    }
    Timer timer;
    std::cout << "Parsing and constructing scene " << params[0] << "." << std::endl;
    tick(timer);
    Scene scene = parse_scene(params[0]);
    std::cout << "Scene parsing done. Took " << tick(timer) << " seconds." << std::endl;

    Flattened flat;
    flatten(scene, flat);
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
