// render.cpp — packing of the description into the C-ABI arrays, and the `Sampler` mirror
// (src/sampler.rs:11-99) over libmrt.so.  See render.hpp.
#include "render.hpp"

#include <cstring>

namespace mrt_host {

mrt_frame Frame::pack() const {
    mrt_frame f{};
    f.res[0] = res[0]; f.res[1] = res[1];
    f.ssaa = ssaa;
    for (int i = 0; i < 3; i++) f.cam_pos[i] = cam.pos[i];
    for (int i = 0; i < 4; i++) f.cam_dir[i] = cam.dir[i];
    f.fov = cam.fov; f.gamma = cam.gamma; f.exp = cam.exp; f.aprt = cam.aprt; f.foc = cam.foc;
    return f;
}

PackedScene::PackedScene(const Scene& scene) {
    static const std::vector<Renderer> no_objs;
    static const std::vector<Light> no_lights;
    const auto& objs = scene.renderer ? *scene.renderer : no_objs;
    const auto& lts = scene.light ? *scene.light : no_lights;

    std::vector<TexturePtr> uniq;
    auto tex_id = [&](const TexturePtr& t) -> int32_t {
        if (!t) return -1;
        for (size_t k = 0; k < uniq.size(); k++)
            if (uniq[k] == t || *uniq[k] == *t) return (int32_t)k;
        uniq.push_back(t);
        return (int32_t)(uniq.size() - 1);
    };

    uint32_t ii = 0;
    for (const Renderer& o : objs) {
        mrt_object po{};
        po.kind = (uint32_t)o.kind;
        switch (o.kind) {
            case Kind::Sphere: po.param[0] = o.r; break;
            case Kind::Plane: for (int i = 0; i < 3; i++) po.param[i] = o.n[i]; break;
            case Kind::Box: for (int i = 0; i < 3; i++) po.param[i] = o.sizes[i]; break;
            case Kind::Triangle: for (int i = 0; i < 9; i++) po.param[i] = o.vtx[i]; break;
            case Kind::Mesh: {
                po.mesh = (uint32_t)meshes.size();
                mrt_mesh m{(uint32_t)(triangles.size() / 9), (uint32_t)(o.mesh.size() / 9)};
                meshes.push_back(m);
                triangles.insert(triangles.end(), o.mesh.begin(), o.mesh.end());
                break;
            }
        }
        po.first_inst = ii;
        po.n_inst = (uint32_t)o.instance.size();
        for (const Instance& in : o.instance) {
            mrt_instance pi{};
            for (int i = 0; i < 3; i++) pi.pos[i] = in.pos[i];
            for (int i = 0; i < 4; i++) pi.dir[i] = in.dir[i];
            instances.push_back(pi);
            ii++;
        }
        const Material& m = o.mat;
        for (int i = 0; i < 3; i++) po.mat.albedo[i] = m.albedo[i];
        po.mat.rough = m.rough; po.mat.metal = m.metal; po.mat.glass = m.glass;
        po.mat.opacity = m.opacity; po.mat.emit = m.emit;
        po.mat.tex = tex_id(m.tex); po.mat.rmap = tex_id(m.rmap); po.mat.mmap = tex_id(m.mmap);
        po.mat.gmap = tex_id(m.gmap); po.mat.omap = tex_id(m.omap); po.mat.emap = tex_id(m.emap);
        objects.push_back(po);
    }
    uint64_t off = 0;
    for (const TexturePtr& t : uniq) {
        mrt_texture pt{};
        pt.w = t->w; pt.h = t->h; pt.first_texel = off; pt.has_dat = t->has_dat ? 1u : 0u;
        if (t->has_dat) {
            texels.insert(texels.end(), t->dat.begin(), t->dat.end());
            off += t->dat.size() / 3;
        }
        textures.push_back(pt);
    }
    for (const Light& l : lts) {
        mrt_light pl{};
        pl.kind = l.kind;
        for (int i = 0; i < 3; i++) { pl.v[i] = l.v[i]; pl.color[i] = l.color[i]; }
        pl.pwr = l.pwr;
        lights.push_back(pl);
    }
    c_.objects = objects.data();     c_.n_objects = (uint32_t)objects.size();
    c_.instances = instances.data(); c_.n_instances = (uint32_t)instances.size();
    c_.textures = textures.data();   c_.n_textures = (uint32_t)textures.size();
    c_.texels = texels.data();       c_.n_texels = off;
    c_.meshes = meshes.data();       c_.n_meshes = (uint32_t)meshes.size();
    c_.triangles = triangles.data(); c_.n_triangles = (uint32_t)(triangles.size() / 9);
    c_.lights = lights.data();       c_.n_lights = (uint32_t)lights.size();
    for (int i = 0; i < 3; i++) c_.sky_color[i] = scene.sky.color[i];
    c_.sky_pwr = scene.sky.pwr;
}

size_t PackedScene::nbytes() const {
    return objects.size() * sizeof(mrt_object) + instances.size() * sizeof(mrt_instance) + textures.size() * sizeof(mrt_texture) +
           texels.size() * 4 + meshes.size() * sizeof(mrt_mesh) + triangles.size() * 4 + lights.size() * sizeof(mrt_light) + sizeof(mrt_scene);
}

template <class T>
static void append(std::string& s, const std::vector<T>& v) {
    if (!v.empty()) s.append(reinterpret_cast<const char*>(v.data()), v.size() * sizeof(T));
}
std::string PackedScene::bytes() const {
    std::string s;
    uint32_t counts[8] = {c_.n_objects, c_.n_instances, c_.n_textures, (uint32_t)c_.n_texels, c_.n_meshes, c_.n_triangles, c_.n_lights, 0};
    s.append(reinterpret_cast<const char*>(counts), sizeof counts);
    append(s, objects); append(s, instances); append(s, textures); append(s, texels);
    append(s, meshes); append(s, triangles); append(s, lights);
    s.append(reinterpret_cast<const char*>(c_.sky_color), 12);
    s.append(reinterpret_cast<const char*>(&c_.sky_pwr), 4);
    return s;
}

// ----------------------------------------------------------------------------- Sampler
int cuda_device_count() {
    int n = 0;
    mrt_device_count(&n);
    return n;
}

Sampler::Sampler(uint32_t workers, uint32_t n_dim, int device, uint64_t seed) : seed_(seed) {
    if (mrt_abi_version() != MRT_ABI_VERSION) throw Error("libmrt.so ABI version mismatch");
    if (mrt_create(&ctx_, device, workers, n_dim)) {
        const char* e = mrt_last_error(nullptr);
        throw Error(e && *e ? e : "mrt_create failed");
    }
}
Sampler::Sampler(const std::vector<int>& devices, uint32_t workers, uint32_t n_dim, uint64_t seed) : seed_(seed) {
    if (mrt_abi_version() != MRT_ABI_VERSION) throw Error("libmrt.so ABI version mismatch");
    if (mrt_create_group(&ctx_, devices.empty() ? nullptr : devices.data(), (int)devices.size(), workers, n_dim)) {
        const char* e = mrt_last_error(nullptr);
        throw Error(e && *e ? e : "mrt_create_group failed");
    }
}
Sampler::~Sampler() {
    if (ctx_) mrt_destroy(ctx_);
}
void Sampler::check(int rc, const char* what) {
    if (rc) {
        const char* e = mrt_last_error(ctx_);
        throw Error(e && *e ? std::string(e) : std::string(what) + " failed");
    }
}
void Sampler::set_option(uint32_t option, uint32_t value) {
    check(mrt_set_option(ctx_, option, value), "mrt_set_option");  // the scene hash covers the options: re-sent when needed
}
void Sampler::bind(const Scene& scene, const Frame& frame, const RayTracer& rt) {
    const PackedScene packed(scene);
    check(mrt_update_scene(ctx_, &packed.c()), "mrt_update_scene");
    const mrt_frame f = frame.pack();
    check(mrt_update_frame(ctx_, &f), "mrt_update_frame");
    have_frame_ = true;
    check(mrt_set_rt(ctx_, rt.bounce, rt.loss, seed_), "mrt_set_rt");
}
double Sampler::execute(const Scene& scene, const Frame& frame, const RayTracer& rt, uint32_t n_passes) {
    bind(scene, frame, rt);
    double sec = 0.0;
    check(mrt_execute(ctx_, n_passes, &sec), "mrt_execute");
    return sec;
}
Image Sampler::img(const Frame& frame) {
    if (!have_frame_) {
        const mrt_frame f = frame.pack();
        check(mrt_update_frame(ctx_, &f), "mrt_update_frame");
        have_frame_ = true;
    }
    Image im;
    im.w = frame.res[0]; im.h = frame.res[1];
    im.rgb.resize((size_t)im.w * im.h * 3);
    check(mrt_img(ctx_, im.rgb.data()), "mrt_img");
    return im;
}
uint32_t Sampler::passes() {
    uint32_t nw = 0, nh = 0, p = 0;
    check(mrt_film_size(ctx_, &nw, &nh, &p), "mrt_film_size");
    return p;
}
uint32_t Sampler::n_devices() {
    uint32_t n = 1;
    check(mrt_group_info(ctx_, &n, nullptr), "mrt_group_info");
    return n;
}
void Sampler::sync() { check(mrt_sync(ctx_), "mrt_sync"); }
double Sampler::device_seconds() {
    double t = 0.0;
    check(mrt_device_seconds(ctx_, &t), "mrt_device_seconds");
    return t;
}

}  // namespace mrt_host
