// render.hpp — native host mirror of the reference's data model and of its `Sampler`:
//
//   Render / RayTracer / Frame / Camera / Scene / Renderer / Material / Texture / Light / Sky
//                                                   /root/reference/src/rt.rs:9-190
//   Sampler::new / execute / img                    /root/reference/src/sampler.rs:19, 28, 80
//
// Defaults are serde's (src/parser.rs:188-271).  `PackedScene` flattens a Scene into the plain
// arrays of include/mrt.h; `Sampler` drives the C ABI (libmrt.so, hand-written sm_100a kernels).
// Nothing in this directory computes a pixel, and there is no CPU fallback: without the CUDA
// library or a device `Sampler` throws.
#pragma once
#include <array>
#include <cstdint>
#include <memory>
#include <optional>
#include <string>
#include <vector>

#include "../../include/mrt.h"
#include "json.hpp"

namespace mrt_host {

using Vec3 = std::array<float, 3>;
using Vec4 = std::array<float, 4>;  // (w, x, y, z), lin.rs:20-25

struct Texture {  // rt.rs:81-86
    uint32_t w = 0, h = 0;
    bool has_dat = false;          // dat == None: every fetch returns zero (rt.rs:626)
    std::vector<float> dat;        // w*h RGB triples, row-major
    bool operator==(const Texture& o) const { return w == o.w && h == o.h && has_dat == o.has_dat && dat == o.dat; }
};
using TexturePtr = std::shared_ptr<const Texture>;

struct Material {  // rt.rs:88-103, defaults parser.rs:242-259
    Vec3 albedo{1.f, 1.f, 1.f};
    float rough = 0.f, metal = 0.f, glass = 0.f, opacity = 1.f, emit = 0.f;
    TexturePtr tex, rmap, mmap, gmap, omap, emap;
};

enum class Kind : uint32_t { Sphere = MRT_SPHERE, Plane = MRT_PLANE, Box = MRT_BOX, Triangle = MRT_TRIANGLE, Mesh = MRT_MESH };

struct Instance {  // rt.rs:146-150
    Vec3 pos{0.f, 0.f, 0.f};
    Vec4 dir{-0.f, -0.f, -1.f, -0.f};  // Vec4f::backward(), lin.rs:143-145
};

struct Renderer {  // rt.rs:152-158
    Kind kind = Kind::Sphere;
    float r = 0.f;                    // sphere
    Vec3 n{0.f, 0.f, 0.f};            // plane
    Vec3 sizes{0.f, 0.f, 0.f};        // box
    std::array<float, 9> vtx{};       // triangle
    std::vector<float> mesh;          // mesh: 9 floats per triangle
    Material mat;
    std::vector<Instance> instance;
    std::optional<std::string> name;
};

struct Light {  // rt.rs:160-175, defaults parser.rs:261-271
    uint32_t kind = MRT_LIGHT_POINT;
    Vec3 v{0.f, 0.f, 0.f};  // Point: pos, Dir: dir
    float pwr = 0.5f;
    Vec3 color{1.f, 1.f, 1.f};
};

struct Sky {  // rt.rs:177-181, defaults parser.rs:222-229
    Vec3 color{0.f, 0.f, 0.f};
    float pwr = 0.5f;
};

struct Scene {  // rt.rs:183-190
    std::optional<std::vector<Renderer>> renderer;
    std::optional<std::vector<Light>> light;
    Sky sky;
};

struct Camera {  // rt.rs:63-72, defaults parser.rs:198-210
    Vec3 pos{-0.f, -1.f, -0.f};
    Vec4 dir{0.f, 0.f, 1.f, 0.f};
    float fov = 70.f, gamma = 0.8f, exp = 0.2f, aprt = 0.001f, foc = 100.f;
};

struct Frame {  // rt.rs:74-79, defaults parser.rs:212-220
    std::array<uint16_t, 2> res{1280, 720};
    float ssaa = 1.f;
    Camera cam;
    // (nw, nh) of the supersampled film, sampler.rs:29-30: f32 product, truncated
    std::array<uint32_t, 2> film_size() const {
        return {(uint32_t)((float)res[0] * ssaa), (uint32_t)((float)res[1] * ssaa)};
    }
    mrt_frame pack() const;
};

struct RayTracer {  // rt.rs:16-22, defaults parser.rs:188-196
    uint32_t bounce = 8, sample = 16;
    float loss = 0.15f;
};

struct Render {  // rt.rs:9-14
    RayTracer rt;
    Frame frame;
    Scene scene;
};

// Owns the arrays an `mrt_scene` points into (identical textures are stored once).
class PackedScene {
public:
    explicit PackedScene(const Scene& scene);
    const mrt_scene& c() const { return c_; }
    size_t nbytes() const;
    // the flat arrays as one byte string (tests compare it with the Python host's packing)
    std::string bytes() const;
    std::vector<mrt_object> objects;
    std::vector<mrt_instance> instances;
    std::vector<mrt_texture> textures;
    std::vector<float> texels;
    std::vector<mrt_mesh> meshes;
    std::vector<float> triangles;
    std::vector<mrt_light> lights;

private:
    mrt_scene c_{};
};

struct Image {  // ≙ image::RgbImage
    uint32_t w = 0, h = 0;
    std::vector<uint8_t> rgb;
};

// ≙ `Sampler` of src/sampler.rs.  `workers` / `n_dim` (--worker / --dim, cli.rs:157) are accepted
// and ignored by the library: the CUDA grid replaces the tile pool.
// The reference's call pattern is the intended one: construct once, `execute(scene, frame, rt)` once per
// pass, `img(frame)` whenever an image is wanted.  One-pass calls are queued inside the library and
// rendered in full-length launches; a Sampler over several devices (second constructor: mrt_create_group)
// splits the samples over them and gathers the films over NVLink peer mappings in img().
class Sampler {
public:
    Sampler(uint32_t workers = 24, uint32_t n_dim = 64, int device = 0, uint64_t seed = 0x5EED);
    // one Sampler over several GPUs of the box; an empty list means all of them
    Sampler(const std::vector<int>& devices, uint32_t workers = 24, uint32_t n_dim = 64, uint64_t seed = 0x5EED);
    ~Sampler();
    Sampler(const Sampler&) = delete;
    Sampler& operator=(const Sampler&) = delete;

    // ≙ n_passes × Sampler::execute (sampler.rs:28); returns device seconds (a one-pass call: the amortised time
    // of the launches that finished since the last call).  scene / frame / rt are borrowed for the call as in the
    // reference — nothing is cached on this side: they are packed and handed over every time, and the library
    // re-uploads only when the CONTENT changed (mrt_update_scene / mrt_update_frame).
    double execute(const Scene& scene, const Frame& frame, const RayTracer& rt, uint32_t n_passes = 1);
    Image img(const Frame& frame);          // ≙ Sampler::img (sampler.rs:80-99)
    void set_option(uint32_t option, uint32_t value);
    uint32_t passes();
    uint32_t n_devices();
    void sync();                 // launch what is queued and wait for the device(s)
    double device_seconds();     // CUDA-event seconds of every finished path launch
    mrt_ctx* ctx() { return ctx_; }
    // hand over the three borrows (what execute does first); for hosts that queue work themselves
    void bind(const Scene& scene, const Frame& frame, const RayTracer& rt);

private:
    void check(int rc, const char* what);
    mrt_ctx* ctx_ = nullptr;
    uint64_t seed_;
    bool have_frame_ = false;
};

int cuda_device_count();  // via libmrt.so; 0 without a driver

}  // namespace mrt_host
