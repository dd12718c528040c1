// mrt_oracle.cpp — CPU parity oracle: a literal restatement of the reference's per-pixel
// path-tracing path.  TEST INFRASTRUCTURE ONLY (see mrt_oracle.h): never linked into or
// called by the product library.
//
// Every function cites the reference lines it follows (paths relative to /root/reference).
// Operation order is kept as written there (two successive mat-vecs, `v * (1/len)`
// normalisation, no FMA contraction: build with -ffp-contract=off), so that the oracle is
// the reference's arithmetic and not a "better" tracer.  The only liberties:
//   * rand::thread_rng (unseedable) is replaced by a counter-based RNG keyed by
//     (seed, pixel, global sample, block) so that runs are reproducible and the CUDA path
//     can consume the very same uniforms;
//   * the overscan pixels Sampler::execute traces and never reads (sampler.rs:32-48) are skipped;
//   * places where the reference panics (texture index out of bounds, rt.rs:624) clamp.
// Parity pinned by doc/out0..out4.png only; see the header.

#include "mrt_oracle.h"

#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <memory>
#include <string>
#include <thread>
#include <vector>

namespace {

constexpr float E = 0.0001f;  // rt.rs:7
constexpr float PI = 3.14159265358979323846f;

// ---------------------------------------------------------------- lin.rs
struct V2 { float x, y; };
struct V3 { float x, y, z; };
struct V4 { float w, x, y, z; };  // lin.rs:18-25: stored w,x,y,z

inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }      // lin.rs:211-221
inline V3 operator+(V3 a, float b) { return {a.x + b, a.y + b, a.z + b}; }          // lin.rs:223-233
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }      // lin.rs:247-257
inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }                            // lin.rs:304-314
inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }          // lin.rs:259-264 (`*`)
inline V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }          // lin.rs:266-275
inline V3 operator*(float s, V3 a) { return a * s; }                                // lin.rs:277-282
inline V3 operator/(V3 a, float s) { return a * (1.0f / s); }                       // lin.rs:296-302
inline V3 hadam(V3 a, V3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }           // lin.rs:107-113
inline V3 cross(V3 a, V3 b) {                                                       // lin.rs:52-58
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline float mag(V3 a) { return std::sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }     // lin.rs:60-62
inline V3 norm(V3 a) { return a * (1.0f / mag(a)); }                                // lin.rs:64-66
inline V3 reflect(V3 v, V3 n) { return v - n * (2.0f * dot(v, n)); }                // lin.rs:68-70
inline V3 recip(V3 a) { return {1.0f / a.x, 1.0f / a.y, 1.0f / a.z}; }              // lin.rs:72-78
inline V3 vabs(V3 a) { return {std::fabs(a.x), std::fabs(a.y), std::fabs(a.z)}; }   // lin.rs:80-86
inline V4 operator-(V4 a) { return {-a.w, -a.x, -a.y, -a.z}; }                      // lin.rs:445-456
inline V3 proj(V4 a) { return {a.x, a.y, a.z}; }                                    // lin.rs:147-153

// lin.rs:96-105
inline bool refract(V3 v, float eta, V3 n, V3* out) {
    float c = dot(-n, v);
    float k = 1.0f - (eta * eta) * (1.0f - c * c);
    if (k < 0.0f) return false;
    *out = v * eta + n * (c * eta + std::sqrt(k));
    return true;
}

struct M3 { float m[9]; };
inline V3 operator*(const M3& m, V3 v) {                                            // lin.rs:344-365
    return {m.m[0] * v.x + m.m[1] * v.y + m.m[2] * v.z,
            m.m[3] * v.x + m.m[4] * v.y + m.m[5] * v.z,
            m.m[6] * v.x + m.m[7] * v.y + m.m[8] * v.z};
}
// lin.rs:175-183
inline M3 rotate_y(V4 d) {
    float cw = std::sqrt(1.0f - d.w * d.w);
    return {{cw, 0.0f, d.w, 0.0f, 1.0f, 0.0f, -d.w, 0.0f, cw}};
}
// lin.rs:197-209 (upper-left 3x3 of the Mat4f; Mat4f*Vec3f uses only those, lin.rs:356-365)
inline M3 lookat(V4 d, V3 up) {
    V3 fwd = norm(proj(d));
    V3 right = norm(cross(fwd, up));
    V3 n_up = cross(right, fwd);
    return {{right.x, -right.y, right.z, -fwd.x, fwd.y, -fwd.z, n_up.x, -n_up.y, n_up.z}};
}
const V3 UP = {0.0f, 0.0f, 1.0f};  // lin.rs:48-50

// f32::max / f32::min: if one argument is NaN the other is returned (== fmaxf/fminf).
inline float fmax_(float a, float b) { return std::fmax(a, b); }
inline float fmin_(float a, float b) { return std::fmin(a, b); }

// f32::total_cmp: -1, 0, 1
inline int total_cmp(float a, float b) {
    int32_t x, y;
    std::memcpy(&x, &a, 4);
    std::memcpy(&y, &b, 4);
    x ^= (int32_t)((uint32_t)(x >> 31) >> 1);
    y ^= (int32_t)((uint32_t)(y >> 31) >> 1);
    return (x > y) - (x < y);
}
// Rust `f32 as usize` (saturating, NaN -> 0)
inline uint64_t f32_as_usize(float v) {
    if (!(v > 0.0f)) return 0;  // NaN, negatives, -0
    if (v >= 18446744073709551616.0f) return UINT64_MAX;
    return (uint64_t)v;
}
inline uint8_t f32_as_u8(float v) {
    if (!(v > 0.0f)) return 0;
    if (v >= 255.0f) return 255;
    return (uint8_t)v;
}

// ---------------------------------------------------------------- RNG (stands in for rand 0.8 thread_rng)
// pcg4d (Jarzynski & Olano, "Hash Functions for GPU Rendering", JCGT 2020) as a counter-based
// generator: one call gives the 4 uniforms of one (pixel, sample, block).
inline void pcg4d(uint32_t v[4]) {
    for (int i = 0; i < 4; i++) v[i] = v[i] * 1664525u + 1013904223u;
    v[0] += v[1] * v[3]; v[1] += v[2] * v[0]; v[2] += v[0] * v[1]; v[3] += v[1] * v[2];
    for (int i = 0; i < 4; i++) v[i] ^= v[i] >> 16;
    v[0] += v[1] * v[3]; v[1] += v[2] * v[0]; v[2] += v[0] * v[1]; v[3] += v[1] * v[2];
}
inline uint32_t fold_seed(uint64_t seed) {
    return (uint32_t)seed ^ ((uint32_t)(seed >> 32) * 0x9E3779B9u);
}
struct Block { float u[4]; };
inline Block rng_block(uint32_t pixel, uint32_t sample, uint32_t block, uint32_t key) {
    uint32_t v[4] = {pixel, sample, block, key};
    pcg4d(v);
    Block b;
    // U[0, 1 - 2^-24]: the scale (1 - 2^-24) 2^-32 keeps a word that rounds up to 2^32 below 1
    for (int i = 0; i < 4; i++) b.u[i] = (float)v[i] * 0x1.fffffep-33f;
    return b;
}
// camera block: lowbias32 (Wellons) of the sample index offset by a per-pixel seed, split into
// two 16-bit uniforms (the lens jitter only needs a cheap hash)
inline uint32_t lowbias32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du;
    x ^= x >> 15; x *= 0x846ca68bu;
    x ^= x >> 16;
    return x;
}
inline Block rng_cam(uint32_t pixel, uint32_t sample, uint32_t key) {
    const uint32_t seed = lowbias32(pixel * 0x9E3779B1u + key);
    const uint32_t h = lowbias32(seed + sample * 0x85EBCA6Bu);
    Block b;
    b.u[0] = (float)(h >> 16) * (1.0f / 65536.0f);
    b.u[1] = (float)(h & 0xffffu) * (1.0f / 65536.0f);
    b.u[2] = b.u[3] = 0.0f;
    return b;
}
// block ids: camera = rng_cam (u0,u1 lens); 1+2b = bounce b "A" (u0 reflect lottery, u1,u2 reflect
// direction, u3 emission draw); 2+2b = bounce b "B" (u0 transmission lottery, u1 refract
// lottery, u2,u3 refract direction).
struct PathRng {
    uint32_t pixel, sample, key;
    Block cam() const { return rng_cam(pixel, sample, key); }
    Block a(uint32_t bounce) const { return rng_block(pixel, sample, 1u + 2u * bounce, key); }
    Block b(uint32_t bounce) const { return rng_block(pixel, sample, 2u + 2u * bounce, key); }
};
// gen_bool(p): true with probability p (p == 1 always, p == 0 never)
inline bool bern(float u, float p) { return u < p; }

// ---------------------------------------------------------------- rt.rs data model
struct Ray {  // rt.rs:45-52
    V3 orig, dir;
    float t = 0.0f, pwr = 1.0f;
    uint32_t bounce = 0;
};
inline V3 point(const Ray& r) { return r.orig + r.dir * r.t; }  // rt.rs:193-197

struct Texture { uint32_t w, h; bool has_dat; const float* dat; };  // rt.rs:81-86
struct Material {  // rt.rs:88-103
    V3 albedo;
    float rough, metal, glass, opacity, emit;
    int tex, rmap, mmap, gmap, omap, emap;
};
struct Tri { V3 a, b, c; };  // rt.rs:128-129
struct BVH {  // rt.rs:110-116
    V3 aabb;
    V3 rel_pos;
    bool has_content = false;
    std::vector<uint32_t> content;
    bool has_childs = false;
    std::vector<BVH> childs;
};
struct Mesh { std::vector<Tri> tris; bool has_bvh = false; BVH bvh; };
struct Instance { V3 pos; V4 dir; };  // rt.rs:146-150
struct Object {  // rt.rs:152-158
    uint32_t kind;
    float r;       // Sphere
    V3 v;          // Plane n / Box sizes
    Tri tri;       // Triangle
    int mesh;
    Material mat;
    std::vector<Instance> inst;
};
struct Light { uint32_t kind; V3 v; float pwr; V3 color; };  // rt.rs:170-175
struct Scene {
    std::vector<Object> objs;
    std::vector<Light> lights;
    std::vector<Texture> tex;
    std::vector<float> texels;
    std::vector<Mesh> meshes;
    V3 sky_color{0, 0, 0};
    float sky_pwr = 0.5f;
    uint32_t normal_space = MRT_NORMAL_FORWARD_XF;  // MRT_OPT_NORMAL_SPACE, see include/mrt.h
};
struct Hit {  // rt.rs:54-61
    int obj = -1, inst = -1, idx = -1;
    Ray ray;
    V3 norm{0, 0, 0};
};

// ---------------------------------------------------------------- intersect, rt.rs:299-412
// Box, rt.rs:299-333
inline bool box_intersect(V3 size, const Ray& ray, V3 pos, float* t0, float* t1) {
    V3 m = recip(ray.dir);
    if (std::isinf(m.x)) m.x = 1.0f / E;
    if (std::isinf(m.y)) m.y = 1.0f / E;
    if (std::isinf(m.z)) m.z = 1.0f / E;
    V3 n = hadam(ray.orig - pos, m);
    V3 k = hadam(0.5f * size, vabs(m));
    V3 a = -n - k;
    V3 b = -n + k;
    float ta = fmax_(fmax_(a.x, a.y), a.z);
    float tb = fmin_(fmin_(b.x, b.y), b.z);
    if (ta > tb || tb < 0.0f) return false;
    *t0 = ta;
    *t1 = tb;
    return true;
}
// Sphere, rt.rs:335-359
inline bool sphere_intersect(float r, const Ray& ray, V3 pos, float* t0, float* t1) {
    V3 o = ray.orig - pos;
    float a = dot(ray.dir, ray.dir);
    float b = 2.0f * dot(o, ray.dir);
    float c = dot(o, o) - r * r;
    float disc = b * b - 4.0f * a * c;
    if (disc < 0.0f) return false;
    float ta = (-b - std::sqrt(disc)) / (2.0f * a);
    float tb = (-b + std::sqrt(disc)) / (2.0f * a);
    if (ta < 0.0f) return false;
    *t0 = ta;
    *t1 = tb;
    return true;
}
// Triangle, rt.rs:361-398
inline bool tri_intersect(const Tri& tr, const Ray& ray, V3 pos, float* tout) {
    V3 e0 = tr.b - tr.a;
    V3 e1 = tr.c - tr.a;
    V3 p = cross(ray.dir, e1);
    float d = dot(e0, p);
    if (d < E && d > -E) return false;
    float inv_d = 1.0f / d;
    V3 t = ray.orig - (tr.a + pos);
    float u = dot(t, p) * inv_d;
    if (u < 0.0f || u > 1.0f) return false;
    V3 q = cross(t, e0);
    float v = dot(ray.dir, q) * inv_d;
    if (v < 0.0f || (u + v) > 1.0f) return false;
    float tt = dot(e1, q) * inv_d;
    if (tt < 0.0f) return false;
    *tout = tt;
    return true;
}
// Plane, rt.rs:400-412
inline bool plane_intersect(V3 n, const Ray& ray, V3 pos, float* tout) {
    float d = dot(-norm(n), pos);
    float t = -(dot(ray.orig, norm(n)) + d) / dot(ray.dir, norm(n));
    if (t <= 0.0f) return false;
    *tout = t;
    return true;
}

// ---------------------------------------------------------------- normal, rt.rs:414-466
inline bool in_range(float lo, float hi, float v) { return lo <= v && v < hi; }  // Range::contains
// Box, rt.rs:414-445 (note the missing `else` before the z test at :435)
inline V3 box_normal(V3 size, V3 hit, V3 pos) {
    V3 p = hadam(hit - pos, recip(size) * 2.0f);
    const float pl = 1.0f - E, ph = 1.0f + E, nl = -1.0f - E, nh = -1.0f + E;
    V3 n = {0, 0, 0};
    if (in_range(pl, ph, p.x)) n = {1, 0, 0};
    else if (in_range(nl, nh, p.x)) n = {-1, -0.0f, -0.0f};
    else if (in_range(pl, ph, p.y)) n = {0, 1, 0};
    else if (in_range(nl, nh, p.y)) n = {-0.0f, -1, -0.0f};
    if (in_range(pl, ph, p.z)) n = {0, 0, 1};
    else if (in_range(nl, nh, p.z)) n = {-0.0f, -0.0f, -1};
    return n;
}
// Triangle, rt.rs:459-466
inline V3 tri_normal(const Tri& tr) { return cross(tr.b - tr.a, tr.c - tr.a); }

// ---------------------------------------------------------------- uv, rt.rs:468-548
// Box, rt.rs:468-516
inline V2 box_uv(V3 size, V3 hit, V3 pos) {
    V3 p = hadam(hit - pos, recip(size) * 2.0f);
    const float pl = 1.0f - E, ph = 1.0f + E, nl = -1.0f - E, nh = -1.0f + E;
    if (in_range(pl, ph, p.x)) return {(0.5f + 0.5f * p.y) / 4.0f + 2.0f / 4.0f, (0.5f - 0.5f * p.z) / 3.0f + 1.0f / 3.0f};
    else if (in_range(nl, nh, p.x)) return {(0.5f - 0.5f * p.y) / 4.0f, (0.5f - 0.5f * p.z) / 3.0f + 1.0f / 3.0f};
    else if (in_range(pl, ph, p.y)) return {(0.5f - 0.5f * p.x) / 4.0f + 3.0f / 4.0f, (0.5f - 0.5f * p.z) / 3.0f + 1.0f / 3.0f};
    else if (in_range(nl, nh, p.y)) return {(0.5f + 0.5f * p.x) / 4.0f + 1.0f / 4.0f, (0.5f - 0.5f * p.z) / 3.0f + 1.0f / 3.0f};
    if (in_range(pl, ph, p.z)) return {(0.5f + 0.5f * p.x) / 4.0f + 1.0f / 4.0f, (0.5f - 0.5f * p.y) / 3.0f};
    else if (in_range(nl, nh, p.z)) return {(0.5f + 0.5f * p.x) / 4.0f + 1.0f / 4.0f, (0.5f + 0.5f * p.y) / 3.0f + 2.0f / 3.0f};
    return {0.0f, 0.0f};
}
// Sphere, rt.rs:518-526
inline V2 sphere_uv(V3 hit, V3 pos) {
    V3 v = norm(hit - pos);
    return {0.5f + 0.5f * std::atan2(v.x, -v.y) / PI, 0.5f - 0.5f * v.z};
}
// Plane, rt.rs:528-542
inline float fract(float v) { return v - std::trunc(v); }
inline V2 plane_uv(V3 hit) {
    float x = fract(hit.x + 0.5f);
    if (x < 0.0f) x = 1.0f + x;
    float y = fract(hit.y + 0.5f);
    if (y < 0.0f) y = 1.0f + y;
    return {x, y};
}

// Texture::get_color, rt.rs:618-628.  The reference panics when the index is out of
// bounds; the oracle clamps the final linear index (documented deviation).
inline V3 tex_get(const Scene& sc, int id, V2 uv) {
    const Texture& t = sc.tex[id];
    uint64_t x = f32_as_usize(uv.x * (float)t.w);
    uint64_t y = f32_as_usize(uv.y * (float)t.h);
    if (!t.has_dat || t.w == 0 || t.h == 0) return {0, 0, 0};
    uint64_t n = (uint64_t)t.w * t.h;
    unsigned __int128 idx = (unsigned __int128)y * t.w + x;
    uint64_t i = idx >= n ? n - 1 : (uint64_t)idx;
    return {t.dat[3 * i], t.dat[3 * i + 1], t.dat[3 * i + 2]};
}

// ---------------------------------------------------------------- BVH, rt.rs:222-249, 261-270, 630-703
// Triangle::check_in_aabb, rt.rs:227-248
inline bool tri_in_aabb(const Tri& tr, V3 aabb, V3 rel_pos) {
    V3 v0 = rel_pos + 0.5f * aabb;
    V3 v1 = rel_pos - 0.5f * aabb;
    auto in = [&](V3 v) {
        if (v.x > v0.x || v.y > v0.y || v.z > v0.z) return false;
        if (v.x < v1.x || v.y < v1.y || v.z < v1.z) return false;
        return true;
    };
    return in(tr.a) || in(tr.b) || in(tr.c);
}
// Mesh::gen_aabb, rt.rs:261-270
inline bool mesh_aabb(const Mesh& m, V3* out) {
    if (m.tris.empty()) return false;
    float mx = 0, my = 0, mz = 0;
    bool first = true;
    for (const Tri& t : m.tris)
        for (V3 v : {t.a, t.b, t.c}) {
            if (first) { mx = std::fabs(v.x); my = std::fabs(v.y); mz = std::fabs(v.z); first = false; continue; }
            if (total_cmp(std::fabs(v.x), mx) >= 0) mx = std::fabs(v.x);
            if (total_cmp(std::fabs(v.y), my) >= 0) my = std::fabs(v.y);
            if (total_cmp(std::fabs(v.z), mz) >= 0) mz = std::fabs(v.z);
        }
    *out = {2.0f * mx, 2.0f * my, 2.0f * mz};
    return true;
}
// BVH::construct, rt.rs:631-674; child order = gen_pos, rt.rs:678-689
const V3 GEN_POS[8] = {{1, 1, 1}, {-1, 1, 1}, {-1, -1, 1}, {1, -1, 1}, {1, 1, -1}, {-1, 1, -1}, {-1, -1, -1}, {1, -1, -1}};
BVH bvh_construct(V3 aabb, V3 rel_pos, const std::vector<Tri>& objs, uint32_t d, uint32_t deep) {
    BVH child;
    child.aabb = aabb;
    child.rel_pos = rel_pos;
    if (d >= deep) {
        std::vector<uint32_t> tmp;
        for (uint32_t i = 0; i < objs.size(); i++)
            if (tri_in_aabb(objs[i], child.aabb, child.rel_pos)) tmp.push_back(i);
        if (!tmp.empty()) { child.has_content = true; child.content = std::move(tmp); }
        return child;
    }
    std::vector<BVH> tmp;
    for (int i = 0; i < 8; i++) {
        BVH c = bvh_construct(0.5f * child.aabb, child.rel_pos + hadam(child.aabb, GEN_POS[i] * 0.25f), objs, d + 1, deep);
        if (c.has_content || c.has_childs) tmp.push_back(std::move(c));
    }
    if (!tmp.empty()) { child.has_childs = true; child.childs = std::move(tmp); }
    return child;
}

// ---------------------------------------------------------------- Renderer, rt.rs:706-864
// Renderer::intersect_bvh, rt.rs:707-723
bool intersect_bvh(const Instance& inst, const Ray& ray, const BVH& bvh, std::vector<uint32_t>* out) {
    float t0, t1;
    if (!box_intersect(bvh.aabb, ray, inst.pos + bvh.rel_pos, &t0, &t1)) return false;
    if (bvh.has_content) { out->insert(out->end(), bvh.content.begin(), bvh.content.end()); return true; }
    // a node with neither content nor childs would `unwrap()` a None (rt.rs:717): cannot be
    // built by bvh_construct except as the root of an empty mesh, which set_scene rejects.
    for (const BVH& c : bvh.childs) {
        std::vector<uint32_t> sub;
        if (intersect_bvh(inst, ray, c, &sub)) out->insert(out->end(), sub.begin(), sub.end());
    }
    return true;
}

struct IsectOut { float t0, t1; int idx0, idx1; };

// Renderer::intersect, rt.rs:725-774
bool obj_intersect(const Scene& sc, const Object& o, const Instance& inst, const Ray& ray, IsectOut* out) {
    M3 rot_y = rotate_y(-inst.dir);
    M3 look = lookat(-inst.dir, UP);
    Ray n_ray = ray;
    n_ray.orig = inst.pos + rot_y * (look * (ray.orig - inst.pos));
    n_ray.dir = rot_y * (look * ray.dir);
    float t0, t1;
    switch (o.kind) {
        case MRT_SPHERE:
            if (!sphere_intersect(o.r, n_ray, inst.pos, &t0, &t1)) return false;
            *out = {t0, t1, -1, -1};
            return true;
        case MRT_PLANE:
            if (!plane_intersect(o.v, n_ray, inst.pos, &t0)) return false;
            *out = {t0, t0, -1, -1};
            return true;
        case MRT_BOX:
            if (!box_intersect(o.v, n_ray, inst.pos, &t0, &t1)) return false;
            *out = {t0, t1, -1, -1};
            return true;
        case MRT_TRIANGLE:
            if (!tri_intersect(o.tri, n_ray, inst.pos, &t0)) return false;
            *out = {t0, t0, -1, -1};
            return true;
        case MRT_MESH: {
            const Mesh& mesh = sc.meshes[o.mesh];
            std::vector<uint32_t> idx;
            if (mesh.has_bvh) {
                if (!intersect_bvh(inst, n_ray, mesh.bvh, &idx)) return false;
            } else {
                for (uint32_t i = 0; i < mesh.tris.size(); i++) idx.push_back(i);
            }
            // Vec::dedup: drops consecutive repeats only (rt.rs:756)
            std::vector<uint32_t> ded;
            for (uint32_t i : idx)
                if (ded.empty() || ded.back() != i) ded.push_back(i);
            bool any = false;
            float best0 = 0, best1 = 0;
            int i0 = -1, i1 = -1;
            for (uint32_t i : ded) {
                float t;
                if (!tri_intersect(mesh.tris[i], n_ray, inst.pos, &t)) continue;
                if (!any) { best0 = best1 = t; i0 = i1 = (int)i; any = true; continue; }
                if (total_cmp(t, best0) < 0) { best0 = t; i0 = (int)i; }   // min_by: first minimum (rt.rs:764)
                if (total_cmp(t, best1) >= 0) { best1 = t; i1 = (int)i; }  // max_by: last maximum (rt.rs:765)
            }
            if (!any) return false;
            *out = {best0, best1, i0, i1};
            return true;
        }
    }
    return false;
}

// Renderer::normal, rt.rs:776-793
V3 obj_normal(const Scene& sc, const Object& o, const Instance& inst, const Hit& hit) {
    V3 hit_p = point(hit.ray);
    M3 rot_y = rotate_y(-inst.dir);
    M3 look = lookat(-inst.dir, UP);
    V3 n_hit = inst.pos + rot_y * (look * (hit_p - inst.pos));
    V3 n;
    switch (o.kind) {
        case MRT_SPHERE: n = n_hit - inst.pos; break;              // rt.rs:447-451
        case MRT_PLANE: n = o.v; break;                            // rt.rs:453-457
        case MRT_BOX: n = box_normal(o.v, n_hit, inst.pos); break;
        case MRT_TRIANGLE: n = tri_normal(o.tri); break;
        default: n = tri_normal(sc.meshes[o.mesh].tris[hit.idx]); break;
    }
    // rt.rs:792 at HEAD pushes n through the FORWARD transform again.  doc/out3.png (the README's
    // CornellBox2 render) was produced by a revision that returned the object-space normal: with
    // this switch the oracle reproduces that image (tests/test_oracle_golden.py).
    if (sc.normal_space == MRT_NORMAL_OBJECT) return norm(n);
    return norm(rot_y * (look * n));
}

// Renderer::to_uv, rt.rs:795-809.  Triangle and Mesh are `todo!()` there; set_scene rejects
// textured triangles/meshes so this is never reached for them.
V2 obj_uv(const Object& o, const Instance& inst, V3 hit) {
    M3 rot_y = rotate_y(-inst.dir);
    M3 look = lookat(-inst.dir, UP);
    V3 n_hit = inst.pos + rot_y * (look * (hit - inst.pos));
    switch (o.kind) {
        case MRT_SPHERE: return sphere_uv(n_hit, inst.pos);
        case MRT_PLANE: return plane_uv(n_hit);
        case MRT_BOX: return box_uv(o.v, n_hit, inst.pos);
        default: return {0, 0};
    }
}

// material getters, rt.rs:592-616 + 811-863
struct HitView {
    const Scene& sc;
    const Hit& h;
    const Object& obj() const { return sc.objs[h.obj]; }
    const Instance& inst() const { return obj().inst[h.inst]; }
    V3 texel(int id) const { return tex_get(sc, id, obj_uv(obj(), inst(), point(h.ray))); }
    V3 color() const { const Material& m = obj().mat; return m.tex >= 0 ? hadam(m.albedo, texel(m.tex)) : m.albedo; }
    float rough() const { const Material& m = obj().mat; return m.rmap >= 0 ? texel(m.rmap).x : m.rough; }
    float metal() const { const Material& m = obj().mat; return m.mmap >= 0 ? texel(m.mmap).x : m.metal; }
    float glass() const { const Material& m = obj().mat; return m.gmap >= 0 ? texel(m.gmap).x : m.glass; }
    float opacity() const { const Material& m = obj().mat; return m.omap >= 0 ? texel(m.omap).x : m.opacity; }
    float emit() const { const Material& m = obj().mat; return m.emap >= 0 ? texel(m.emap).x : m.emit; }
};

struct Stats {
    uint64_t paths = 0, segments = 0, hits = 0, shadow_rays = 0, nan_normals = 0;
    uint64_t hist[34] = {0};
    void add(const Stats& o) {
        paths += o.paths; segments += o.segments; hits += o.hits; shadow_rays += o.shadow_rays;
        nan_normals += o.nan_normals;
        for (int i = 0; i < 34; i++) hist[i] += o.hist[i];
    }
};

// RayTracer::closest_hit, rt.rs:867-898
bool closest_hit(const Scene& sc, const Ray& ray, Hit* hit0, Hit* hit1) {
    bool any = false;
    IsectOut best{};
    int bo = -1, bi = -1;
    for (size_t oi = 0; oi < sc.objs.size(); oi++) {
        const Object& o = sc.objs[oi];
        for (size_t ii = 0; ii < o.inst.size(); ii++) {
            IsectOut p;
            if (!obj_intersect(sc, o, o.inst[ii], ray, &p)) continue;
            if (!any || total_cmp(p.t0, best.t0) < 0) {  // min_by keeps the first of equal minima (rt.rs:872)
                best = p; bo = (int)oi; bi = (int)ii; any = true;
            }
        }
    }
    if (!any) return false;
    hit0->obj = bo; hit0->inst = bi; hit0->idx = best.idx0; hit0->ray = ray; hit0->ray.t = best.t0;
    hit0->norm = obj_normal(sc, sc.objs[bo], sc.objs[bo].inst[bi], *hit0);
    hit1->obj = bo; hit1->inst = bi; hit1->idx = best.idx1; hit1->ray = ray; hit1->ray.t = best.t1;
    hit1->norm = obj_normal(sc, sc.objs[bo], sc.objs[bo].inst[bi], *hit1);
    return true;
}

// Ray::cast, rt.rs:551-553 ; cast_default rt.rs:555-557
inline Ray ray_cast(V3 orig, V3 dir, float pwr, uint32_t bounce) {
    Ray r; r.orig = orig + dir * E; r.dir = dir; r.pwr = pwr; r.bounce = bounce; r.t = 0.0f; return r;
}
inline Ray ray_cast_default(V3 orig, V3 dir) { return ray_cast(orig, dir, 1.0f, 0); }

// RayTracer::rand, rt.rs:996-1007
inline V3 rt_rand(V3 n, float r, float u1, float u2) {
    float th = std::acos(1.0f - 2.0f * u1);
    float phi = u2 * 2.0f * PI;
    V3 v = {std::sin(th) * std::cos(phi), std::sin(th) * std::sin(phi), std::cos(th)};
    return norm(n + r * v);
}

struct RtParams { uint32_t bounce; float loss; };

// Ray::reflect, rt.rs:559-572 (self = hit.ray)
Ray ray_reflect(const Scene& sc, const RtParams& rt, const Hit& hit, float u_lot, float u1, float u2) {
    HitView hv{sc, hit};
    float rough = hv.rough();
    float opacity = hv.opacity();
    if (hv.obj().mat.metal == 0.0f && opacity != 0.0f && bern(u_lot, 0.80f)) rough = 1.0f;
    V3 n = rt_rand(hit.norm, rough, u1, u2);
    V3 dir = norm(reflect(hit.ray.dir, n));
    return ray_cast(point(hit.ray), dir, hit.ray.pwr * (1.0f - fmin_(rt.loss, 1.0f)), hit.ray.bounce + 1);
}
// Ray::refract, rt.rs:574-589
bool ray_refract(const Scene& sc, const RtParams& rt, const Hit& hit, float u_lot, float u1, float u2, Ray* out) {
    HitView hv{sc, hit};
    float rough = hv.rough();
    float opacity = hv.opacity();
    if (hv.obj().mat.metal == 0.0f && opacity != 0.0f && bern(u_lot, 0.80f)) rough = 1.0f;
    V3 n = rt_rand(hit.norm, rough, u1, u2);
    float eta = 1.0f + 0.5f * hv.glass();
    V3 d;
    if (!refract(hit.ray.dir, eta, n, &d)) return false;
    d = norm(d);
    *out = ray_cast(point(hit.ray), d, hit.ray.pwr * (1.0f - fmin_(rt.loss, 1.0f)), hit.ray.bounce + 1);
    return true;
}

struct Camera { V3 pos; V4 dir; float fov, gamma, exp, aprt, foc; };
struct Frame { uint16_t res[2]; float ssaa; Camera cam; };

// RayTracer::cast, rt.rs:900-931
Ray rt_cast(V2 uv, const Frame& f, float u1, float u2) {
    float tan_fov = std::tan((0.5f * f.cam.fov) * (PI / 180.0f));
    V3 dir = norm(V3{uv.x, 1.0f / (2.0f * tan_fov), -uv.y});
    Ray ray = ray_cast_default(f.cam.pos, dir);
    ray.t = f.cam.foc;
    V3 p = point(ray);
    V3 pos = {f.cam.pos.x + (u1 - 0.5f) * f.cam.aprt, f.cam.pos.y, f.cam.pos.z + (u2 - 0.5f) * f.cam.aprt};
    V3 new_dir = norm(p - pos);
    M3 look = lookat(f.cam.dir, UP);
    M3 rot_y = rotate_y(f.cam.dir);
    return ray_cast_default(pos, rot_y * (look * new_dir));
}
// RayTracer::iter, rt.rs:937-954
Ray rt_iter(float cx, float cy, const Frame& f, float u1, float u2) {
    float w = (float)f.res[0] * f.ssaa;
    float h = (float)f.res[1] * f.ssaa;
    float aspect = w / h;
    V2 uv = {aspect * (cx - 0.5f * w) / w, (cy - 0.5f * h) / h};
    return rt_cast(uv, f, u1, u2);
}

struct PathItem { Hit hit; std::vector<int> lights; bool has_lights = false; };

// RaytraceIterator::next, rt.rs:1014-1066
bool iter_next(const Scene& sc, const RtParams& rt, Ray* next_ray, const PathRng& rng, PathItem* item, Stats* st) {
    if (next_ray->bounce > rt.bounce) return false;
    uint32_t b = next_ray->bounce;
    Hit h0, h1;
    st->segments++;
    if (!closest_hit(sc, *next_ray, &h0, &h1)) return false;
    st->hits++;
    if (!std::isfinite(h0.norm.x) || !std::isfinite(h0.norm.y) || !std::isfinite(h0.norm.z)) st->nan_normals++;
    item->lights.clear();
    item->has_lights = false;
    for (size_t li = 0; li < sc.lights.size(); li++) {
        const Light& light = sc.lights[li];
        V3 l = light.kind == MRT_LIGHT_POINT ? light.v - point(h0.ray) : -norm(light.v);
        Ray ray_l = ray_cast_default(point(h0.ray), norm(l));
        Hit s0, s1;
        st->shadow_rays++;
        if (closest_hit(sc, ray_l, &s0, &s1)) continue;  // no distance limit (rt.rs:1036)
        item->lights.push_back((int)li);
        item->has_lights = true;
    }
    Block A = rng.a(b);
    *next_ray = ray_reflect(sc, rt, h0, A.u[0], A.u[1], A.u[2]);
    item->hit = h0;
    float opacity = HitView{sc, h0}.opacity();
    float p = fmin_(1.0f - opacity, 0.85f);
    if (p > 0.0f) {  // gen_bool(0) is always false: the B block is drawn only when it can matter
        Block B = rng.b(b);
        if (bern(B.u[0], p)) {
            Ray r;
            if (ray_refract(sc, rt, h1, B.u[1], B.u[2], B.u[3], &r)) { *next_ray = r; item->hit = h1; }
        }
    }
    return true;
}

// direct light of one recorded hit, rt.rs:973-987
V3 direct_light(const Scene& sc, const PathItem& it) {
    V3 l_col = {0, 0, 0};
    if (!it.has_lights) return l_col;
    HitView hv{sc, it.hit};
    for (int li : it.lights) {
        const Light& light = sc.lights[li];
        V3 l = light.kind == MRT_LIGHT_POINT ? light.v - point(it.hit.ray) : -norm(light.v);
        float diff = fmax_(dot(norm(l), it.hit.norm), 0.0f);
        float s = fmax_(dot(it.hit.ray.dir, reflect(norm(l), it.hit.norm)), 0.0f);
        float s2 = s * s, s4 = s2 * s2, s8 = s4 * s4, s16 = s8 * s8, s32 = s16 * s16;  // powi(32)
        float spec = s32 * (1.0f - hv.rough());
        V3 o_col = hv.color() * (1.0f - hv.metal());
        l_col = l_col + (hadam(o_col * diff, light.color) + spec) * light.pwr;
    }
    return l_col;
}

// RayTracer::reduce_light as written, rt.rs:956-994
V3 reduce_light_literal(const Scene& sc, const RtParams& rt, const Frame& f, float cx, float cy,
                        const PathRng& rng, Stats* st) {
    st->paths++;
    // RayTracer::iter casts the camera ray ONCE (rt.rs:948); it.clone() copies it, so the
    // throw-away count trace and the collected trace start from the same primary ray.
    Block c = rng.cam();
    const Ray ray0 = rt_iter(cx, cy, f, c.u[0], c.u[1]);
    {   // it.clone().count() == 0  (rt.rs:957): a full throw-away trace with its own bounce randoms
        PathRng rng0 = rng;
        rng0.key ^= 0xA5A5A5A5u;
        Ray ray = ray0;
        PathItem item;
        Stats dummy;
        size_t count = 0;
        while (iter_next(sc, rt, &ray, rng0, &item, &dummy)) count++;
        if (count == 0) { st->hist[0]++; return sc.sky_color; }
    }
    Ray ray = ray0;
    std::vector<PathItem> path;
    PathItem item;
    while (iter_next(sc, rt, &ray, rng, &item, st)) path.push_back(item);
    st->hist[path.size() < 33 ? path.size() : 33]++;
    V3 col = sc.sky_color * sc.sky_pwr;
    for (size_t k = path.size(); k-- > 0;) {
        const PathItem& it = path[k];
        HitView hv{sc, it.hit};
        float emit = hv.emit();
        // the emission draw of this hit: block A slot 3 of the bounce that produced it
        if (bern(rng.a(it.hit.ray.bounce).u[3], emit)) { col = hv.color(); continue; }
        V3 l_col = direct_light(sc, it);
        V3 d_col = 0.5f * col + hadam(hv.color(), col);
        col = (d_col + l_col) * it.hit.ray.pwr;
    }
    return col;
}

// Forward form of the same estimator: L = sum_i T_i (l_col_i pwr_i) + T_end * tail,
// T_{i+1} = T_i (0.5 + color_i) pwr_i, stopping at the first hit whose emission draw passes.
// The emission draws are independent of the path construction (they are made in the fold,
// rt.rs:968), so this has the same law; with shared RNG slots it equals the literal fold of
// the same path up to float association.  A primary miss returns sky.color (rt.rs:958).
V3 reduce_light_forward(const Scene& sc, const RtParams& rt, const Frame& f, float cx, float cy,
                        const PathRng& rng, Stats* st) {
    st->paths++;
    Block c = rng.cam();
    Ray ray = rt_iter(cx, cy, f, c.u[0], c.u[1]);
    V3 L = {0, 0, 0};
    V3 T = {1, 1, 1};
    PathItem item;
    size_t n = 0;
    for (;;) {
        uint32_t b = ray.bounce;
        if (!iter_next(sc, rt, &ray, rng, &item, st)) {
            if (n == 0) { st->hist[0]++; return sc.sky_color; }
            L = L + hadam(T, sc.sky_color * sc.sky_pwr);
            break;
        }
        n++;
        HitView hv{sc, item.hit};
        if (bern(rng.a(b).u[3], hv.emit())) { L = L + hadam(T, hv.color()); break; }
        V3 l_col = direct_light(sc, item);
        L = L + hadam(T, l_col * item.hit.ray.pwr);
        T = hadam(T, (hv.color() + 0.5f) * item.hit.ray.pwr);
    }
    st->hist[n < 33 ? n : 33]++;
    return L;
}

// ---------------------------------------------------------------- film, sampler.rs:80-99
inline uint8_t tonemap1(float v, float gamma, float exp) {
    float g = std::pow(v, gamma);
    float d = (1.0f - exp);
    float t = g * (1.0f + g / (d * d)) / (1.0f + g);
    return f32_as_u8(255.0f * t);
}

// image 0.24 imageops::resize with FilterType::Lanczos3 (third-party, not vendored in the
// reference; restated from the crate's published algorithm: vertical pass into an f32
// intermediate, then horizontal pass, clamp + round at the end).
inline float sinc(float t) {
    float a = t * PI;
    return t == 0.0f ? 1.0f : std::sin(a) / a;
}
inline float lanczos3(float x) { return std::fabs(x) < 3.0f ? sinc(x) * sinc(x / 3.0f) : 0.0f; }

void resample_weights(uint32_t in_n, uint32_t out_n, std::vector<uint32_t>* left_out, std::vector<std::vector<float>>* ws_out) {
    float ratio = (float)in_n / (float)out_n;
    float sratio = ratio < 1.0f ? 1.0f : ratio;
    float src_support = 3.0f * sratio;
    left_out->resize(out_n);
    ws_out->resize(out_n);
    for (uint32_t o = 0; o < out_n; o++) {
        float input = ((float)o + 0.5f) * ratio;
        int64_t left = (int64_t)std::floor(input - src_support);
        if (left < 0) left = 0;
        if (left > (int64_t)in_n - 1) left = (int64_t)in_n - 1;
        int64_t right = (int64_t)std::ceil(input + src_support);
        if (right < left + 1) right = left + 1;
        if (right > (int64_t)in_n) right = (int64_t)in_n;
        input = input - 0.5f;
        std::vector<float>& ws = (*ws_out)[o];
        ws.clear();
        float sum = 0.0f;
        for (int64_t i = left; i < right; i++) {
            float w = lanczos3(((float)i - input) / sratio);
            ws.push_back(w);
            sum += w;
        }
        for (float& w : ws) w /= sum;
        (*left_out)[o] = (uint32_t)left;
    }
}

void resize_lanczos3(const uint8_t* src, uint32_t w, uint32_t h, uint8_t* dst, uint32_t nw, uint32_t nh) {
    if (nw == w && nh == h) { std::memcpy(dst, src, (size_t)w * h * 3); return; }
    std::vector<uint32_t> left;
    std::vector<std::vector<float>> ws;
    // vertical_sample: w x nh f32
    resample_weights(h, nh, &left, &ws);
    std::vector<float> tmp((size_t)w * nh * 3);
    for (uint32_t oy = 0; oy < nh; oy++)
        for (uint32_t x = 0; x < w; x++) {
            float t[3] = {0, 0, 0};
            for (size_t i = 0; i < ws[oy].size(); i++) {
                const uint8_t* p = src + ((size_t)(left[oy] + i) * w + x) * 3;
                for (int c = 0; c < 3; c++) t[c] += (float)p[c] * ws[oy][i];
            }
            for (int c = 0; c < 3; c++) tmp[((size_t)oy * w + x) * 3 + c] = t[c];
        }
    // horizontal_sample: nw x nh u8
    resample_weights(w, nw, &left, &ws);
    for (uint32_t ox = 0; ox < nw; ox++)
        for (uint32_t y = 0; y < nh; y++) {
            float t[3] = {0, 0, 0};
            for (size_t i = 0; i < ws[ox].size(); i++) {
                const float* p = &tmp[((size_t)y * w + left[ox] + i) * 3];
                for (int c = 0; c < 3; c++) t[c] += p[c] * ws[ox][i];
            }
            for (int c = 0; c < 3; c++) {
                float v = t[c] < 0.0f ? 0.0f : (t[c] > 255.0f ? 255.0f : t[c]);
                dst[((size_t)y * nw + ox) * 3 + c] = (uint8_t)std::round(v);
            }
        }
}

}  // namespace

// ================================================================ context + C entry points
struct mrt_cpu_ctx {
    uint32_t workers = 1, n_dim = 64;
    Scene scene;
    bool have_scene = false, have_frame = false;
    Frame frame{};
    RtParams rt{8, 0.15f};
    uint64_t seed = 0x5EED;
    uint32_t rank = 0, world = 1;
    int mode = MRT_CPU_FORWARD;
    uint32_t normal_space = MRT_NORMAL_FORWARD_XF;
    uint32_t nw = 0, nh = 0, passes = 0;
    std::vector<float> colors;  // nw*nh*3, f32 `+=` per pass like Sampler.colors (sampler.rs:60-70)
    Stats stats;
    std::string err;
};

namespace {
thread_local std::string g_create_err;
int fail(mrt_cpu_ctx* c, int code, const std::string& m) { c->err = m; return code; }

void film_dims(const Frame& f, uint32_t* nw, uint32_t* nh) {  // sampler.rs:29-30
    *nw = (uint32_t)f32_as_usize((float)f.res[0] * f.ssaa);
    *nh = (uint32_t)f32_as_usize((float)f.res[1] * f.ssaa);
}
V3 one_path(mrt_cpu_ctx* c, uint32_t x, uint32_t y, uint32_t sample, Stats* st) {
    PathRng rng{y * c->nw + x, sample, fold_seed(c->seed)};
    if (c->mode == MRT_CPU_LITERAL) return reduce_light_literal(c->scene, c->rt, c->frame, (float)x, (float)y, rng, st);
    return reduce_light_forward(c->scene, c->rt, c->frame, (float)x, (float)y, rng, st);
}
}  // namespace

extern "C" {

int mrt_cpu_create(mrt_cpu_ctx** out, uint32_t workers, uint32_t n_dim) {
    if (!out) return MRT_ERR_INVALID;
    auto* c = new mrt_cpu_ctx();
    c->workers = workers ? workers : std::max(1u, std::thread::hardware_concurrency());
    c->n_dim = n_dim ? n_dim : 64;
    *out = c;
    return MRT_OK;
}
void mrt_cpu_destroy(mrt_cpu_ctx* c) { delete c; }
const char* mrt_cpu_last_error(const mrt_cpu_ctx* c) { return c ? c->err.c_str() : g_create_err.c_str(); }

int mrt_cpu_set_scene(mrt_cpu_ctx* c, const mrt_scene* s) {
    if (!c || !s) return MRT_ERR_INVALID;
    Scene sc;
    sc.sky_color = {s->sky_color[0], s->sky_color[1], s->sky_color[2]};
    sc.sky_pwr = s->sky_pwr;
    sc.texels.assign(s->texels, s->texels + 3 * s->n_texels);
    for (uint32_t i = 0; i < s->n_textures; i++) {
        const mrt_texture& t = s->textures[i];
        if (t.has_dat && t.first_texel + (uint64_t)t.w * t.h > s->n_texels) return fail(c, MRT_ERR_INVALID, "texture out of range");
        sc.tex.push_back({t.w, t.h, t.has_dat != 0, nullptr});
    }
    for (uint32_t i = 0; i < s->n_meshes; i++) {
        const mrt_mesh& m = s->meshes[i];
        if ((uint64_t)m.first_tri + m.n_tri > s->n_triangles) return fail(c, MRT_ERR_INVALID, "mesh out of range");
        Mesh mesh;
        for (uint32_t k = 0; k < m.n_tri; k++) {
            const float* p = s->triangles + 9 * (size_t)(m.first_tri + k);
            mesh.tris.push_back({{p[0], p[1], p[2]}, {p[3], p[4], p[5]}, {p[6], p[7], p[8]}});
        }
        V3 aabb;
        if (!mesh_aabb(mesh, &aabb)) return fail(c, MRT_ERR_INVALID, "empty mesh");
        mesh.bvh = bvh_construct(aabb, {0, 0, 0}, mesh.tris, 0, 3);  // parser.rs:815-817
        mesh.has_bvh = true;
        if (!mesh.bvh.has_childs && !mesh.bvh.has_content) return fail(c, MRT_ERR_INVALID, "mesh octree is empty");
        sc.meshes.push_back(std::move(mesh));
    }
    for (uint32_t i = 0; i < s->n_objects; i++) {
        const mrt_object& o = s->objects[i];
        Object ob;
        ob.kind = o.kind;
        ob.r = o.param[0];
        ob.v = {o.param[0], o.param[1], o.param[2]};
        ob.tri = {{o.param[0], o.param[1], o.param[2]}, {o.param[3], o.param[4], o.param[5]}, {o.param[6], o.param[7], o.param[8]}};
        ob.mesh = (int)o.mesh;
        if (o.kind > MRT_MESH) return fail(c, MRT_ERR_INVALID, "unknown object kind");
        if (o.kind == MRT_MESH && o.mesh >= s->n_meshes) return fail(c, MRT_ERR_INVALID, "mesh index out of range");
        const mrt_material& m = o.mat;
        ob.mat = {{m.albedo[0], m.albedo[1], m.albedo[2]}, m.rough, m.metal, m.glass, m.opacity, m.emit,
                  m.tex, m.rmap, m.mmap, m.gmap, m.omap, m.emap};
        for (int id : {m.tex, m.rmap, m.mmap, m.gmap, m.omap, m.emap}) {
            if (id >= (int)s->n_textures) return fail(c, MRT_ERR_INVALID, "texture index out of range");
            if (id >= 0 && (o.kind == MRT_TRIANGLE || o.kind == MRT_MESH))
                return fail(c, MRT_ERR_INVALID, "textured triangle/mesh: to_uv is todo!() in the reference (rt.rs:546,806)");
        }
        if (!(m.emit >= 0.0f && m.emit <= 1.0f)) return fail(c, MRT_ERR_INVALID, "emit outside [0,1]: gen_bool panics (rt.rs:968)");
        if (!(m.opacity >= 0.0f && m.opacity <= 1.0f)) return fail(c, MRT_ERR_INVALID, "opacity outside [0,1]: gen_bool panics (rt.rs:1054)");
        if ((uint64_t)o.first_inst + o.n_inst > s->n_instances) return fail(c, MRT_ERR_INVALID, "instance range");
        for (uint32_t k = 0; k < o.n_inst; k++) {
            const mrt_instance& in = s->instances[o.first_inst + k];
            ob.inst.push_back({{in.pos[0], in.pos[1], in.pos[2]}, {in.dir[0], in.dir[1], in.dir[2], in.dir[3]}});
        }
        sc.objs.push_back(std::move(ob));
    }
    for (uint32_t i = 0; i < s->n_lights; i++) {
        const mrt_light& l = s->lights[i];
        sc.lights.push_back({l.kind, {l.v[0], l.v[1], l.v[2]}, l.pwr, {l.color[0], l.color[1], l.color[2]}});
    }
    sc.normal_space = c->normal_space;
    c->scene = std::move(sc);
    for (uint32_t i = 0; i < s->n_textures; i++)
        c->scene.tex[i].dat = c->scene.texels.data() + 3 * s->textures[i].first_texel;
    c->have_scene = true;
    mrt_cpu_reset(c);
    return MRT_OK;
}

int mrt_cpu_set_frame(mrt_cpu_ctx* c, const mrt_frame* f) {
    if (!c || !f) return MRT_ERR_INVALID;
    Frame fr;
    fr.res[0] = f->res[0]; fr.res[1] = f->res[1]; fr.ssaa = f->ssaa;
    fr.cam = {{f->cam_pos[0], f->cam_pos[1], f->cam_pos[2]}, {f->cam_dir[0], f->cam_dir[1], f->cam_dir[2], f->cam_dir[3]},
              f->fov, f->gamma, f->exp, f->aprt, f->foc};
    uint32_t nw, nh;
    film_dims(fr, &nw, &nh);
    if (nw == 0 || nh == 0) return fail(c, MRT_ERR_INVALID, "empty film");
    c->frame = fr; c->nw = nw; c->nh = nh; c->have_frame = true;
    mrt_cpu_reset(c);
    return MRT_OK;
}
int mrt_cpu_set_rt(mrt_cpu_ctx* c, uint32_t bounce, float loss, uint64_t seed) {
    if (!c) return MRT_ERR_INVALID;
    c->rt = {bounce, loss};
    c->seed = seed;
    return MRT_OK;
}
int mrt_cpu_set_partition(mrt_cpu_ctx* c, uint32_t rank, uint32_t world) {
    if (!c || world == 0 || rank >= world) return MRT_ERR_INVALID;
    c->rank = rank; c->world = world;
    return MRT_OK;
}
int mrt_cpu_set_mode(mrt_cpu_ctx* c, int mode) {
    if (!c || (mode != MRT_CPU_LITERAL && mode != MRT_CPU_FORWARD)) return MRT_ERR_INVALID;
    c->mode = mode;
    return MRT_OK;
}
int mrt_cpu_set_option(mrt_cpu_ctx* c, uint32_t option, uint32_t value) {
    if (!c) return MRT_ERR_INVALID;
    if (option != MRT_OPT_NORMAL_SPACE || value > MRT_NORMAL_OBJECT) return fail(c, MRT_ERR_INVALID, "unknown option or value");
    c->normal_space = value;
    c->scene.normal_space = value;
    return MRT_OK;
}
int mrt_cpu_reset(mrt_cpu_ctx* c) {
    if (!c) return MRT_ERR_INVALID;
    c->passes = 0;
    c->colors.assign((size_t)c->nw * c->nh * 3, 0.0f);
    c->stats = Stats();
    return MRT_OK;
}

// Sampler::execute, sampler.rs:28-78: n_dim x n_dim tile jobs on a pool of `workers`
// threads, one path per supersampled pixel per pass.
int mrt_cpu_execute(mrt_cpu_ctx* c, uint32_t n_passes, double* seconds) {
    if (!c) return MRT_ERR_INVALID;
    if (!c->have_scene || !c->have_frame) return fail(c, MRT_ERR_STATE, "execute before set_scene/set_frame");
    auto t0 = std::chrono::steady_clock::now();
    const uint32_t nw = c->nw, nh = c->nh, nd = c->n_dim;
    const uint32_t g_w = (uint32_t)std::ceil((float)nw / (float)nd);  // sampler.rs:32-33
    const uint32_t g_h = (uint32_t)std::ceil((float)nh / (float)nd);
    for (uint32_t pass = 0; pass < n_passes; pass++) {
        const uint32_t sample = c->rank + (c->passes) * c->world;  // global sample index of this pass
        std::atomic<uint32_t> next{0};
        std::vector<Stats> tstats(c->workers);
        auto job = [&](uint32_t wid) {
            Stats st;
            for (;;) {
                uint32_t tile = next.fetch_add(1);
                if (tile >= nd * nd) break;
                uint32_t g_x = tile / nd, g_y = tile % nd;  // sampler.rs:40-41 loop order
                for (uint32_t x = g_x * g_w; x < (g_x + 1) * g_w && x < nw; x++)
                    for (uint32_t y = g_y * g_h; y < (g_y + 1) * g_h && y < nh; y++) {
                        V3 col = one_path(c, x, y, sample, &st);
                        float* p = &c->colors[((size_t)y * nw + x) * 3];
                        p[0] += col.x; p[1] += col.y; p[2] += col.z;  // sampler.rs:63-67
                    }
            }
            tstats[wid] = st;
        };
        std::vector<std::thread> th;
        for (uint32_t w = 1; w < c->workers; w++) th.emplace_back(job, w);
        job(0);
        for (auto& t : th) t.join();
        for (auto& s : tstats) c->stats.add(s);
        c->passes += 1;  // sampler.rs:76
    }
    if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return MRT_OK;
}

int mrt_cpu_film_size(mrt_cpu_ctx* c, uint32_t* nw, uint32_t* nh, uint32_t* passes) {
    if (!c) return MRT_ERR_INVALID;
    if (nw) *nw = c->nw;
    if (nh) *nh = c->nh;
    if (passes) *passes = c->passes;
    return MRT_OK;
}
int mrt_cpu_accum(mrt_cpu_ctx* c, float* rgb, uint32_t* passes) {
    if (!c || !rgb) return MRT_ERR_INVALID;
    std::memcpy(rgb, c->colors.data(), c->colors.size() * sizeof(float));
    if (passes) *passes = c->passes;
    return MRT_OK;
}
// Sampler::img before the resize, sampler.rs:84-96
int mrt_cpu_img_ss(mrt_cpu_ctx* c, uint8_t* rgb) {
    if (!c || !rgb) return MRT_ERR_INVALID;
    if (c->passes == 0) return fail(c, MRT_ERR_STATE, "img before any pass");
    float inv = 1.0f / (float)c->passes;  // Vec3f / f32 = self * rhs.recip(), lin.rs:296-302
    for (size_t i = 0; i < c->colors.size(); i++)
        rgb[i] = tonemap1(c->colors[i] * inv, c->frame.cam.gamma, c->frame.cam.exp);
    return MRT_OK;
}
int mrt_cpu_img(mrt_cpu_ctx* c, uint8_t* rgb) {
    if (!c || !rgb) return MRT_ERR_INVALID;
    std::vector<uint8_t> ss((size_t)c->nw * c->nh * 3);
    int rc = mrt_cpu_img_ss(c, ss.data());
    if (rc) return rc;
    resize_lanczos3(ss.data(), c->nw, c->nh, rgb, c->frame.res[0], c->frame.res[1]);  // sampler.rs:98
    return MRT_OK;
}
int mrt_cpu_resize_lanczos3(const uint8_t* src, uint32_t w, uint32_t h, uint8_t* dst, uint32_t nw, uint32_t nh) {
    if (!src || !dst || !w || !h || !nw || !nh) return MRT_ERR_INVALID;
    resize_lanczos3(src, w, h, dst, nw, nh);
    return MRT_OK;
}
void mrt_cpu_tonemap(const float* v, float gamma, float exp, uint8_t* out, size_t n) {
    for (size_t i = 0; i < n; i++) out[i] = tonemap1(v[i], gamma, exp);
}
void mrt_cpu_rng_block(uint32_t pixel, uint32_t sample, uint32_t block, uint64_t seed, float out[4]) {
    Block b = rng_block(pixel, sample, block, fold_seed(seed));
    for (int i = 0; i < 4; i++) out[i] = b.u[i];
}

int mrt_cpu_trace_primary(mrt_cpu_ctx* c, mrt_hit* out) {
    if (!c || !out) return MRT_ERR_INVALID;
    if (!c->have_scene || !c->have_frame) return fail(c, MRT_ERR_STATE, "trace before set_scene/set_frame");
    const uint32_t nw = c->nw, nh = c->nh;
    std::atomic<uint32_t> next{0};
    auto job = [&]() {
        for (;;) {
            uint32_t y = next.fetch_add(1);
            if (y >= nh) break;
            for (uint32_t x = 0; x < nw; x++) {
                Ray ray = rt_iter((float)x, (float)y, c->frame, 0.5f, 0.5f);
                mrt_hit h;
                std::memset(&h, 0, sizeof h);
                h.orig[0] = ray.orig.x; h.orig[1] = ray.orig.y; h.orig[2] = ray.orig.z;
                h.dir[0] = ray.dir.x; h.dir[1] = ray.dir.y; h.dir[2] = ray.dir.z;
                Hit h0, h1;
                if (closest_hit(c->scene, ray, &h0, &h1)) {
                    h.t0 = h0.ray.t; h.t1 = h1.ray.t;
                    h.obj = h0.obj; h.inst = h0.inst; h.tri0 = h0.idx; h.tri1 = h1.idx;
                    h.n0[0] = h0.norm.x; h.n0[1] = h0.norm.y; h.n0[2] = h0.norm.z;
                    h.n1[0] = h1.norm.x; h.n1[1] = h1.norm.y; h.n1[2] = h1.norm.z;
                    const Object& o = c->scene.objs[h0.obj];
                    if (o.kind <= MRT_BOX) {
                        V2 uv = obj_uv(o, o.inst[h0.inst], point(h0.ray));
                        h.uv[0] = uv.x; h.uv[1] = uv.y;
                    }
                } else {
                    h.t0 = -1.0f; h.t1 = -1.0f; h.obj = h.inst = h.tri0 = h.tri1 = -1;
                }
                out[(size_t)y * nw + x] = h;
            }
        }
    };
    std::vector<std::thread> th;
    for (uint32_t w = 1; w < c->workers; w++) th.emplace_back(job);
    job();
    for (auto& t : th) t.join();
    return MRT_OK;
}

int mrt_cpu_path(mrt_cpu_ctx* c, uint32_t x, uint32_t y, uint32_t sample, float rgb[3]) {
    if (!c || !rgb) return MRT_ERR_INVALID;
    if (!c->have_scene || !c->have_frame) return fail(c, MRT_ERR_STATE, "path before set_scene/set_frame");
    Stats st;
    V3 col = one_path(c, x, y, sample, &st);
    rgb[0] = col.x; rgb[1] = col.y; rgb[2] = col.z;
    return MRT_OK;
}

int mrt_cpu_get_stats(mrt_cpu_ctx* c, mrt_cpu_stats* out) {
    if (!c || !out) return MRT_ERR_INVALID;
    out->paths = c->stats.paths; out->segments = c->stats.segments; out->hits = c->stats.hits;
    out->shadow_rays = c->stats.shadow_rays; out->nan_normals = c->stats.nan_normals;
    for (int i = 0; i < 34; i++) out->hit_hist[i] = c->stats.hist[i];
    return MRT_OK;
}

}  // extern "C"
