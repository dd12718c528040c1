// mrt_api.cu — the C ABI of include/mrt.h: context, scene packing into the device layout
// (SlimInst/FatInst/Xf/lights/textures/flattened mesh octree), launch scheduling, film
// read-out.  Host code only; every pixel is computed by the kernels in mrt_kernels.cu.
// There is no CPU fallback: without a usable CUDA device every compute entry point fails.
#include "mrt_device.cuh"
#include "mrt_kernels.h"
#include "mrt_jit.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace {

thread_local std::string g_create_err;

// ---------------------------------------------------------------- host f32 math (lin.rs order)
struct H3 { float x, y, z; };
inline H3 hsub(H3 a, H3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline H3 hcross(H3 a, H3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline float hdot(H3 a, H3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline H3 hnorm(H3 a) { float r = 1.0f / std::sqrt(a.x * a.x + a.y * a.y + a.z * a.z); return {a.x * r, a.y * r, a.z * r}; }
struct HM { float m[9]; };
inline H3 hmul(const HM& m, H3 v) {
    return {m.m[0] * v.x + m.m[1] * v.y + m.m[2] * v.z, m.m[3] * v.x + m.m[4] * v.y + m.m[5] * v.z,
            m.m[6] * v.x + m.m[7] * v.y + m.m[8] * v.z};
}
// M = rotate_y(dir) * lookat(dir, up): lin.rs:175-183, 197-209; applied as rot_y * (look * v)
HM transform_of(const float dir[4]) {
    const float w = dir[0];
    const float cw = std::sqrt(1.0f - w * w);
    const HM ry = {{cw, 0.0f, w, 0.0f, 1.0f, 0.0f, -w, 0.0f, cw}};
    const H3 fwd = hnorm({dir[1], dir[2], dir[3]});
    const H3 right = hnorm(hcross(fwd, {0.0f, 0.0f, 1.0f}));
    const H3 up = hcross(right, fwd);
    const HM lk = {{right.x, -right.y, right.z, -fwd.x, fwd.y, -fwd.z, up.x, -up.y, up.z}};
    HM out;
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++)
            out.m[3 * r + c] = ry.m[3 * r] * lk.m[c] + ry.m[3 * r + 1] * lk.m[3 + c] + ry.m[3 * r + 2] * lk.m[6 + c];
    return out;
}
bool is_identity(const HM& m) {
    const float id[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    for (int i = 0; i < 9; i++)
        if (!(m.m[i] == id[i])) return false;
    return true;
}
bool finite_m(const HM& m) {
    for (float v : m.m) if (!std::isfinite(v)) return false;
    return true;
}
inline float u2f(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    cudaError_t upload(const std::vector<T>& h) {
        release();
        n = h.size();
        const size_t bytes = std::max<size_t>(1, n) * sizeof(T);
        cudaError_t e = cudaMalloc((void**)&p, bytes);
        if (e != cudaSuccess) { p = nullptr; return e; }
        if (n) e = cudaMemcpy(p, h.data(), n * sizeof(T), cudaMemcpyHostToDevice);
        return e;
    }
    cudaError_t alloc(size_t count) {
        if (p && n == count) return cudaSuccess;
        release();
        n = count;
        cudaError_t e = cudaMalloc((void**)&p, std::max<size_t>(1, n) * sizeof(T));
        if (e != cudaSuccess) p = nullptr;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
};

}  // namespace

struct mrt_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;      // the stream work is queued on
    cudaStream_t own_stream = nullptr;  // created by mrt_create
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::string err;
    uint64_t launches = 0;

    // scene
    bool have_scene = false, in_param = false;
    uint32_t features = 0;
    ParamScene* pscene = nullptr;  // host staging copies of the kernel-parameter structs
    GlobalScene gscene{};
    DevBuf<SlimInst> d_slim[K_NKIND];
    DevBuf<Xf> d_mesh_m;
    DevBuf<BoxPair> d_boxp;
    DevBuf<BvhNode> d_bvh;
    DevBuf<BxfInst> d_bxf;
    DevBuf<FatInst> d_fat;
    DevBuf<DTex> d_tex;
    DevBuf<float4> d_texels;
    DevBuf<DMesh> d_mesh;
    DevBuf<DMeshLeaf> d_leaf;
    DevBuf<uint32_t> d_leaf_idx;
    DevBuf<DTri> d_tri;
    DevBuf<BvhNode> d_tbvh;          // triangle BVHs of the meshes
    DevBuf<DTriLeaf> d_tri_leaf;     // per triangle: the octree leaves that list it
    DevBuf<uint32_t> d_obj_inst;

    // frame / rt
    bool have_frame = false;
    mrt_frame frame{};
    uint32_t nw = 0, nh = 0;
    uint32_t bounce = 8;
    float loss = 0.15f;
    uint64_t seed = 0x5EED;
    uint32_t rank = 0, world = 1;
    uint32_t passes = 0;        // passes this context rendered (local)
    uint32_t passes_total = 0;  // passes the accumulator holds (after an external reduce)
    bool tiled = true;  // warp = 8x4 pixel tile (FilmParams::tiles_x); MRT_TILE=0: 32 pixels of a row (A/B knob)
    uint32_t spp_per_launch = 1024;  // measured: 128 -> 8917, 256 -> 9058, 1024 -> 9234 Mpaths/s (intra-warp tail)
    uint32_t normal_space = MRT_NORMAL_FORWARD_XF;  // MRT_OPT_NORMAL_SPACE

    // run-time scene specialisation (mrt_jit.cu)
    uint32_t jit_mode = MRT_JIT_AUTO;   // MRT_OPT_JIT
    std::string jit_header;             // "" = scene not eligible
    cudaKernel_t jit_kernel = nullptr;  // compiled for jit_header
    bool jit_requested = false, jit_failed = false, jit_from_disk = false;
    double jit_seconds = 0.0;
    std::string jit_err;
    uint64_t jit_launches = 0;

    // film
    DevBuf<float4> d_accum;
    DevBuf<uint8_t> d_ss, d_out;
    DevBuf<float> d_tmp, d_rgb, d_wv, d_wh;
    DevBuf<int32_t> d_lv, d_cv, d_lh, d_ch;
    DevBuf<mrt_hit> d_hits;
    uint32_t taps_v = 0, taps_h = 0;
    bool weights_ready = false;

    ~mrt_ctx() { delete pscene; }
};

namespace {

int fail(mrt_ctx* c, int code, const std::string& m) { c->err = m; return code; }
int cuda_fail(mrt_ctx* c, cudaError_t e, const char* what) {
    c->err = std::string(what) + ": " + cudaGetErrorString(e);
    return MRT_ERR_CUDA;
}
#define CK(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return cuda_fail(c, e__, #call); } while (0)

uint32_t fold_seed(uint64_t seed) { return (uint32_t)seed ^ ((uint32_t)(seed >> 32) * 0x9E3779B9u); }

void film_dims(const mrt_frame& f, uint32_t* nw, uint32_t* nh) {  // sampler.rs:29-30
    auto cast = [](float v) -> uint32_t { return v > 0.0f ? (v >= 4294967040.0f ? 0xffffffffu : (uint32_t)v) : 0u; };
    *nw = cast((float)f.res[0] * f.ssaa);
    *nh = cast((float)f.res[1] * f.ssaa);
}

FilmParams make_film_params(const mrt_ctx* c) {
    FilmParams fp{};
    const mrt_frame& f = c->frame;
    fp.accum = c->d_accum.p;
    fp.nw = c->nw; fp.nh = c->nh;
    fp.key = fold_seed(c->seed);
    fp.max_bounce = c->bounce;
    fp.keep = 1.0f - std::fmin(c->loss, 1.0f);
    fp.cam_pos[0] = f.cam_pos[0]; fp.cam_pos[1] = f.cam_pos[1]; fp.cam_pos[2] = f.cam_pos[2];
    fp.aprt = f.aprt; fp.foc = f.foc;
    const HM m = transform_of(f.cam_dir);
    for (int i = 0; i < 9; i++) fp.cam_M[i] = m.m[i];
    fp.cam_identity = is_identity(m) ? 1u : 0u;
    fp.fw = (float)f.res[0] * f.ssaa;
    fp.fh = (float)f.res[1] * f.ssaa;
    const float tan_fov = std::tan((0.5f * f.fov) * (3.14159265358979323846f / 180.0f));  // rt.rs:902
    fp.fy = 1.0f / (2.0f * tan_fov);
    fp.tiles_x = c->tiled ? (c->nw + 15u) / 16u : 0u;
    return fp;
}

// Flattened depth-3 octree of one mesh: the non-empty leaves in the reference's depth-first
// child order (rt.rs:631-689, parser.rs:805-824), each with the triangles that have a vertex
// inside it (rt.rs:227-248).
struct LeafBuild { H3 center, size; std::vector<uint32_t> idx; };
void build_leaves(const float* tris, uint32_t n_tri, std::vector<LeafBuild>* out, float root_half[3]) {
    static const float G[8][3] = {{1, 1, 1}, {-1, 1, 1}, {-1, -1, 1}, {1, -1, 1}, {1, 1, -1}, {-1, 1, -1}, {-1, -1, -1}, {1, -1, -1}};
    float mx = 0, my = 0, mz = 0;  // Mesh::gen_aabb, rt.rs:261-270
    for (uint32_t t = 0; t < n_tri; t++)
        for (int v = 0; v < 3; v++) {
            const float* p = tris + 9 * (size_t)t + 3 * v;
            mx = std::fmax(mx, std::fabs(p[0])); my = std::fmax(my, std::fabs(p[1])); mz = std::fmax(mz, std::fabs(p[2]));
        }
    const H3 A0 = {2.0f * mx, 2.0f * my, 2.0f * mz};
    root_half[0] = 0.5f * A0.x; root_half[1] = 0.5f * A0.y; root_half[2] = 0.5f * A0.z;  // Box::intersect halves the size, rt.rs:318
    const H3 A1 = {0.5f * A0.x, 0.5f * A0.y, 0.5f * A0.z};
    const H3 A2 = {0.5f * A1.x, 0.5f * A1.y, 0.5f * A1.z};
    const H3 A3 = {0.5f * A2.x, 0.5f * A2.y, 0.5f * A2.z};
    for (int i0 = 0; i0 < 8; i0++) {
        const H3 r1 = {0.0f + A0.x * (G[i0][0] * 0.25f), 0.0f + A0.y * (G[i0][1] * 0.25f), 0.0f + A0.z * (G[i0][2] * 0.25f)};
        for (int i1 = 0; i1 < 8; i1++) {
            const H3 r2 = {r1.x + A1.x * (G[i1][0] * 0.25f), r1.y + A1.y * (G[i1][1] * 0.25f), r1.z + A1.z * (G[i1][2] * 0.25f)};
            for (int i2 = 0; i2 < 8; i2++) {
                const H3 r3 = {r2.x + A2.x * (G[i2][0] * 0.25f), r2.y + A2.y * (G[i2][1] * 0.25f), r2.z + A2.z * (G[i2][2] * 0.25f)};
                const H3 hi = {r3.x + 0.5f * A3.x, r3.y + 0.5f * A3.y, r3.z + 0.5f * A3.z};
                const H3 lo = {r3.x - 0.5f * A3.x, r3.y - 0.5f * A3.y, r3.z - 0.5f * A3.z};
                LeafBuild lb{r3, A3, {}};
                for (uint32_t t = 0; t < n_tri; t++) {
                    bool in = false;
                    for (int v = 0; v < 3 && !in; v++) {
                        const float* p = tris + 9 * (size_t)t + 3 * v;
                        in = !(p[0] > hi.x || p[1] > hi.y || p[2] > hi.z) && !(p[0] < lo.x || p[1] < lo.y || p[2] < lo.z);
                    }
                    if (in) lb.idx.push_back(t);
                }
                if (!lb.idx.empty()) out->push_back(std::move(lb));
            }
        }
    }
}

// float literal that round-trips exactly (C++17 hex float)
void lit(std::string* o, float v) {
    char b[48];
    std::snprintf(b, sizeof b, "%af", (double)v);
    *o += b;
}
void lits(std::string* o, const float* v, int n) {
    for (int i = 0; i < n; i++) { *o += ", "; lit(o, v[i]); }
}
bool all_finite(const float* v, int n) {
    for (int i = 0; i < n; i++) if (!std::isfinite(v[i])) return false;
    return true;
}

// ---- BVH builder (scene-level BVH over the finite instances, triangle BVHs of the meshes; mrt_device.cuh:
// BvhNode): median split of the centroids along the widest axis, one primitive per leaf (measured best).
struct PrimBox { float lo[3], hi[3]; uint32_t ref; };
bool g_bvh_sah = true;  // MRT_BVH_SAH=0: median splits only (A/B knob, read in mrt_create)
// Returns the reference of the subtree over prims[begin, end): a leaf (MRT_BVH_LEAF | prims[begin].ref) for a
// single primitive, else the index of a node that holds the boxes and references of its two halves.
uint32_t bvh_build(std::vector<PrimBox>& prims, size_t begin, size_t end, std::vector<BvhNode>* nodes, int depth = 0, int* max_depth = nullptr) {
    if (max_depth) *max_depth = std::max(*max_depth, depth);
    if (end - begin == 1) return MRT_BVH_LEAF | prims[begin].ref;
    float clo[3] = {INFINITY, INFINITY, INFINITY}, chi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (size_t i = begin; i < end; i++)
        for (int a = 0; a < 3; a++) {
            const float cc = 0.5f * (prims[i].lo[a] + prims[i].hi[a]);
            clo[a] = std::fmin(clo[a], cc); chi[a] = std::fmax(chi[a], cc);
        }
    int ax = 0;
    if (chi[1] - clo[1] > chi[ax] - clo[ax]) ax = 1;
    if (chi[2] - clo[2] > chi[ax] - clo[ax]) ax = 2;
    size_t mid = begin + (end - begin) / 2;
    bool split_done = false;
    if (g_bvh_sah && end - begin > 4) {
        // binned surface-area heuristic over the three axes (16 bins of the centroid range); falls back to the
        // median when every centroid lands in one bin or the best split is lopsided beyond the stack's depth budget
        constexpr int NB = 16;
        float best_cost = INFINITY; int best_ax = -1, best_bin = -1;
        for (int a = 0; a < 3; a++) {
            const float ext = chi[a] - clo[a];
            if (!(ext > 0.0f)) continue;
            struct Bin { float lo[3], hi[3]; size_t n; } bins[NB];
            for (auto& b : bins) { for (int k = 0; k < 3; k++) { b.lo[k] = INFINITY; b.hi[k] = -INFINITY; } b.n = 0; }
            const float scale = (float)NB / ext;
            for (size_t i = begin; i < end; i++) {
                const float cc = 0.5f * (prims[i].lo[a] + prims[i].hi[a]);
                const int bi = std::min(NB - 1, std::max(0, (int)((cc - clo[a]) * scale)));
                Bin& b = bins[bi];
                for (int k = 0; k < 3; k++) { b.lo[k] = std::fmin(b.lo[k], prims[i].lo[k]); b.hi[k] = std::fmax(b.hi[k], prims[i].hi[k]); }
                b.n++;
            }
            auto area = [](const float* lo, const float* hi) {
                const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
                return dx * dy + dy * dz + dz * dx;
            };
            float la[NB], ra[NB]; size_t ln[NB], rn[NB];
            float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY}; size_t n = 0;
            for (int b = 0; b < NB; b++) {
                if (bins[b].n) for (int k = 0; k < 3; k++) { lo[k] = std::fmin(lo[k], bins[b].lo[k]); hi[k] = std::fmax(hi[k], bins[b].hi[k]); }
                n += bins[b].n; ln[b] = n; la[b] = n ? area(lo, hi) : 0.0f;
            }
            for (int k = 0; k < 3; k++) { lo[k] = INFINITY; hi[k] = -INFINITY; } n = 0;
            for (int b = NB - 1; b >= 0; b--) {
                if (bins[b].n) for (int k = 0; k < 3; k++) { lo[k] = std::fmin(lo[k], bins[b].lo[k]); hi[k] = std::fmax(hi[k], bins[b].hi[k]); }
                n += bins[b].n; rn[b] = n; ra[b] = n ? area(lo, hi) : 0.0f;
            }
            for (int b = 0; b + 1 < NB; b++) {  // split after bin b
                if (ln[b] == 0 || rn[b + 1] == 0) continue;
                const float cost = la[b] * (float)ln[b] + ra[b + 1] * (float)rn[b + 1];
                if (cost < best_cost) { best_cost = cost; best_ax = a; best_bin = b; }
            }
        }
        if (best_ax >= 0) {
            const int a = best_ax;
            const float scale = 16.0f / (chi[a] - clo[a]), c0 = clo[a];
            auto it = std::partition(prims.begin() + begin, prims.begin() + end, [&](const PrimBox& p) {
                const float cc = 0.5f * (p.lo[a] + p.hi[a]);
                return std::min(15, std::max(0, (int)((cc - c0) * scale))) <= best_bin;
            });
            const size_t m = (size_t)(it - prims.begin());
            const size_t small = std::min(m - begin, end - m);
            if (m > begin && m < end && small * 16 >= (end - begin) / 4 + 1) { mid = m; split_done = true; }  // keep the depth bounded
        }
    }
    if (!split_done) {
        mid = begin + (end - begin) / 2;
        std::nth_element(prims.begin() + begin, prims.begin() + mid, prims.begin() + end, [ax](const PrimBox& a, const PrimBox& b) {
            return a.lo[ax] + a.hi[ax] < b.lo[ax] + b.hi[ax];
        });
    }
    auto bounds = [&](size_t b0, size_t e0, float* lo, float* hi) {
        for (int a = 0; a < 3; a++) { lo[a] = INFINITY; hi[a] = -INFINITY; }
        for (size_t i = b0; i < e0; i++)
            for (int a = 0; a < 3; a++) { lo[a] = std::fmin(lo[a], prims[i].lo[a]); hi[a] = std::fmax(hi[a], prims[i].hi[a]); }
    };
    const size_t node = nodes->size();
    nodes->emplace_back();
    float llo[3], lhi[3], rlo[3], rhi[3];
    bounds(begin, mid, llo, lhi);
    bounds(mid, end, rlo, rhi);
    const uint32_t l = bvh_build(prims, begin, mid, nodes, depth + 1, max_depth);
    const uint32_t r = bvh_build(prims, mid, end, nodes, depth + 1, max_depth);
    // centre / half-extent form, left child in the low lane; the half extent is taken from the centre AS ROUNDED
    // and padded, so the stored box still covers [lo, hi]
    float cl[3], hl[3], cr[3], hr[3];
    auto centre_half = [](const float* lo, const float* hi, float* c, float* h) {
        for (int a = 0; a < 3; a++) {
            c[a] = 0.5f * (lo[a] + hi[a]);
            const float e = std::fmax(hi[a] - c[a], c[a] - lo[a]);
            h[a] = e * (1.0f + 4e-7f) + 1e-30f;
        }
    };
    centre_half(llo, lhi, cl, hl);
    centre_half(rlo, rhi, cr, hr);
    BvhNode& n = (*nodes)[node];
    n.q0 = make_float4(cl[0], cr[0], cl[1], cr[1]);
    n.q1 = make_float4(cl[2], cr[2], hl[0], hr[0]);
    n.q2 = make_float4(hl[1], hr[1], hl[2], hr[2]);
    n.ref = make_uint4(l, r, 0u, 0u);
    return (uint32_t)node;
}
// world-space AABB of an object-space box of half extents h centred on pos, under world->object matrix M
// (object->world is M^T), padded so that rounding in the primitive tests cannot leave the node
void world_box(const HM& M, H3 pos, H3 h, PrimBox* b) {
    const float hw[3] = {std::fabs(M.m[0]) * h.x + std::fabs(M.m[3]) * h.y + std::fabs(M.m[6]) * h.z,
                         std::fabs(M.m[1]) * h.x + std::fabs(M.m[4]) * h.y + std::fabs(M.m[7]) * h.z,
                         std::fabs(M.m[2]) * h.x + std::fabs(M.m[5]) * h.y + std::fabs(M.m[8]) * h.z};
    const float p[3] = {pos.x, pos.y, pos.z};
    for (int a = 0; a < 3; a++) {
        const float pad = 1e-4f * (std::fabs(hw[a]) + std::fabs(p[a])) + 1e-5f;
        b->lo[a] = p[a] - std::fabs(hw[a]) - pad;
        b->hi[a] = p[a] + std::fabs(hw[a]) + pad;
    }
}

uint32_t pack_ids(int32_t lo, int32_t hi) { return ((uint32_t)(lo < 0 ? 0xffff : lo) & 0xffffu) | (((uint32_t)(hi < 0 ? 0xffff : hi) & 0xffffu) << 16); }

}  // namespace

extern "C" {

int mrt_abi_version(void) { return MRT_ABI_VERSION; }

const char* mrt_last_error(const mrt_ctx* c) { return c ? c->err.c_str() : g_create_err.c_str(); }

int mrt_device_count(int* n) {
    if (!n) return MRT_ERR_INVALID;
    *n = 0;
    int k = 0;
    if (cudaGetDeviceCount(&k) != cudaSuccess) { cudaGetLastError(); return MRT_ERR_CUDA; }
    *n = k;
    return MRT_OK;
}

int mrt_create(mrt_ctx** out, int device, uint32_t workers, uint32_t n_dim) {
    (void)workers; (void)n_dim;  // --worker / --dim: the CUDA grid replaces the tile pool
    if (!out) { g_create_err = "mrt_create: null out"; return MRT_ERR_INVALID; }
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        g_create_err = std::string("mrt_create: no CUDA device (") + (e != cudaSuccess ? cudaGetErrorString(e) : "count 0") + "); there is no CPU fallback";
        return MRT_ERR_CUDA;
    }
    if (device < 0 || device >= n) { g_create_err = "mrt_create: device index out of range"; return MRT_ERR_INVALID; }
    mrt_ctx* c = new mrt_ctx();
    c->device = device;
    e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking);
    c->stream = c->own_stream;
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev1);
    if (e != cudaSuccess) {
        g_create_err = std::string("mrt_create: ") + cudaGetErrorString(e);
        delete c;
        return MRT_ERR_CUDA;
    }
    if (const char* s = std::getenv("MRT_TILE")) c->tiled = std::atoi(s) != 0;
    { const char* s = std::getenv("MRT_BVH_SAH"); g_bvh_sah = s ? std::atoi(s) != 0 : true; }
    if (const char* s = std::getenv("MRT_SPP_PER_LAUNCH")) {
        const int v = std::atoi(s);
        if (v > 0) c->spp_per_launch = (uint32_t)v;
    }
    if (const char* s = std::getenv("MRT_JIT")) {  // default MRT_OPT_JIT of new contexts (experiments, CI)
        const int v = std::atoi(s);
        if (v >= 0 && v <= (int)MRT_JIT_FORCE) c->jit_mode = (uint32_t)v;
    }
    c->pscene = new ParamScene();
    *out = c;
    return MRT_OK;
}

void mrt_destroy(mrt_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->jit_requested && !c->jit_header.empty()) mrt_jit_wait(c->jit_header);
    for (auto& b : c->d_slim) b.release();
    c->d_boxp.release(); c->d_bvh.release(); c->d_bxf.release(); c->d_mesh_m.release(); c->d_fat.release(); c->d_tex.release(); c->d_texels.release();
    c->d_mesh.release(); c->d_leaf.release(); c->d_leaf_idx.release(); c->d_tri.release(); c->d_tbvh.release(); c->d_tri_leaf.release(); c->d_obj_inst.release();
    c->d_accum.release(); c->d_ss.release(); c->d_out.release(); c->d_tmp.release(); c->d_rgb.release();
    c->d_wv.release(); c->d_wh.release(); c->d_lv.release(); c->d_cv.release(); c->d_lh.release(); c->d_ch.release();
    c->d_hits.release();
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
}

int mrt_reset(mrt_ctx* c) {
    if (!c) return MRT_ERR_INVALID;
    CK(cudaSetDevice(c->device));
    c->passes = 0;
    c->passes_total = 0;
    if (c->d_accum.p) CK(cudaMemsetAsync(c->d_accum.p, 0, c->d_accum.n * sizeof(float4), c->stream));
    return MRT_OK;
}

int mrt_set_scene(mrt_ctx* c, const mrt_scene* s) {
    if (!c || !s) return MRT_ERR_INVALID;
    CK(cudaSetDevice(c->device));
    if (s->n_lights > MRT_MAX_LIGHTS) return fail(c, MRT_ERR_INVALID, "more than 16 lights are not supported");
    if (s->n_objects > 0xffffu) return fail(c, MRT_ERR_INVALID, "too many objects");
    if (s->n_textures >= 0xffffu) return fail(c, MRT_ERR_INVALID, "too many textures");

    uint32_t feat = s->n_lights ? F_LIGHTS : 0u;
    // textures -> float4 texels
    std::vector<DTex> tex(s->n_textures);
    std::vector<float4> texels;
    for (uint32_t i = 0; i < s->n_textures; i++) {
        const mrt_texture& t = s->textures[i];
        const uint64_t n = (uint64_t)t.w * t.h;
        tex[i] = {t.w, t.h, (uint32_t)texels.size(), (t.has_dat && n > 0) ? 1u : 0u};
        if (tex[i].has_dat) {
            if (t.first_texel + n > s->n_texels) return fail(c, MRT_ERR_INVALID, "texture texel range out of bounds");
            if (texels.size() + n > 0x7fffffffull) return fail(c, MRT_ERR_INVALID, "textures too large");
            for (uint64_t k = 0; k < n; k++) {
                const float* p = s->texels + 3 * (t.first_texel + k);
                texels.push_back(make_float4(p[0], p[1], p[2], 0.0f));
            }
        }
    }
    // meshes -> leaves + triangles
    std::vector<DMesh> meshes(s->n_meshes);
    std::vector<DMeshLeaf> leaves;
    std::vector<uint32_t> leaf_idx;
    std::vector<DTri> tris;
    std::vector<BvhNode> tbvh;
    std::vector<DTriLeaf> tri_leaf;
    const bool mesh_bvh = !std::getenv("MRT_NO_MESH_BVH");  // test knob: the sequential leaf walk instead
    for (uint32_t i = 0; i < s->n_meshes; i++) {
        const mrt_mesh& m = s->meshes[i];
        if ((uint64_t)m.first_tri + m.n_tri > s->n_triangles) return fail(c, MRT_ERR_INVALID, "mesh triangle range out of bounds");
        if (m.n_tri == 0) return fail(c, MRT_ERR_INVALID, "empty mesh");
        const float* tp = s->triangles + 9 * (size_t)m.first_tri;
        std::vector<LeafBuild> lb;
        float root_half[3];
        build_leaves(tp, m.n_tri, &lb, root_half);
        if (lb.empty()) return fail(c, MRT_ERR_INVALID, "mesh octree is empty (the reference would panic, rt.rs:717)");
        meshes[i] = {(uint32_t)leaves.size(), (uint32_t)lb.size(), (uint32_t)tris.size(), m.n_tri, {root_half[0], root_half[1], root_half[2]}, 0xffffffffu};
        for (const LeafBuild& l : lb) {
            DMeshLeaf dl;
            dl.lo = make_float4(l.center.x - 0.5f * l.size.x, l.center.y - 0.5f * l.size.y, l.center.z - 0.5f * l.size.z, u2f((uint32_t)leaf_idx.size()));
            dl.hi = make_float4(l.center.x + 0.5f * l.size.x, l.center.y + 0.5f * l.size.y, l.center.z + 0.5f * l.size.z, u2f((uint32_t)l.idx.size()));
            leaves.push_back(dl);
            leaf_idx.insert(leaf_idx.end(), l.idx.begin(), l.idx.end());
        }
        // per triangle: its occurrences in the candidate sequence (leaf order, then list order), ascending
        std::vector<std::vector<DTriLeaf>> occ(m.n_tri);
        {
            size_t listed = 0;
            for (const LeafBuild& l : lb) listed += l.idx.size();
            uint32_t rank = (uint32_t)(leaf_idx.size() - listed);  // = this mesh's first position in leaf_idx
            for (size_t l = 0; l < lb.size(); l++)
                for (uint32_t ti : lb[l].idx) occ[ti].push_back(DTriLeaf{(uint32_t)(meshes[i].first_leaf + l), rank++});
        }
        for (uint32_t t = 0; t < m.n_tri; t++) {
            const float* p = tp + 9 * (size_t)t;
            DTri d;
            d.v0 = make_float4(p[0], p[1], p[2], u2f((uint32_t)tri_leaf.size()));
            d.e0 = make_float4(p[3] - p[0], p[4] - p[1], p[5] - p[2], u2f((uint32_t)occ[t].size()));
            d.e1 = make_float4(p[6] - p[0], p[7] - p[1], p[8] - p[2], 0.0f);
            tris.push_back(d);
            tri_leaf.insert(tri_leaf.end(), occ[t].begin(), occ[t].end());
        }
        // triangle BVH (median split, padded boxes: rounding in tri_test must not be able to leave a node).
        // A triangle no leaf lists can never be a candidate and is left out.
        meshes[i].bvh_root = 0xffffffffu;
        if (mesh_bvh && m.n_tri < (1u << 26)) {
            std::vector<PrimBox> pb;
            pb.reserve(m.n_tri);
            bool finite = true;
            for (uint32_t t = 0; t < m.n_tri; t++) {
                if (occ[t].empty()) continue;
                const float* p = tp + 9 * (size_t)t;
                PrimBox b;
                for (int a = 0; a < 3; a++) {
                    const float lo = std::fmin(p[a], std::fmin(p[3 + a], p[6 + a])), hi = std::fmax(p[a], std::fmax(p[3 + a], p[6 + a]));
                    const float pad = 1e-4f * (std::fabs(lo) + std::fabs(hi)) + 1e-5f;
                    b.lo[a] = lo - pad; b.hi[a] = hi + pad;
                    finite &= std::isfinite(b.lo[a]) && std::isfinite(b.hi[a]);
                }
                b.ref = t;
                pb.push_back(b);
            }
            if (finite && !pb.empty()) {
                const size_t mark = tbvh.size();
                int depth = 0;
                const uint32_t root = bvh_build(pb, 0, pb.size(), &tbvh, 0, &depth);
                if (depth <= 30) meshes[i].bvh_root = root;  // the traversal stack holds 32 entries
                else tbvh.resize(mark);
            }
        }
    }
    // instances, grouped by kind (declaration order inside a kind)
    std::vector<SlimInst> by_kind[K_NKIND];
    std::vector<FatInst> fat_k[K_NKIND];
    std::vector<uint32_t> oi_k[K_NKIND];
    std::vector<Xf> mesh_m;
    std::vector<BxfInst> bxf;
    int rot_class = 0;  // MRT_JIT_ROT: 0 no rotated instance, 1 yaw-only, 2 general
    std::vector<PrimBox> prim_boxes;  // finite instances, for the scene-level BVH
    bool prim_boxes_ok = true;
    for (uint32_t oi = 0; oi < s->n_objects; oi++) {
        const mrt_object& o = s->objects[oi];
        const mrt_material& mt = o.mat;
        if (o.kind > MRT_MESH) return fail(c, MRT_ERR_INVALID, "unknown object kind");
        if (o.kind == MRT_TRIANGLE)
            return fail(c, MRT_ERR_INVALID, "top-level triangle objects panic in the reference (Triangle::gen_aabb is todo!(), rt.rs:224); use a mesh");
        if (o.kind == MRT_MESH && o.mesh >= s->n_meshes) return fail(c, MRT_ERR_INVALID, "mesh index out of range");
        if ((uint64_t)o.first_inst + o.n_inst > s->n_instances) return fail(c, MRT_ERR_INVALID, "instance range out of bounds");
        if (o.n_inst > 0xffffu) return fail(c, MRT_ERR_INVALID, "too many instances in one object");
        if (!(mt.emit >= 0.0f && mt.emit <= 1.0f)) return fail(c, MRT_ERR_INVALID, "material emit outside [0,1] (gen_bool panics, rt.rs:968)");
        if (!(mt.opacity >= 0.0f && mt.opacity <= 1.0f)) return fail(c, MRT_ERR_INVALID, "material opacity outside [0,1] (gen_bool panics, rt.rs:1054)");
        const int32_t ids[6] = {mt.tex, mt.rmap, mt.mmap, mt.gmap, mt.omap, mt.emap};
        bool textured = false;
        for (int32_t id : ids) {
            if (id >= (int32_t)s->n_textures) return fail(c, MRT_ERR_INVALID, "texture index out of range");
            textured |= id >= 0;
        }
        if (textured && o.kind == MRT_MESH) return fail(c, MRT_ERR_INVALID, "textured mesh: to_uv is todo!() in the reference (rt.rs:806)");
        if (textured) feat |= F_TEX;
        if (mt.opacity < 1.0f || mt.omap >= 0) feat |= F_TRANSMIT;
        if (o.kind == MRT_MESH) feat |= F_MESH;
        for (uint32_t k = 0; k < o.n_inst; k++) {
            const mrt_instance& in = s->instances[o.first_inst + k];
            const float nd[4] = {-in.dir[0], -in.dir[1], -in.dir[2], -in.dir[3]};  // rt.rs:726: -inst.dir
            const HM M = transform_of(nd);
            if (!finite_m(M)) return fail(c, MRT_ERR_INVALID, "instance dir gives a non-finite transform (|w| > 1, zero or vertical facing vector)");
            const bool ident = is_identity(M);
            if (!ident) {
                const bool yaw = M.m[2] == 0.0f && M.m[5] == 0.0f && M.m[6] == 0.0f && M.m[7] == 0.0f && M.m[8] == 1.0f;
                rot_class = std::max(rot_class, yaw ? 1 : 2);
            }
            const H3 pos = {in.pos[0], in.pos[1], in.pos[2]};
            SlimInst si{};
            FatInst fi{};
            uint32_t kind;
            Xf x{};
            for (int r = 0; r < 3; r++) for (int cc = 0; cc < 3; cc++) x.m[4 * r + cc] = M.m[3 * r + cc];
            PrimBox pb{};
            bool finite_prim = true;
            if (o.kind == MRT_SPHERE) {
                kind = K_SPHERE;
                const float r = o.param[0];
                world_box(M, pos, {std::fabs(r), std::fabs(r), std::fabs(r)}, &pb);
                si.a = make_float4(pos.x, pos.y, pos.z, 0.0f);
                si.b = make_float4(r * r, 0.0f, 0.0f, 0.0f);
                fi.A = make_float4(1.0f / r, r, 0.0f, 0.0f);
            } else if (o.kind == MRT_PLANE) {
                kind = K_PLANE;
                finite_prim = false;
                const H3 nraw = {o.param[0], o.param[1], o.param[2]};
                const H3 nh = hnorm(nraw);  // Plane::intersect normalises, rt.rs:404
                // t = -((o_l - pos).n^)/(d_l.n^) with o_l - pos = M(o - pos), d_l = M d  =>  n_w = M^T n^
                const H3 nw = {M.m[0] * nh.x + M.m[3] * nh.y + M.m[6] * nh.z, M.m[1] * nh.x + M.m[4] * nh.y + M.m[7] * nh.z,
                               M.m[2] * nh.x + M.m[5] * nh.y + M.m[8] * nh.z};
                si.a = make_float4(nw.x, nw.y, nw.z, 0.0f);
                si.b = make_float4(hdot(pos, nw), 0.0f, 0.0f, 0.0f);
                const H3 ns = c->normal_space == MRT_NORMAL_OBJECT ? hnorm(nraw) : hnorm(hmul(M, nraw));  // Renderer::normal, rt.rs:786,792
                fi.A = make_float4(ns.x, ns.y, ns.z, 0.0f);
            } else if (o.kind == MRT_BOX) {
                kind = ident ? K_BOX : K_BOX_XF;
                world_box(M, pos, {0.5f * std::fabs(o.param[0]), 0.5f * std::fabs(o.param[1]), 0.5f * std::fabs(o.param[2])}, &pb);
                si.a = make_float4(pos.x, pos.y, pos.z, 0.0f);
                if (ident) {
                    si.a.w = 0.5f * o.param[0];
                    si.b = make_float4(0.5f * o.param[1], 0.5f * o.param[2], 0.0f, 0.0f);
                } else {
                    si.b = make_float4(0.5f * o.param[0], 0.5f * o.param[1], 0.5f * o.param[2], 0.0f);
                    const H3 mp = hmul(M, pos);
                    bxf.push_back({make_float4(M.m[0], M.m[1], M.m[2], -mp.x), make_float4(M.m[3], M.m[4], M.m[5], -mp.y),
                                   make_float4(M.m[6], M.m[7], M.m[8], -mp.z), si.b});
                }
                fi.A = make_float4((1.0f / o.param[0]) * 2.0f, (1.0f / o.param[1]) * 2.0f, (1.0f / o.param[2]) * 2.0f, 0.0f);  // rt.rs:416
            } else {
                kind = K_MESH;
                world_box(M, pos, {meshes[o.mesh].half[0], meshes[o.mesh].half[1], meshes[o.mesh].half[2]}, &pb);
                si.a = make_float4(pos.x, pos.y, pos.z, 0.0f);
                si.b = make_float4(u2f(ident ? 0u : 1u), u2f(o.mesh), 0.0f, 0.0f);
                mesh_m.push_back(x);
                fi.A = make_float4(u2f(meshes[o.mesh].first_tri), 0.0f, 0.0f, 0.0f);
            }
            fi.P = make_float4(pos.x, pos.y, pos.z, u2f(kind | (ident ? FAT_IDENT : 0u) | (textured ? FAT_TEX : 0u) |
                                                        ((!ident && c->normal_space == MRT_NORMAL_FORWARD_XF) ? FAT_NXF : 0u)));
            fi.m0 = make_float4(M.m[0], M.m[1], M.m[2], u2f(pack_ids(mt.tex, mt.rmap)));
            fi.m1 = make_float4(M.m[3], M.m[4], M.m[5], u2f(pack_ids(mt.mmap, mt.gmap)));
            fi.m2 = make_float4(M.m[6], M.m[7], M.m[8], u2f(pack_ids(mt.omap, mt.emap)));
            fi.C = make_float4(mt.albedo[0], mt.albedo[1], mt.albedo[2], mt.emit);
            fi.R = make_float4(mt.rough, mt.metal, mt.glass, mt.opacity);
            if (finite_prim) {
                pb.ref = (kind << 28) | (uint32_t)by_kind[kind].size();
                for (int a = 0; a < 3; a++) prim_boxes_ok &= std::isfinite(pb.lo[a]) && std::isfinite(pb.hi[a]);
                prim_boxes.push_back(pb);
            }
            by_kind[kind].push_back(si);
            fat_k[kind].push_back(fi);
            oi_k[kind].push_back(oi | (k << 16));
        }
    }
    std::vector<FatInst> fat;
    std::vector<uint32_t> obj_inst;
    uint32_t first[K_NKIND], cnt[K_NKIND];
    for (uint32_t k = 0; k < K_NKIND; k++) {
        first[k] = (uint32_t)fat.size();
        cnt[k] = (uint32_t)by_kind[k].size();
        fat.insert(fat.end(), fat_k[k].begin(), fat_k[k].end());
        obj_inst.insert(obj_inst.end(), oi_k[k].begin(), oi_k[k].end());
    }
    if (fat.size() > 0x7fffffffu) return fail(c, MRT_ERR_INVALID, "too many instances");

    CK(cudaStreamSynchronize(c->stream));
    // axis-aligned boxes, two per BoxPair (see mrt_device.cuh)
    std::vector<BoxPair> boxp((by_kind[K_BOX].size() + 1) / 2);
    for (size_t k = 0; k < boxp.size(); k++) {
        const SlimInst& a = by_kind[K_BOX][2 * k];
        SlimInst b{};
        if (2 * k + 1 < by_kind[K_BOX].size()) b = by_kind[K_BOX][2 * k + 1];
        else { b.a = make_float4(0.0f, 0.0f, 0.0f, -1.0f); b.b = make_float4(-1.0f, -1.0f, 0.0f, 0.0f); }  // never hit
        boxp[k].q0 = make_float4(a.a.x, b.a.x, a.a.y, b.a.y);
        boxp[k].q1 = make_float4(a.a.z, b.a.z, a.a.w, b.a.w);
        boxp[k].q2 = make_float4(a.b.x, b.b.x, a.b.y, b.b.y);
    }
    for (uint32_t k = 0; k < K_NKIND; k++) CK(c->d_slim[k].upload(by_kind[k]));
    CK(c->d_boxp.upload(boxp));
    CK(c->d_bxf.upload(bxf));
    // scene-level BVH: only for scenes too large to unroll (the specialised kernel covers <= 128 primitives)
    std::vector<BvhNode> bvh_nodes;
    uint32_t bvh_root = 0;
    // BVH or brute force?  Measured on random scenes of N boxes / N spheres, both through their specialised kernels
    // (unrolled / BVH, Mpaths/s): boxes 48: 11 584 / 9 157, 56: 8 621 / 8 264, 64: 6 695 / 7 307; spheres 16: 26 240 /
    // 24 744, 24: 17 986 / 18 105, 32: 14 179 / 14 699, 40: 10 865 / 12 521, 64: 5 658 / 8 041 — the cross-over sits at
    // ~60 boxes or ~26 spheres, i.e. ~60 box-equivalents with a sphere at 2.3 (a rotated box 2.5, a mesh far more).
    // Minecraft.json (84 boxes): 4 144 unrolled, 5 343 through the BVH.
    size_t bvh_min = 60;
    if (const char* e = std::getenv("MRT_BVH_MIN")) bvh_min = (size_t)std::max(0, std::atoi(e));  // experiment knob
    const size_t brute_cost = (6 * by_kind[K_BOX].size() + 14 * by_kind[K_SPHERE].size() + 15 * bxf.size() + 36 * by_kind[K_MESH].size()) / 6;
    bool use_bvh = brute_cost > bvh_min && prim_boxes.size() > 1 && prim_boxes_ok && prim_boxes.size() < (1u << 28) && !std::getenv("MRT_NO_BVH");
    if (use_bvh) {
        int depth = 0;
        bvh_root = bvh_build(prim_boxes, 0, prim_boxes.size(), &bvh_nodes, 0, &depth);
        if (depth > 30) use_bvh = false;  // the traversal stack holds 32 entries (median splits: never for < 2^28 primitives)
    }
    CK(c->d_bvh.upload(bvh_nodes));
    CK(c->d_mesh_m.upload(mesh_m));
    CK(c->d_fat.upload(fat));
    CK(c->d_tex.upload(tex));
    CK(c->d_texels.upload(texels));
    CK(c->d_mesh.upload(meshes));
    CK(c->d_leaf.upload(leaves));
    CK(c->d_leaf_idx.upload(leaf_idx));
    CK(c->d_tri.upload(tris));
    CK(c->d_tbvh.upload(tbvh));
    CK(c->d_tri_leaf.upload(tri_leaf));
    CK(c->d_obj_inst.upload(obj_inst));

    // ---- text of the scene for the run-time specialised kernel (mrt_jit.cu); small scenes only
    c->jit_header.clear();
    if (c->jit_requested && !c->jit_header.empty()) mrt_jit_wait(c->jit_header);  // never abandon a running compile
    c->jit_kernel = nullptr;
    c->jit_requested = c->jit_failed = c->jit_from_disk = false;
    c->jit_err.clear();
    {
        const size_t n_prim = 2 * boxp.size() + by_kind[K_SPHERE].size() + by_kind[K_PLANE].size() + bxf.size() + by_kind[K_MESH].size();
        // Scenes that go through the BVH get a specialised kernel too, but one that only folds what does not
        // depend on the instance tables (kinds present, material scalars, lights, sky, rotation class): their
        // header has empty tables, so scenes of the same shape share one kernel.
        bool ok = n_prim > 0 && (n_prim <= 128 || use_bvh);
        std::string h = "// generated by mrt_set_scene\n";
        if (use_bvh) h += "#define MRT_JIT_BVH 1\n";
        const bool tables = !use_bvh;
        auto tab = [&](const char* name, size_t n, auto&& row) {
            h += std::string("#define ") + name + "(X)";
            for (size_t k = 0; k < n; k++) { h += " X(" + std::to_string(k); row(k); h += ")"; }
            h += "\n";
        };
        // box pairs: X = packed FFMA2 pair, XS = the two boxes one at a time, X1 = single box (odd count).
        // A pair constant whose two lanes differ costs two uniform-register moves per use in the packed form
        // (only equal lanes are an immediate broadcast), so lopsided pairs are cheaper unpacked.
        h += "#define MRT_JIT_BOXPAIRS(X, XS, X1, CB, CE)";
        // scenes of many boxes: consecutive pairs are bracketed, four at a time, by their bounding box
        // (declaration order is kept, so the first-minimum rule is untouched)
        size_t cluster = 4;
        if (const char* e = std::getenv("MRT_JIT_CLUSTER")) cluster = (size_t)std::max(0, std::atoi(e));  // experiment knob, 0 = off
        const bool clustered = cluster > 0 && boxp.size() >= 3 * cluster;
        for (size_t k = 0; tables && k < boxp.size(); k++) {
            const float* q = &boxp[k].q0.x;  // (cA.x,cB.x, cA.y,cB.y, cA.z,cB.z, hA.x,hB.x, hA.y,hB.y, hA.z,hB.z)
            ok &= all_finite(q, 12);
            const bool odd = 2 * k + 1 >= by_kind[K_BOX].size();
            if (clustered && k % cluster == 0) {
                float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
                for (size_t j = k; j < std::min(boxp.size(), k + cluster); j++) {
                    const float* p = &boxp[j].q0.x;
                    const int lanes = (2 * j + 1 >= by_kind[K_BOX].size()) ? 1 : 2;
                    for (int l = 0; l < lanes; l++)
                        for (int a = 0; a < 3; a++) {
                            lo[a] = std::fmin(lo[a], p[2 * a + l] - std::fabs(p[6 + 2 * a + l]));
                            hi[a] = std::fmax(hi[a], p[2 * a + l] + std::fabs(p[6 + 2 * a + l]));
                        }
                }
                for (int a = 0; a < 3; a++) {  // the cluster test and the box tests round differently: keep a margin
                    const float pad = 1e-5f * (std::fabs(lo[a]) + std::fabs(hi[a])) + 1e-6f;
                    lo[a] -= pad; hi[a] += pad;
                }
                const float v[6] = {lo[0], lo[1], lo[2], hi[0], hi[1], hi[2]};
                ok &= all_finite(v, 6);
                std::string t;
                lits(&t, v, 6);
                h += " CB(" + t.substr(2) + ")";
            }
            int packed = 6, scalar = 12;
            for (int a = 0; a < 3; a++) {
                const float ca = q[2 * a], cb = q[2 * a + 1], ha = q[6 + 2 * a], hb = q[7 + 2 * a];
                if (ca != 0.0f || cb != 0.0f) packed += 1 + (ca != cb ? 2 : 0);
                if (ha != hb) packed += 4;  // +h and -h pairs
                scalar += (ca != 0.0f) + (cb != 0.0f);
            }
            h += odd ? " X1(" : (scalar < packed ? " XS(" : " X(");
            h += std::to_string(k);
            lits(&h, q, 12);
            h += ")";
            if (clustered && (k % cluster == cluster - 1 || k + 1 == boxp.size())) h += " CE";
        }
        h += "\n";
        tab("MRT_JIT_SPHERES", tables ? by_kind[K_SPHERE].size() : 0, [&](size_t k) {
            const SlimInst& e = by_kind[K_SPHERE][k];
            const float v[4] = {e.a.x, e.a.y, e.a.z, e.b.x};
            ok &= all_finite(v, 4); lits(&h, v, 4); });
        tab("MRT_JIT_PLANES", tables ? by_kind[K_PLANE].size() : 0, [&](size_t k) {
            const SlimInst& e = by_kind[K_PLANE][k];
            const float v[4] = {e.a.x, e.a.y, e.a.z, e.b.x};
            ok &= all_finite(v, 4); lits(&h, v, 4); });
        tab("MRT_JIT_BXFS", tables ? bxf.size() : 0, [&](size_t k) { ok &= all_finite(&bxf[k].r0.x, 15); lits(&h, &bxf[k].r0.x, 12); lits(&h, &bxf[k].h.x, 3); });
        tab("MRT_JIT_MESHES", tables ? by_kind[K_MESH].size() : 0, [&](size_t k) {
            const SlimInst& e = by_kind[K_MESH][k];
            ok &= all_finite(&e.a.x, 3) && all_finite(mesh_m[k].m, 12);
            lits(&h, &e.a.x, 3);
            uint32_t rot, mid;
            std::memcpy(&rot, &e.b.x, 4); std::memcpy(&mid, &e.b.y, 4);
            h += ", " + std::to_string(rot) + "u, " + std::to_string(mid) + "u";
            lits(&h, mesh_m[k].m, 12); });
        // big unrolled scenes: without a register budget ptxas hoists every operand (254 registers, 2 blocks
        // per SM on Minecraft.json); 3 blocks (168 registers) measured best there: 2594 -> 2757 Mpaths/s
        // BVH kernels are latency bound (long_scoreboard): 64 registers / 8 blocks per SM measured best
        // (Minecraft.json 5 429 -> 5 628, Instance.json 2 378 -> 2 401 Mpaths/s against ptxas' own 96 / 64)
        if (use_bvh && !(std::getenv("MRT_JIT_MINBLOCKS") && *std::getenv("MRT_JIT_MINBLOCKS"))) h += "#define MRT_JIT_MINBLOCKS 8\n";
        if (tables && n_prim > 48 && !(std::getenv("MRT_JIT_MINBLOCKS") && *std::getenv("MRT_JIT_MINBLOCKS"))) h += "#define MRT_JIT_MINBLOCKS 3\n";
        {   // rough/metal/glass/opacity shared by every material (and no map overrides them): fold them in
            bool uni = s->n_objects > 0;
            const mrt_material& m0 = s->objects[0].mat;
            for (uint32_t oi = 0; oi < s->n_objects && uni; oi++) {
                const mrt_material& m = s->objects[oi].mat;
                uni = m.rough == m0.rough && m.metal == m0.metal && m.glass == m0.glass && m.opacity == m0.opacity &&
                      m.rmap < 0 && m.mmap < 0 && m.gmap < 0 && m.omap < 0;
            }
            const float v[4] = {m0.rough, m0.metal, m0.glass, m0.opacity};
            if (uni && all_finite(v, 4)) {
                h += "#define MRT_JIT_UNIFORM_R ";
                std::string t;
                lits(&t, v, 4);
                h += t.substr(2) + "\n";
            }
        }
        {
            bool binary = true;
            for (uint32_t oi = 0; oi < s->n_objects; oi++) {
                const mrt_material& m = s->objects[oi].mat;
                binary &= (m.emit == 0.0f || m.emit == 1.0f) && m.emap < 0;
            }
            if (binary) h += "#define MRT_JIT_EMIT_BINARY 1\n";
        }
        if (s->sky_color[0] == 0.0f && s->sky_color[1] == 0.0f && s->sky_color[2] == 0.0f) h += "#define MRT_JIT_SKY_BLACK 1\n";
        h += "#define MRT_JIT_ROT " + std::to_string(rot_class) + "\n";
        h += "#define MRT_JIT_N_BOX " + std::to_string(cnt[K_BOX] + cnt[K_BOX_XF]) + "\n";
        h += "#define MRT_JIT_N_SPHERE " + std::to_string(cnt[K_SPHERE]) + "\n";
        if (use_bvh) {  // what the BVH leaves and the loops around the traversal may assume
            h += "#define MRT_JIT_N_ABOX " + std::to_string(cnt[K_BOX]) + "\n";
            h += "#define MRT_JIT_N_BXF " + std::to_string(cnt[K_BOX_XF]) + "\n";
            h += "#define MRT_JIT_N_MESH " + std::to_string(cnt[K_MESH]) + "\n";
            h += "#define MRT_JIT_N_LIGHTS " + std::to_string(s->n_lights) + "\n";
        }
        h += "#define MRT_JIT_N_PLANE " + std::to_string(cnt[K_PLANE]) + "\n";
        h += "#define MRT_JIT_FIRST_SPHERE " + std::to_string(first[K_SPHERE]) + "\n";
        h += "#define MRT_JIT_FIRST_PLANE " + std::to_string(first[K_PLANE]) + "\n";
        h += "#define MRT_JIT_FIRST_BXF " + std::to_string(first[K_BOX_XF]) + "\n";
        h += "#define MRT_JIT_FIRST_MESH " + std::to_string(first[K_MESH]) + "\n";
        if (ok) c->jit_header = h;  // the feature mask is appended in mrt_set_scene's tail
    }

    SceneCommon sc{};
    sc.fat = c->d_fat.p; sc.tex = c->d_tex.p; sc.texels = c->d_texels.p;
    sc.mesh = c->d_mesh.p; sc.leaf = c->d_leaf.p; sc.leaf_idx = c->d_leaf_idx.p; sc.tri = c->d_tri.p;
    sc.tbvh = c->d_tbvh.p; sc.tri_leaf = c->d_tri_leaf.p;
    sc.n_inst = (uint32_t)fat.size();
    sc.n_lights = s->n_lights;
    for (uint32_t k = 0; k < K_NKIND; k++) { sc.first[k] = first[k]; sc.cnt[k] = cnt[k]; }
    for (int k = 0; k < 3; k++) { sc.sky[k] = s->sky_color[k]; sc.sky_tail[k] = s->sky_color[k] * s->sky_pwr; }
    for (uint32_t i = 0; i < s->n_lights; i++) {
        const mrt_light& l = s->lights[i];
        if (l.kind > MRT_LIGHT_DIR) return fail(c, MRT_ERR_INVALID, "unknown light kind");
        H3 v = {l.v[0], l.v[1], l.v[2]};
        if (l.kind == MRT_LIGHT_DIR) { const H3 n = hnorm(v); v = {-n.x, -n.y, -n.z}; }  // rt.rs:977,1031: -dir.norm()
        sc.light[i].v_kind = make_float4(v.x, v.y, v.z, u2f(l.kind));
        sc.light[i].color_pwr = make_float4(l.color[0], l.color[1], l.color[2], l.pwr);
    }
    c->gscene.c = sc;
    c->gscene.boxp = c->d_boxp.p; c->gscene.box = c->d_slim[K_BOX].p; c->gscene.sph = c->d_slim[K_SPHERE].p; c->gscene.pln = c->d_slim[K_PLANE].p;
    c->gscene.bxf = c->d_bxf.p;
    c->gscene.bvh = use_bvh ? c->d_bvh.p : nullptr;
    c->gscene.bvh_root = bvh_root;
    c->gscene.mesh = c->d_slim[K_MESH].p; c->gscene.mesh_m = c->d_mesh_m.p;
    c->in_param = cnt[K_BOX] <= MRT_PB && cnt[K_SPHERE] <= MRT_PS && cnt[K_PLANE] <= MRT_PP && cnt[K_BOX_XF] <= MRT_PX &&
                  cnt[K_MESH] <= MRT_PM && !use_bvh && !std::getenv("MRT_FORCE_GLOBAL_SCENE");
    if (c->in_param) {
        ParamScene& ps = *c->pscene;
        ps.c = sc;
        std::copy(boxp.begin(), boxp.end(), ps.boxp);
        std::copy(by_kind[K_SPHERE].begin(), by_kind[K_SPHERE].end(), ps.sph);
        std::copy(by_kind[K_PLANE].begin(), by_kind[K_PLANE].end(), ps.pln);
        std::copy(bxf.begin(), bxf.end(), ps.bxf);
        std::copy(by_kind[K_MESH].begin(), by_kind[K_MESH].end(), ps.mesh);
        std::copy(mesh_m.begin(), mesh_m.end(), ps.mesh_m);
    }
    if (const char* f = std::getenv("MRT_FORCE_FEATURES")) feat |= (uint32_t)std::atoi(f) & F_ALL;
    c->features = feat;
    if (!c->jit_header.empty()) c->jit_header += "#define MRT_JIT_F " + std::to_string(feat & F_ALL) + "u\n";
    c->have_scene = true;
    return mrt_reset(c);
}

int mrt_set_frame(mrt_ctx* c, const mrt_frame* f) {
    if (!c || !f) return MRT_ERR_INVALID;
    CK(cudaSetDevice(c->device));
    uint32_t nw, nh;
    film_dims(*f, &nw, &nh);
    if (nw == 0 || nh == 0 || f->res[0] == 0 || f->res[1] == 0) return fail(c, MRT_ERR_INVALID, "empty film");
    if ((uint64_t)nw * nh > 0x7fffffffull) return fail(c, MRT_ERR_INVALID, "film too large");
    CK(cudaStreamSynchronize(c->stream));
    c->frame = *f;
    c->nw = nw; c->nh = nh;
    cudaError_t e = c->d_accum.alloc((size_t)nw * nh);
    if (e != cudaSuccess) return cuda_fail(c, e, "accumulator allocation");
    c->weights_ready = false;
    c->have_frame = true;
    return mrt_reset(c);
}

int mrt_set_rt(mrt_ctx* c, uint32_t bounce, float loss, uint64_t seed) {
    if (!c) return MRT_ERR_INVALID;
    if (bounce > 0x3fffffffu) return fail(c, MRT_ERR_INVALID, "bounce too large");
    c->bounce = bounce; c->loss = loss; c->seed = seed;
    return MRT_OK;
}

int mrt_set_option(mrt_ctx* c, uint32_t option, uint32_t value) {
    if (!c) return MRT_ERR_INVALID;
    if (option == MRT_OPT_NORMAL_SPACE && value <= MRT_NORMAL_OBJECT) {
        c->normal_space = value;  // read by the next mrt_set_scene
        return MRT_OK;
    }
    if (option == MRT_OPT_JIT && value <= MRT_JIT_FORCE) {
        c->jit_mode = value;
        return MRT_OK;
    }
    return fail(c, MRT_ERR_INVALID, "unknown option or value");
}

int mrt_set_partition(mrt_ctx* c, uint32_t rank, uint32_t world) {
    if (!c) return MRT_ERR_INVALID;
    if (world == 0 || rank >= world) return fail(c, MRT_ERR_INVALID, "bad partition");
    c->rank = rank; c->world = world;
    return MRT_OK;
}

int mrt_execute_async(mrt_ctx* c, uint32_t n_passes) {
    if (!c) return MRT_ERR_INVALID;
    if (!c->have_scene || !c->have_frame) return fail(c, MRT_ERR_STATE, "execute before set_scene/set_frame");
    CK(cudaSetDevice(c->device));
    FilmParams fp = make_film_params(c);
    // Scene-specialised kernel (mrt_jit.cu).  MRT_JIT_AUTO never stalls a small call: the first
    // execute after set_scene starts the NVRTC compile on a background thread (or finds the cubin in
    // the process / on-disk cache) and this and later calls switch over as soon as it is ready; a call
    // big enough to amortise the ~0.15 s compile (>= 2^33 paths) waits for it.  MRT_JIT_FORCE waits.
    const uint64_t paths = (uint64_t)c->nw * c->nh * n_passes;
    const bool want_jit = !c->jit_header.empty() && c->jit_mode != MRT_JIT_OFF;
    if (want_jit && !c->jit_kernel && !c->jit_failed) {
        MrtJitInfo info;
        c->jit_kernel = mrt_jit_kernel(c->jit_header, c->jit_mode == MRT_JIT_FORCE || paths >= (1ull << 33), &info);
        c->jit_requested = true;
        if (!info.pending) {
            c->jit_seconds = info.seconds;
            c->jit_from_disk = info.from_disk;
            c->jit_err = info.err;
            c->jit_failed = !c->jit_kernel;
            if (c->jit_failed && c->jit_mode == MRT_JIT_FORCE) return fail(c, MRT_ERR_CUDA, "scene specialisation failed: " + c->jit_err);
        }
    }
    const bool use_jit = want_jit && c->jit_kernel;
    uint32_t left = n_passes;
    while (left) {
        const uint32_t n = std::min(left, c->spp_per_launch);
        fp.sample0 = c->rank + c->passes * c->world;
        fp.sample_stride = c->world;
        fp.n_samples = n;
        cudaError_t e = use_jit ? (c->gscene.bvh ? mrt_jit_launch_bvh(c->jit_kernel, c->gscene, fp, c->stream)
                                                 : mrt_jit_launch(c->jit_kernel, c->gscene.c, fp, c->stream))
                                : mrt_launch_path(c->features, c->in_param, c->pscene, &c->gscene, fp, c->stream);
        if (e != cudaSuccess) return cuda_fail(c, e, "path kernel launch");
        if (use_jit) c->jit_launches++;
        c->launches++;
        c->passes += n;
        c->passes_total += n;
        left -= n;
    }
    return MRT_OK;
}

int mrt_sync(mrt_ctx* c) {
    if (!c) return MRT_ERR_INVALID;
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    return MRT_OK;
}

int mrt_execute(mrt_ctx* c, uint32_t n_passes, double* seconds) {
    if (!c) return MRT_ERR_INVALID;
    CK(cudaSetDevice(c->device));
    if (!c->have_scene || !c->have_frame) return fail(c, MRT_ERR_STATE, "execute before set_scene/set_frame");
    CK(cudaEventRecord(c->ev0, c->stream));
    // While the scene-specialised kernel is still compiling (MRT_JIT_AUTO, first big call after
    // mrt_set_scene), a blocking call feeds the generic kernel in slices of ~2^27 paths and looks
    // again after each, so a one-shot render (the CLI case) switches over after ~0.15 s instead of
    // finishing on the slower kernel.  mrt_execute_async cannot wait and keeps what it has.
    uint32_t left = n_passes;
    while (left) {
        uint32_t n = left;
        const bool pending = c->jit_mode == MRT_JIT_AUTO && !c->jit_header.empty() && !c->jit_kernel && !c->jit_failed;
        if (pending) {
            const uint64_t npix = (uint64_t)c->nw * c->nh;
            n = (uint32_t)std::min<uint64_t>(left, std::max<uint64_t>(1, (1ull << 27) / std::max<uint64_t>(1, npix)));
        }
        int rc = mrt_execute_async(c, n);
        if (rc) return rc;
        left -= n;
        if (pending && left) CK(cudaStreamSynchronize(c->stream));
    }
    CK(cudaEventRecord(c->ev1, c->stream));
    CK(cudaEventSynchronize(c->ev1));
    if (seconds) {
        float ms = 0.0f;
        CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        *seconds = (double)ms * 1e-3;
    }
    return MRT_OK;
}

int mrt_film_size(mrt_ctx* c, uint32_t* nw, uint32_t* nh, uint32_t* passes) {
    if (!c) return MRT_ERR_INVALID;
    if (nw) *nw = c->nw;
    if (nh) *nh = c->nh;
    if (passes) *passes = c->passes_total;
    return MRT_OK;
}

int mrt_accum(mrt_ctx* c, float* rgb, uint32_t* passes) {
    if (!c || !rgb) return MRT_ERR_INVALID;
    if (!c->have_frame) return fail(c, MRT_ERR_STATE, "accum before set_frame");
    CK(cudaSetDevice(c->device));
    const uint32_t npix = c->nw * c->nh;
    CK(c->d_rgb.alloc((size_t)npix * 3));
    CK(mrt_launch_unpack(c->d_accum.p, c->d_rgb.p, npix, c->stream));
    c->launches++;
    CK(cudaMemcpyAsync(rgb, c->d_rgb.p, (size_t)npix * 3 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (passes) *passes = c->passes_total;
    return MRT_OK;
}

int mrt_accum_device(mrt_ctx* c, void** dptr, size_t* n_floats, void** cuda_stream) {
    if (!c) return MRT_ERR_INVALID;
    if (!c->have_frame) return fail(c, MRT_ERR_STATE, "accum_device before set_frame");
    if (dptr) *dptr = c->d_accum.p;
    if (n_floats) *n_floats = (size_t)c->nw * c->nh * 4;
    if (cuda_stream) *cuda_stream = (void*)c->stream;
    return MRT_OK;
}

int mrt_set_stream(mrt_ctx* c, void* cuda_stream) {
    if (!c) return MRT_ERR_INVALID;
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    c->stream = cuda_stream ? (cudaStream_t)cuda_stream : c->own_stream;
    return MRT_OK;
}

int mrt_set_passes(mrt_ctx* c, uint32_t passes) {
    if (!c) return MRT_ERR_INVALID;
    c->passes_total = passes;
    return MRT_OK;
}

static int tonemap_ss(mrt_ctx* c) {
    if (!c->have_frame) return fail(c, MRT_ERR_STATE, "img before set_frame");
    if (c->passes_total == 0) return fail(c, MRT_ERR_STATE, "img before any pass");
    const uint32_t npix = c->nw * c->nh;
    CK(c->d_ss.alloc((size_t)npix * 3));
    const float inv_n = 1.0f / (float)c->passes_total;  // Vec3f / f32 = v * (1/n), lin.rs:296-302
    CK(mrt_launch_tonemap(c->d_accum.p, c->d_ss.p, npix, inv_n, c->frame.gamma, c->frame.exp, c->stream));
    c->launches++;
    return MRT_OK;
}

int mrt_img_ss(mrt_ctx* c, uint8_t* rgb) {
    if (!c || !rgb) return MRT_ERR_INVALID;
    CK(cudaSetDevice(c->device));
    int rc = tonemap_ss(c);
    if (rc) return rc;
    CK(cudaMemcpyAsync(rgb, c->d_ss.p, (size_t)c->nw * c->nh * 3, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return MRT_OK;
}

int mrt_img(mrt_ctx* c, uint8_t* rgb) {
    if (!c || !rgb) return MRT_ERR_INVALID;
    CK(cudaSetDevice(c->device));
    int rc = tonemap_ss(c);
    if (rc) return rc;
    const uint32_t w = c->nw, h = c->nh, ow = c->frame.res[0], oh = c->frame.res[1];
    if (ow == w && oh == h) {  // imageops::resize copies when the size is unchanged
        CK(cudaMemcpyAsync(rgb, c->d_ss.p, (size_t)w * h * 3, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        return MRT_OK;
    }
    if (!c->weights_ready) {
        auto taps = [](uint32_t in_n, uint32_t out_n) {
            const float ratio = (float)in_n / (float)out_n;
            const float sr = ratio < 1.0f ? 1.0f : ratio;
            return (uint32_t)std::ceil(2.0f * 3.0f * sr) + 3u;
        };
        c->taps_v = taps(h, oh);
        c->taps_h = taps(w, ow);
        CK(c->d_lv.alloc(oh)); CK(c->d_cv.alloc(oh)); CK(c->d_wv.alloc((size_t)oh * c->taps_v));
        CK(c->d_lh.alloc(ow)); CK(c->d_ch.alloc(ow)); CK(c->d_wh.alloc((size_t)ow * c->taps_h));
        CK(mrt_launch_lanczos_weights(h, oh, c->taps_v, c->d_lv.p, c->d_cv.p, c->d_wv.p, c->stream));
        CK(mrt_launch_lanczos_weights(w, ow, c->taps_h, c->d_lh.p, c->d_ch.p, c->d_wh.p, c->stream));
        c->launches += 2;
        c->weights_ready = true;
    }
    CK(c->d_tmp.alloc((size_t)w * oh * 3));
    CK(c->d_out.alloc((size_t)ow * oh * 3));
    CK(mrt_launch_lanczos_vertical(c->d_ss.p, c->d_tmp.p, w, oh, c->taps_v, c->d_lv.p, c->d_cv.p, c->d_wv.p, c->stream));
    CK(mrt_launch_lanczos_horizontal(c->d_tmp.p, c->d_out.p, w, ow, oh, c->taps_h, c->d_lh.p, c->d_ch.p, c->d_wh.p, c->stream));
    c->launches += 2;
    CK(cudaMemcpyAsync(rgb, c->d_out.p, (size_t)ow * oh * 3, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return MRT_OK;
}

int mrt_trace_primary(mrt_ctx* c, mrt_hit* out) {
    if (!c || !out) return MRT_ERR_INVALID;
    if (!c->have_scene || !c->have_frame) return fail(c, MRT_ERR_STATE, "trace_primary before set_scene/set_frame");
    CK(cudaSetDevice(c->device));
    const uint32_t npix = c->nw * c->nh;
    CK(c->d_hits.alloc(npix));
    FilmParams fp = make_film_params(c);
    CK(mrt_launch_primary(c->gscene, fp, c->d_hits.p, c->d_obj_inst.p, c->stream));
    c->launches++;
    CK(cudaMemcpyAsync(out, c->d_hits.p, (size_t)npix * sizeof(mrt_hit), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return MRT_OK;
}

int mrt_spp_per_launch(mrt_ctx* c, uint32_t spp, uint32_t* current) {
    if (!c) return MRT_ERR_INVALID;
    if (spp) c->spp_per_launch = spp;
    if (current) *current = c->spp_per_launch;
    return MRT_OK;
}

int mrt_jit_status(mrt_ctx* c, uint32_t* eligible, uint32_t* compiled, uint64_t* launches, double* compile_seconds) {
    if (!c) return MRT_ERR_INVALID;
    if (compile_seconds && c->jit_from_disk) *compile_seconds = -c->jit_seconds;  // negative: loaded from the disk cache
    if (eligible) *eligible = c->jit_header.empty() ? 0u : 1u;
    if (compiled) *compiled = c->jit_kernel ? 1u : 0u;
    if (launches) *launches = c->jit_launches;
    if (compile_seconds && !c->jit_from_disk) *compile_seconds = c->jit_seconds;
    if (!c->jit_err.empty()) c->err = c->jit_err;  // readable through mrt_last_error
    return MRT_OK;
}

int mrt_launch_count(mrt_ctx* c, uint64_t* n) {
    if (!c || !n) return MRT_ERR_INVALID;
    *n = c->launches;
    return MRT_OK;
}

int mrt_fp32_peak(mrt_ctx* c, double* tflops, double* seconds) {
    if (!c || !tflops) return MRT_ERR_INVALID;
    CK(cudaSetDevice(c->device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, c->device));
    DevBuf<float> sink;
    CK(sink.alloc(4));
    const int blocks = prop.multiProcessorCount * 8;
    const int iters = 16384;
    CK(mrt_launch_fp32_peak(sink.p, blocks, 256, c->stream));  // warm-up
    CK(cudaStreamSynchronize(c->stream));
    double best = 0.0, best_s = 0.0;
    for (int rep = 0; rep < 3; rep++) {
        CK(cudaEventRecord(c->ev0, c->stream));
        CK(mrt_launch_fp32_peak(sink.p, blocks, iters, c->stream));
        CK(cudaEventRecord(c->ev1, c->stream));
        CK(cudaEventSynchronize(c->ev1));
        float ms = 0.0f;
        CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        const double flops = (double)blocks * 256.0 * (double)iters * 64.0 * 2.0;
        const double tf = flops / ((double)ms * 1e-3) / 1e12;
        if (tf > best) { best = tf; best_s = (double)ms * 1e-3; }
    }
    c->launches += 4;
    sink.release();
    *tflops = best;
    if (seconds) *seconds = best_s;
    return MRT_OK;
}

}  // extern "C"
