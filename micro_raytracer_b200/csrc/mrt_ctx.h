// mrt_ctx.h — the context behind the C ABI of include/mrt.h (shared by mrt_api.cu: entry points, launch
// scheduling, device groups, film read-out; and mrt_scene.cu: validation + packing of a scene into the device layout).
// Host code only.
#pragma once
#include "mrt_device.cuh"
#include "mrt_kernels.h"
#include "mrt_jit.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <deque>
#include <string>
#include <utility>
#include <vector>

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
inline HM transform_of(const float dir[4]) {
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
inline bool is_identity(const HM& m) {
    const float id[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    for (int i = 0; i < 9; i++)
        if (!(m.m[i] == id[i])) return false;
    return true;
}
inline bool finite_m(const HM& m) {
    for (float v : m.m) if (!std::isfinite(v)) return false;
    return true;
}
inline float u2f(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    // Copies `h` to the device on `st` WITHOUT waiting: the caller synchronises the stream before `h` goes away.  An
    // allocation of the right size is reused (a host that re-sends its scene per frame pays copies, not cudaMalloc).
    cudaError_t upload(const std::vector<T>& h, cudaStream_t st) {
        cudaError_t e = alloc(h.size());
        if (e != cudaSuccess) return e;
        if (n) e = cudaMemcpyAsync(p, h.data(), n * sizeof(T), cudaMemcpyHostToDevice, st);
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


// Experiment / test knobs, read from the environment ONCE per context (mrt_create), so that contexts living in
// different threads never see each other's settings change under them.
struct Knobs {
    bool tiled = true;          // MRT_TILE=0: a warp renders 32 pixels of a row instead of an 8x4 tile
    bool bvh_sah = true;        // MRT_BVH_SAH=0: median splits only
    bool mesh_bvh = true;       // MRT_NO_MESH_BVH: the sequential walk of the octree leaves
    bool no_bvh = false;        // MRT_NO_BVH: brute force whatever the scene size
    bool force_global = false;  // MRT_FORCE_GLOBAL_SCENE: never the kernel-parameter scene
    bool jit_minblocks_env = false;  // MRT_JIT_MINBLOCKS set: mrt_jit.cu passes it on, no default from the scene
    size_t bvh_min = 60;        // MRT_BVH_MIN: BVH above this many box-equivalents
    size_t jit_cluster = 4;     // MRT_JIT_CLUSTER: box pairs per bracket in big unrolled scenes, 0 = off
    uint32_t force_features = 0;  // MRT_FORCE_FEATURES
    int refine_spheres = -1;    // MRT_REFINE_SPHERES: -1 decided per scene, 0 never, 1 always (sphere hits in the reference's own arithmetic)
    bool pinhole = true;        // MRT_NO_PINHOLE: frames with aperture 0 go through the thin-lens entry point too (the A/B of the bit-identity test)
    bool mesh_via_bvh = false;  // MRT_MESH_VIA_BVH=1: scenes with a mesh go through the scene BVH whatever their size (measured: Mesh.json 2 355 vs 2 482 unrolled)
    void read() {
        if (const char* s = std::getenv("MRT_TILE")) tiled = std::atoi(s) != 0;
        if (const char* s = std::getenv("MRT_BVH_SAH")) bvh_sah = std::atoi(s) != 0;
        mesh_bvh = !std::getenv("MRT_NO_MESH_BVH");
        pinhole = !std::getenv("MRT_NO_PINHOLE");
        no_bvh = std::getenv("MRT_NO_BVH") != nullptr;
        force_global = std::getenv("MRT_FORCE_GLOBAL_SCENE") != nullptr;
        if (const char* s = std::getenv("MRT_JIT_MINBLOCKS")) jit_minblocks_env = *s != 0;
        if (const char* s = std::getenv("MRT_BVH_MIN")) bvh_min = (size_t)std::max(0, std::atoi(s));
        if (const char* s = std::getenv("MRT_JIT_CLUSTER")) jit_cluster = (size_t)std::max(0, std::atoi(s));
        if (const char* s = std::getenv("MRT_FORCE_FEATURES")) force_features = (uint32_t)std::atoi(s) & F_ALL;
        if (const char* s = std::getenv("MRT_MESH_VIA_BVH")) mesh_via_bvh = std::atoi(s) != 0;
        if (const char* s = std::getenv("MRT_REFINE_SPHERES")) refine_spheres = std::atoi(s);
    }
};

struct mrt_ctx {
    int device = 0;
    Knobs knobs;
    // A GROUP context (mrt_create_group) owns one ordinary context per device and no device buffers of its own: the
    // entry points of mrt_api.cu render through `members` (a plain context renders through itself) and read the
    // film out on members[0].
    std::vector<mrt_ctx*> members;
    bool p2p = false;                   // group: every member can map every other member's memory (NVLink / PCIe P2P)
    DevBuf<float4> d_stage;             // group without P2P: staging copy of a peer's accumulator on members[0]
    // Pass coalescing (MRT_OPT_COALESCE): one-pass calls of mrt_execute are queued here and rendered in launches of
    // up to spp_per_launch passes, the way the reference's `for _ in 0..sample { execute }` loop (cli.rs:162) wants.
    bool coalesce = true;
    uint32_t pending = 0;
    // Device time of the path launches.  One Round per run of launches: a (start, stop) event pair per rendering
    // context, read back lazily; `unreported_s` is what mrt_execute's *seconds has not handed out yet.
    struct Round { std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev; };
    std::deque<Round> timing;
    std::vector<cudaEvent_t> event_pool;  // per device: events are recycled, not created per launch
    double unreported_s = 0.0, total_s = 0.0;
    cudaEvent_t ev_sync = nullptr, ev_band = nullptr;  // cross-device ordering inside a group
    // film gather across processes (mrt_ipc_*): the other ranks' accumulators and the film rank's image, mapped over CUDA IPC
    uint32_t ipc_rank = 0, ipc_world = 0;
    std::vector<void*> ipc_mapped;      // what cudaIpcOpenMemHandle returned (closed on detach)
    PeerAccums ipc_accums{};            // by rank; this rank's own entry is its local pointer
    uint8_t* ipc_image = nullptr;       // the film rank's supersampled image (local on the film rank)
    uint64_t scene_hash = 0;            // content hash of the scene the context holds (mrt_update_scene)
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
    uint32_t passes = 0;        // passes this context (or group) launched: its next pass is number `passes` of its partition
    uint32_t passes_total = 0;  // passes the accumulator holds (launched here + summed in by an external reduce)
    uint32_t spp_per_launch = 1024;  // measured: 128 -> 8917, 256 -> 9058, 1024 -> 9234 Mpaths/s (intra-warp tail)
    uint32_t normal_space = MRT_NORMAL_FORWARD_XF;  // MRT_OPT_NORMAL_SPACE

    // run-time scene specialisation (mrt_jit.cu)
    uint32_t jit_mode = MRT_JIT_AUTO;   // MRT_OPT_JIT
    std::string jit_header;             // "" = scene not eligible
    const MrtJitKernels* jit_kernel = nullptr;  // compiled for jit_header (both entry points)
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


inline int fail(mrt_ctx* c, int code, const std::string& m) { c->err = m; return code; }
inline int cuda_fail(mrt_ctx* c, cudaError_t e, const char* what) {
    c->err = std::string(what) + ": " + cudaGetErrorString(e);
    return MRT_ERR_CUDA;
}
#define CK(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return cuda_fail(c, e__, #call); } while (0)

// mrt_scene.cu: validate `s`, pack it into the device layout and replace the scene of the (single-device) context.
// On failure the context holds NO scene (have_scene == false) and the message is in c->err.
int mrt_scene_upload(mrt_ctx* c, const mrt_scene* s);
// content hash of a scene description + the options that shape its packing (mrt_update_scene)
uint64_t mrt_scene_hash(const mrt_scene* s, uint32_t normal_space);
