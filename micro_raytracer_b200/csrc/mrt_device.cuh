// mrt_device.cuh — device-side scene layout and the per-ray building blocks of the
// path-tracing hot path (sm_100a).  Reference semantics: /root/reference/src/rt.rs, lin.rs
// (cited per function).  Design notes in DESIGN.md.
//
// Scene storage, two views of the same instance table:
//   * SlimInst  (32 B/instance): what the closest-hit loop reads.  Every lane of a warp walks
//     the same instance list, so these are warp-uniform loads: from the kernel-parameter
//     constant bank (ParamView, scenes up to MRT_PARAM_INST instances; operands come straight
//     from c[0][..]) or from global memory through L1 (GlobalView, any size).
//   * FatInst   (128 B/instance = one L1 line): what a lane reads about the ONE instance it
//     hit (transform, normal data, material); divergent, always global memory.
#pragma once
#ifdef __CUDACC_RTC__  // compiled at run time by NVRTC (mrt_jit.cu): no host headers
typedef unsigned char uint8_t;
typedef unsigned short uint16_t;
typedef unsigned int uint32_t;
typedef int int32_t;
typedef unsigned long long uint64_t;
#else
#include <cstdint>
#include <cuda_runtime.h>
#endif

#define MRT_E 0.0001f  // rt.rs:7

// MRT_CHECKED (tools/build_variant.sh checked "-DMRT_CHECKED=1"): every table index the kernels form is checked against
// the table's size and a violation traps the kernel (the launch then fails with a CUDA error the tests see).  This
// pool's GPUs do not admit compute-sanitizer; the checked build run over the whole GPU test suite is the substitute.
#ifdef MRT_CHECKED
#define MRT_CHECK(cond) do { if (!(cond)) __trap(); } while (0)
#else
#define MRT_CHECK(cond) do { } while (0)
#endif

enum : uint32_t {
    K_BOX = 0, K_SPHERE = 1, K_PLANE = 2, K_BOX_XF = 3, K_MESH = 4, K_NKIND = 5
};
// feature bits of a scene -> kernel specialisation
enum : uint32_t {
    F_LIGHTS = 1u,    // scene has lights: shadow rays + direct shading
    F_TEX = 2u,       // some material has a texture map: uv + texel fetches
    F_TRANSMIT = 4u,  // some material has opacity < 1 or an omap: exit hit + refraction
    F_MESH = 8u,      // some object is a mesh
    F_ALL = 15u
};

// The closest-hit loops walk one table per primitive kind (no per-instance dispatch), each in
// declaration order; the FatInst table is sorted the same way (kind, then declaration order).
struct SlimInst {
    // box    : a = (centre.xyz, half.x)  b.xy = half.yz            (identity transform; 6 floats = 2 uniform loads)
    // sphere : a.xyz = centre            b.x = r^2
    // plane  : a.xyz = n_w = M^T n^     b.x = pos . n_w          (t = -(o.n_w - b.x)/(d.n_w))
    // box_xf : a.xyz = pos               b.xyz = half              (+ its Xf at the same index)
    // mesh   : a.xyz = pos               b.x = 1 if rotated (bits), b.y = mesh index (bits)   (+ its Xf)
    float4 a, b;
};
struct Xf { float m[12]; };  // rows of M = rot_y * look (rt.rs:726-727), padded to 3 x float4
// Two axis-aligned boxes A, B interleaved for the packed FFMA2 slab test (one f32x2 lane each):
//   q0 = (cA.x, cB.x, cA.y, cB.y)  q1 = (cA.z, cB.z, hA.x, hB.x)  q2 = (hA.y, hB.y, hA.z, hB.z)
// c = centre, h = half extents.  An odd box count is padded with B = a box of half extents -1,
// which can never be hit (its t0 > t1 for every ray).
struct BoxPair { float4 q0, q1, q2; };
// Rotated box: rows of M with -(M pos)_i in .w  (o_l - pos = M o - M pos), and the half extents.
struct BxfInst { float4 r0, r1, r2, h; };

enum : uint32_t { FAT_IDENT = 0x100u, FAT_TEX = 0x200u, FAT_NXF = 0x400u /* kind normal goes through M again, rt.rs:792 */ };
struct FatInst {
    float4 P;          // pos.xyz, w = kind | FAT_IDENT | FAT_TEX (bits)
    float4 A;          // box: 2/size   sphere: (1/r, r, 0)   plane: shading normal norm(M n)   mesh: x = first_tri (bits)
    float4 C;          // albedo.rgb, emit
    float4 R;          // rough, metal, glass, opacity
    float4 m0, m1, m2; // rows of M; .w = packed texture ids (tex|rmap<<16, mmap|gmap<<16, omap|emap<<16), 0xFFFF = none
    float4 S;          // spare
};
static_assert(sizeof(FatInst) == 128, "FatInst must be one cache line");

// Scene-level BVH over the FINITE instances (everything but planes) of scenes too large to unroll
// (> 128 primitives).  It only narrows the candidate set: every candidate runs the same
// primitive test as the brute-force loop and the winner is the lexicographic minimum of (t0, instance
// index), i.e. exactly the brute-force result.  Children of an inner node are adjacent.
// A node holds the boxes of BOTH its children, so one visit is one round of loads (the node's own box was
// tested at its parent).  The two boxes are stored like a BoxPair — centre / half extents, left child in the
// low f32x2 lane, right child in the high one — so both slab intervals come out of 9 FFMA2 + 4 FMNMX3 with no
// per-axis min/max (a lo/hi node costs 12 FFMA + 12 FMNMX + 4 FMNMX3, and the half-rate ALU pipe is what these
// kernels wait for):  q0 = (cL.x, cR.x, cL.y, cR.y)  q1 = (cL.z, cR.z, hL.x, hR.x)  q2 = (hL.y, hR.y, hL.z, hR.z)
// ref.x / ref.y = child references: bit 31 set = leaf, the low bits are the primitive (scene BVH: kind << 28 |
// index within the kind's table; mesh BVH: triangle index within the mesh), else the index of the child node.
// One primitive per leaf, so leaves need no node at all.
// 48 bytes, three 16-byte words: the BVH kernels are bound by the L1 data pipe (ncu, Instance.json: l1tex__data_pipe_lsu_wavefronts
// 81 % of peak at 64 % issue; the node reads of the divergent walks are 4/5 of the wavefronts, and a 256-bit load costs two), so
// the six half extents are stored as fp16 ROUNDED UP — a node box may only grow: the candidate set stays a superset and the
// results bit-identical — and a visit reads three words instead of four:
//   w0 = (cL.x, cR.x, cL.y, cR.y)   w1 = (cL.z, cR.z, h2(hL.x, hR.x), h2(hL.y, hR.y))   w2 = (h2(hL.z, hR.z), 0, ref.x, ref.y)
struct BvhNode { float4 w0; float2 cz; uint32_t hx, hy; uint32_t hz, pad, refl, refr; };
static_assert(sizeof(BvhNode) == 48, "a BVH node is three 16-byte words");
#define MRT_BVH_LEAF 0x80000000u
struct DLight { float4 v_kind; float4 color_pwr; };  // v.xyz (pos or unit -dir), w = kind bits ; color.rgb, pwr
struct DTex { uint32_t w, h, first, has_dat; };      // texel offset into the float4 texel array
struct DMeshLeaf { float4 lo, hi; };                  // leaf box relative to instance pos; lo.w = first index (bits), hi.w = count (bits)
struct DMesh { uint32_t first_leaf, n_leaf, first_tri, n_tri; float half[3]; uint32_t bvh_root; };  // half = half extents of the root AABB; bvh_root = 0xffffffff: no triangle BVH
// v0, e0 = v1 - v0, e1 = v2 - v0 (object space, before + pos); v0.w / e0.w (bits) = first entry / entry count of the
// triangle's DTriLeaf list
struct DTri { float4 v0, e0, e1, pad; };  // 64 bytes: read as one 32-byte word (v0, e0) + e1
static_assert(sizeof(DTri) == 64, "DTri is read with 32-byte loads");
// One octree leaf that lists a triangle: `leaf` indexes SceneCommon::leaf, `rank` is the position of this
// occurrence in the reference's candidate sequence (leaf order, then list order).  Ascending rank per triangle.
struct DTriLeaf { uint32_t leaf, rank; };

// capacities of the kernel-parameter scene (constant bank); larger scenes use GlobalScene
#define MRT_PB 160  // axis-aligned boxes
#define MRT_PS 96   // spheres
#define MRT_PP 32   // planes
#define MRT_PX 24   // rotated boxes
#define MRT_PM 8    // mesh instances
#define MRT_MAX_LIGHTS 16

struct SceneCommon {
    const FatInst* fat;   // sorted by (kind, declaration order): index = first[kind] + k
    const DTex* tex;
    const float4* texels;
    const DMesh* mesh;
    const DMeshLeaf* leaf;
    const uint32_t* leaf_idx;
    const DTri* tri;
    const BvhNode* tbvh;         // triangle BVHs of all meshes (boxes relative to the instance pos); root reference in DMesh::bvh_root
    const DTriLeaf* tri_leaf;    // per triangle: the octree leaves that list it
    uint32_t n_inst, n_lights;
    uint32_t n_tex, n_texels, n_tri, n_leaf, n_leaf_idx, n_tri_leaf, n_tbvh, n_bvh, n_mesh;  // table sizes (read by MRT_CHECK only)
    uint32_t refine_spheres;  // sphere hits are recomputed in the reference's own arithmetic (refine_sphere_hit)
    uint32_t cnt[K_NKIND];    // instances per kind
    uint32_t first[K_NKIND];  // FatInst index of the kind's first instance
    float sky[3];       // sky.color (primary miss, rt.rs:958)
    float sky_tail[3];  // sky.color * sky.pwr (rt.rs:964)
    DLight light[MRT_MAX_LIGHTS];
};
struct ParamScene {
    SceneCommon c;
    BoxPair boxp[MRT_PB / 2];
    SlimInst sph[MRT_PS];
    SlimInst pln[MRT_PP];
    BxfInst bxf[MRT_PX];
    SlimInst mesh[MRT_PM];
    Xf mesh_m[MRT_PM];
};
static_assert(sizeof(ParamScene) + 256 < 32764, "kernel parameters are limited to 32764 bytes");
struct GlobalScene {
    SceneCommon c;
    const BoxPair* boxp;
    const SlimInst* box;       // the same axis-aligned boxes one by one (a = centre.xyz, half.x; b.xy = half.yz): what a BVH leaf reads
    const SlimInst* sph;
    const SlimInst* pln;
    const BxfInst* bxf;
    const SlimInst* mesh;
    const Xf* mesh_m;
    const BvhNode* bvh;        // nullptr: brute force
    uint32_t bvh_root;         // reference of the root (a leaf when the scene has one finite instance)
};

struct FilmParams {
    float4* accum;
    uint32_t nw, nh;
    uint32_t sample0, sample_stride, n_samples;  // this launch: global samples sample0 + j*stride
    uint32_t key;
    uint32_t max_bounce;
    float keep;           // 1 - min(loss, 1)   (rt.rs:571)
    float cam_pos[3];
    float aprt, foc;
    float cam_M[9];       // rot_y(cam.dir) * lookat(cam.dir, up)   (rt.rs:925-930)
    float fw, fh;         // res * ssaa as f32 (rt.rs:938-939)
    float fy;             // 1 / (2 tan(fov/2))  (rt.rs:902-906)
    uint32_t cam_identity;  // cam_M == I: skip the rotation
    uint32_t tiles_x;       // > 0: a warp renders an 8x4 pixel tile, a block of 128 threads 16x8 (tiles_x blocks per row); 0: 32 pixels of a row
};
// number of 128-thread blocks a path-kernel launch needs for this film
__host__ __device__ inline uint32_t path_grid_blocks(const FilmParams& fp) {
    return fp.tiles_x ? fp.tiles_x * ((fp.nh + 7u) / 8u) : (fp.nw * fp.nh + 127u) / 128u;
}

// ------------------------------------------------------------------ small vector helpers
struct f3 { float x, y, z; };
__device__ __forceinline__ f3 mk(float x, float y, float z) { return {x, y, z}; }
__device__ __forceinline__ f3 operator+(f3 a, f3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ f3 operator-(f3 a, f3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ f3 operator-(f3 a) { return {-a.x, -a.y, -a.z}; }
__device__ __forceinline__ f3 operator*(f3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ f3 operator*(f3 a, f3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
__device__ __forceinline__ float dot(f3 a, f3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }
__device__ __forceinline__ f3 cross(f3 a, f3 b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
__device__ __forceinline__ f3 fma3(f3 a, float s, f3 b) { return {fmaf(a.x, s, b.x), fmaf(a.y, s, b.y), fmaf(a.z, s, b.z)}; }
// MRT_PRECISE (ablation build, tools/build_variant.sh precise): IEEE reciprocal / square root / division and libm's
// sincos instead of the approximate units (MUFU.RCP / MUFU.RSQ / MUFU.SIN), to separate what the approximations cost
// in parity against the oracle from what the reference's own ill-conditioning costs ("precision ablation").
#ifdef MRT_PRECISE
__device__ __forceinline__ f3 normalize(f3 a) { return a * __frcp_rn(__fsqrt_rn(dot(a, a))); }  // lin.rs:64-66: self * mag().recip()
#else
__device__ __forceinline__ f3 normalize(f3 a) { return a * rsqrtf(dot(a, a)); }  // lin.rs:64-66
#endif
__device__ __forceinline__ f3 xyz(float4 v) { return {v.x, v.y, v.z}; }
__device__ __forceinline__ f3 mulM(const float4& r0, const float4& r1, const float4& r2, f3 v) {
    return {dot(xyz(r0), v), dot(xyz(r1), v), dot(xyz(r2), v)};
}
__device__ __forceinline__ f3 mulXf(const Xf& x, f3 v) {
    return {fmaf(x.m[2], v.z, fmaf(x.m[1], v.y, x.m[0] * v.x)),
            fmaf(x.m[6], v.z, fmaf(x.m[5], v.y, x.m[4] * v.x)),
            fmaf(x.m[10], v.z, fmaf(x.m[9], v.y, x.m[8] * v.x))};
}
__device__ __forceinline__ float frcp(float x) {  // one MUFU.RCP
#ifdef MRT_PRECISE
    return __frcp_rn(x);
#endif
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// 256-bit read-only load (sm_100: LDG.E.256, ld.global.nc.v8): the BVH kernels are bound by the L1 data pipe, not by issue
// slots (ncu, Instance.json: l1tex__data_pipe_lsu_wavefronts 82 % of peak at 64 % issue) — every lane of a divergent walk reads
// its own node, so a load instruction costs one wavefront per active lane whatever its width; 32 bytes per instruction halve
// the wavefronts of a node visit (64 B), of a leaf (32 B) and of the hit's FatInst rows.  `p` must be 32-byte aligned.
struct f8 { float4 lo, hi; };
__device__ __forceinline__ f8 ldg256(const void* p) {
    f8 r;
#ifdef MRT_NO_LD256
    r.lo = __ldg(reinterpret_cast<const float4*>(p)); r.hi = __ldg(reinterpret_cast<const float4*>(p) + 1);
#else
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(r.lo.x), "=f"(r.lo.y), "=f"(r.lo.z), "=f"(r.lo.w), "=f"(r.hi.x), "=f"(r.hi.y), "=f"(r.hi.z), "=f"(r.hi.w) : "l"(p));
#endif
    return r;
}

// Packed f32x2 arithmetic (sm_100 FFMA2): one issue slot for two FMAs.  Measured on B200
// (tools/microbench.cu): same 74 TFLOP/s as FFMA in half the issue slots; operands may be
// register pairs, uniform-register pairs or a scalar broadcast to both lanes.
typedef unsigned long long f2;
__device__ __forceinline__ f2 pk2(float lo, float hi) { f2 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ f2 bc2(float v) { return pk2(v, v); }
__device__ __forceinline__ void up2(f2 v, float& lo, float& hi) { asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { f2 d; asm("fma.rn.ftz.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

// Box::intersect's reciprocal with the zero-division workaround, rt.rs:303-316
__device__ __forceinline__ float rcp_fixed(float d) {
    float m = frcp(d);
    return (fabsf(m) == __int_as_float(0x7f800000)) ? (1.0f / MRT_E) : m;
}
__device__ __forceinline__ f3 rcp_fixed3(f3 d) { return {rcp_fixed(d.x), rcp_fixed(d.y), rcp_fixed(d.z)}; }
// The geometrically true slab reciprocal for acceleration-structure nodes, given m = rcp_fixed3(d): a zero
// component becomes a huge finite slope (inside the slab -> (-huge, +huge), outside -> an empty interval)
// instead of the reference's 1/E, which belongs to the primitive tests only.
__device__ __forceinline__ f3 true_rcp3(f3 d, f3 m) {
    // 1e18: huge against any ray parameter, small enough that (centre - origin) * slope cannot overflow
    return {d.x == 0.0f ? 1e18f : m.x, d.y == 0.0f ? 1e18f : m.y, d.z == 0.0f ? 1e18f : m.z};
}

// Slab intervals (entry, exit) of the two children of a BVH node for a ray given as bm = true_rcp3 reciprocal,
// bnom = -o * bm, bam = |bm| (mesh BVHs: o relative to the instance).
struct NodeRay { f3 bm, bnom, bam; };
__device__ __forceinline__ void node_slabs(const NodeRay& n, float4 q0, float4 q1, float4 q2,
                                           float* tnl, float* tfl, float* tnr, float* tfr) {
    const f2 cx = fma2(pk2(q0.x, q0.y), bc2(n.bm.x), bc2(n.bnom.x));
    const f2 cy = fma2(pk2(q0.z, q0.w), bc2(n.bm.y), bc2(n.bnom.y));
    const f2 cz = fma2(pk2(q1.x, q1.y), bc2(n.bm.z), bc2(n.bnom.z));
    const f2 hx = pk2(q1.z, q1.w), hy = pk2(q2.x, q2.y), hz = pk2(q2.z, q2.w);
    float lxa, lxb, lya, lyb, lza, lzb, hxa, hxb, hya, hyb, hza, hzb;
    up2(fma2(hx, bc2(-n.bam.x), cx), lxa, lxb);
    up2(fma2(hy, bc2(-n.bam.y), cy), lya, lyb);
    up2(fma2(hz, bc2(-n.bam.z), cz), lza, lzb);
    up2(fma2(hx, bc2(n.bam.x), cx), hxa, hxb);
    up2(fma2(hy, bc2(n.bam.y), cy), hya, hyb);
    up2(fma2(hz, bc2(n.bam.z), cz), hza, hzb);
    *tnl = fmaxf(fmaxf(lxa, lya), lza); *tfl = fminf(fminf(hxa, hya), hza);
    *tnr = fmaxf(fmaxf(lxb, lyb), lzb); *tfr = fminf(fminf(hxb, hyb), hzb);
}

// fp16 pair -> two floats (cvt.f32.f16 x 2; no cuda_fp16.h under NVRTC)
__device__ __forceinline__ void half2_to_floats(uint32_t u, float& lo, float& hi) {
    asm("{\n\t.reg .b16 a, b;\n\tmov.b32 {a, b}, %2;\n\tcvt.f32.f16 %0, a;\n\tcvt.f32.f16 %1, b;\n\t}" : "=f"(lo), "=f"(hi) : "r"(u));
}
// One node visit's data: three 16-byte loads, the half extents widened to f32, in node_slabs' operand layout.
__device__ __forceinline__ void load_node(const BvhNode* p, float4* q0, float4* q1, float4* q2, uint2* ref) {
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(p));
    const float4 w1 = __ldg(reinterpret_cast<const float4*>(p) + 1);
    const uint4 w2 = __ldg(reinterpret_cast<const uint4*>(p) + 2);
    *q0 = w0;
    q1->x = w1.x; q1->y = w1.y;
    half2_to_floats(__float_as_uint(w1.z), q1->z, q1->w);
    half2_to_floats(__float_as_uint(w1.w), q2->x, q2->y);
    half2_to_floats(w2.x, q2->z, q2->w);
    *ref = make_uint2(w2.z, w2.w);
}

// ------------------------------------------------------------------ RNG: pcg4d counter hash
// (Jarzynski & Olano, JCGT 2020).  One call = the 4 uniforms of (pixel, sample, block);
// identical to oracle/mrt_oracle.cpp rng_block so both consume the same numbers.
// The four 32-bit words of a block; a uniform is word * MRT_U32_TO_UNIT in [0, 1 - 2^-24] (the scale
// is (1 - 2^-24) 2^-32 so that a word that rounds up to 2^32 still maps below 1).
#define MRT_U32_TO_UNIT 0x1.fffffep-33f
// u < 0.8f  <=>  word < MRT_LOTTERY_80 (exact: float(word) * MRT_U32_TO_UNIT is monotone in word)
#define MRT_LOTTERY_80 0xCCCCCD80u
__device__ __forceinline__ uint4 rng_words(uint32_t pixel, uint32_t sample, uint32_t block, uint32_t key) {
    uint32_t x = pixel * 1664525u + 1013904223u;
    uint32_t y = sample * 1664525u + 1013904223u;
    uint32_t z = block * 1664525u + 1013904223u;
    uint32_t w = key * 1664525u + 1013904223u;
    x += y * w; y += z * x; z += x * y; w += y * z;
    x ^= x >> 16; y ^= y >> 16; z ^= z >> 16; w ^= w >> 16;
    x += y * w; y += z * x; z += x * y; w += y * z;
    return make_uint4(x, y, z, w);
}
__device__ __forceinline__ float4 rng_block(uint32_t pixel, uint32_t sample, uint32_t block, uint32_t key) {
    const uint4 v = rng_words(pixel, sample, block, key);
    return make_float4((float)v.x * MRT_U32_TO_UNIT, (float)v.y * MRT_U32_TO_UNIT, (float)v.z * MRT_U32_TO_UNIT, (float)v.w * MRT_U32_TO_UNIT);
}

// Camera block: the 2 lens uniforms of (pixel, sample).  The lens jitter only needs a cheap
// hash: lowbias32 (Wellons) of the sample index offset by a per-pixel seed, split into two
// 16-bit uniforms (lens positions on a 65536^2 grid).  Identical in oracle/mrt_oracle.cpp.
__device__ __forceinline__ uint32_t lowbias32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du;
    x ^= x >> 15; x *= 0x846ca68bu;
    x ^= x >> 16;
    return x;
}
__device__ __forceinline__ uint32_t cam_hash_seed(uint32_t pixel, uint32_t key) { return lowbias32(pixel * 0x9E3779B1u + key); }
__device__ __forceinline__ float2 rng_cam(uint32_t seed, uint32_t sample) {
    const uint32_t h = lowbias32(seed + sample * 0x85EBCA6Bu);
    const float s = 1.0f / 65536.0f;
    return make_float2((float)(h >> 16) * s, (float)(h & 0xffffu) * s);
}

// ------------------------------------------------------------------ scene views
struct ParamView {
    static constexpr bool kParam = true;
    static constexpr bool kJit = false;
    static constexpr bool kBvh = false;
    const ParamScene& s;
    __device__ __forceinline__ const SceneCommon& c() const { return s.c; }
    __device__ __forceinline__ BoxPair boxp(uint32_t k) const { return s.boxp[k]; }
    __device__ __forceinline__ SlimInst sph(uint32_t k) const { return s.sph[k]; }
    __device__ __forceinline__ SlimInst pln(uint32_t k) const { return s.pln[k]; }
    __device__ __forceinline__ BxfInst bxf(uint32_t k) const { return s.bxf[k]; }
    __device__ __forceinline__ SlimInst mesh(uint32_t k) const { return s.mesh[k]; }
    __device__ __forceinline__ const Xf& mesh_m(uint32_t k) const { return s.mesh_m[k]; }
};
__device__ __forceinline__ SlimInst ldg_slim(const SlimInst* p) { const f8 v = ldg256(p); return {v.lo, v.hi}; }
struct GlobalView {
    static constexpr bool kParam = false;
    static constexpr bool kJit = false;
    static constexpr bool kBvh = true;
    const GlobalScene& s;
    __device__ __forceinline__ const SceneCommon& c() const { return s.c; }
    __device__ __forceinline__ BoxPair boxp(uint32_t k) const { return {__ldg(&s.boxp[k].q0), __ldg(&s.boxp[k].q1), __ldg(&s.boxp[k].q2)}; }
    __device__ __forceinline__ SlimInst sph(uint32_t k) const { return ldg_slim(s.sph + k); }
    __device__ __forceinline__ SlimInst pln(uint32_t k) const { return ldg_slim(s.pln + k); }
    __device__ __forceinline__ BxfInst bxf(uint32_t k) const { return {__ldg(&s.bxf[k].r0), __ldg(&s.bxf[k].r1), __ldg(&s.bxf[k].r2), __ldg(&s.bxf[k].h)}; }
    __device__ __forceinline__ SlimInst mesh(uint32_t k) const { return ldg_slim(s.mesh + k); }
    __device__ __forceinline__ const Xf& mesh_m(uint32_t k) const { return s.mesh_m[k]; }
};

// Scene-specialised view (mrt_jit.cu): the instance tables are not data at all — the generated
// header spells every primitive test out with literal operands (MRT_JIT_* X-macros), so the
// closest-hit "loops" are straight-line code whose constants are FFMA/FFMA2 immediates.
struct JitView {
    static constexpr bool kParam = true;
    static constexpr bool kJit = true;
    static constexpr bool kBvh = false;
    const SceneCommon& s;
    __device__ __forceinline__ const SceneCommon& c() const { return s; }
};

// ------------------------------------------------------------------ primitive tests
// Möller–Trumbore as written in rt.rs:361-398 (two-sided, |det| < E rejected, t >= 0).
__device__ __forceinline__ bool tri_test(const DTri& tr, f3 o_rel /* ray.orig - pos */, f3 d, float* t_out) {
    f3 e0 = xyz(tr.e0), e1 = xyz(tr.e1);
    f3 p = cross(d, e1);
    float det = dot(e0, p);
    if (det < MRT_E && det > -MRT_E) return false;
    float inv = frcp(det);
    f3 t = o_rel - xyz(tr.v0);
    // The acceptance tests are written so that a NaN fails them (the reference's `if u < 0.0 || u > 1.0 { return None }`
    // lets NaN through: a ray that comes back from 1e30 away — a grazing hit on an infinite plane — would "hit"
    // every triangle with t = NaN).  For finite values the two forms are the same predicate.
    float u = dot(t, p) * inv;
    if (!(u >= 0.0f && u <= 1.0f)) return false;
    f3 q = cross(t, e0);
    float v = dot(d, q) * inv;
    if (!(v >= 0.0f && (u + v) <= 1.0f)) return false;
    float tt = dot(e1, q) * inv;
    if (!(tt >= 0.0f)) return false;
    *t_out = tt;
    return true;
}

// The reference's Box::intersect on an octree leaf (relative to the instance), 1/E quirk included: m = rcp_fixed3(d),
// om = o_rel * m.  Part of the candidate rule: a zero direction component behaves like a slope of 1/E, so a leaf
// the ray runs inside of can still be "missed".
__device__ __forceinline__ bool leaf_slab_hit(float4 lo, float4 hi, f3 m, f3 om) {
    const float ax = fmaf(lo.x, m.x, -om.x), bx = fmaf(hi.x, m.x, -om.x);
    const float ay = fmaf(lo.y, m.y, -om.y), by = fmaf(hi.y, m.y, -om.y);
    const float az = fmaf(lo.z, m.z, -om.z), bz = fmaf(hi.z, m.z, -om.z);
    const float tn = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz));
    const float tf = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz));
    return !(tn > tf || tf < 0.0f);
}
// Is a triangle the ray hits a CANDIDATE of the reference's walk, i.e. does the ray pierce one of the octree leaves
// that list it?  rf / rl = rank (position in the candidate sequence) of its first / last pierced occurrence.
template <bool WANT_LAST>
__device__ __forceinline__ bool tri_candidate(const SceneCommon& c, const DTri& tr, f3 m, f3 om, uint32_t* rf, uint32_t* rl) {
    const uint32_t e0 = __float_as_uint(tr.v0.w), en = __float_as_uint(tr.e0.w);
    bool cand = false;
    for (uint32_t e = 0; e < en; e++) {
        MRT_CHECK(e0 + e < c.n_tri_leaf);
        const DTriLeaf tl = c.tri_leaf[e0 + e];
        MRT_CHECK(tl.leaf < c.n_leaf);
        if (!leaf_slab_hit(__ldg(&c.leaf[tl.leaf].lo), __ldg(&c.leaf[tl.leaf].hi), m, om)) continue;
        if (!cand) *rf = tl.rank;
        *rl = tl.rank;
        cand = true;
        if (!WANT_LAST) break;
    }
    return cand;
}

// Mesh leg of Renderer::intersect (rt.rs:740-772), triangle-BVH form (DMesh::bvh_root; the sequential walk of the
// octree leaves it replaces follows below and stays as the reference the tests compare with).  The reference's result only depends on the SET of
// candidates — every triangle listed by a pierced leaf — and, between equal t, on their order.  So instead of
// walking every pierced leaf's list (Mesh.json: ~200 leaf slab tests + ~50 triangle tests per ray), a tight BVH
// over the triangles finds the few triangles the ray actually hits, and each HIT is then checked for candidacy
// against the handful of octree leaves that list that triangle (DTriLeaf: same slab arithmetic as below, so the
// holes of the vertex-containment lists, Q15, are reproduced exactly).  Entry = lexicographic minimum of
// (t, rank of the first pierced occurrence), exit = maximum of (t, rank of the last): the first-min / last-max
// rules of the sequential walk, independent of the visiting order.  Node boxes are padded on the host so that
// rounding cannot hide a hit; pruning by the best entry is only allowed when no exit is wanted.
template <bool ANY, bool WANT_T1>
__device__ __forceinline__ bool mesh_test_bvh(const SceneCommon& c, const DMesh& mh, f3 o_rel, f3 d, f3 m, f3 om,
                                              float* t0, float* t1, int* i0, int* i1) {
    const float INF = __int_as_float(0x7f800000);
    float b0 = INF, b1 = -INF;
    uint32_t r0 = 0xffffffffu, r1 = 0u;
    int k0 = -1, k1 = -1;
    uint2 stack[32];  // (reference, entry parameter of the pushed subtree's box): one 8-byte word per entry
    int sp = 0;
    // Two slab tests.  `leaf_slab` is the reference's Box::intersect on an octree leaf, 1/E quirk included
    // (a zero direction component behaves like a slope of 1/E, so a leaf the ray runs inside of can still be
    // "missed": part of the candidate rule).  `slab` is for the BVH's own nodes and must be geometrically
    // true, or a ray running exactly in a mesh's symmetry plane would lose the triangles that end there.
    auto leaf_slab = [&](float4 lo, float4 hi) -> bool {
        const float ax = fmaf(lo.x, m.x, -om.x), bx = fmaf(hi.x, m.x, -om.x);
        const float ay = fmaf(lo.y, m.y, -om.y), by = fmaf(hi.y, m.y, -om.y);
        const float az = fmaf(lo.z, m.z, -om.z), bz = fmaf(hi.z, m.z, -om.z);
        const float tn = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz));
        const float tf = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz));
        return !(tn > tf || tf < 0.0f);
    };
    NodeRay nr;
    nr.bm = true_rcp3(d, m);
    nr.bnom = mk(-o_rel.x * nr.bm.x, -o_rel.y * nr.bm.y, -o_rel.z * nr.bm.z);
    nr.bam = mk(fabsf(nr.bm.x), fabsf(nr.bm.y), fabsf(nr.bm.z));
    constexpr bool PRUNE = !ANY && !WANT_T1;
    uint32_t cur = mh.bvh_root;
    for (;;) {
        if (cur & MRT_BVH_LEAF) {
            const uint32_t ti = cur & ~MRT_BVH_LEAF;
            MRT_CHECK(ti < mh.n_tri && mh.first_tri + ti < c.n_tri);
            const DTri* tp = &c.tri[mh.first_tri + ti];
            DTri tr;
            { const f8 ve = ldg256(tp); tr.v0 = ve.lo; tr.e0 = ve.hi; tr.e1 = __ldg(&tp->e1); }
            float t;
            if (tri_test(tr, o_rel, d, &t) && !(PRUNE && !(t <= b0))) {
                // is the triangle a candidate?  first / last pierced leaf that lists it
                const uint32_t e0 = __float_as_uint(tr.v0.w), en = __float_as_uint(tr.e0.w);
                uint32_t rf = 0xffffffffu, rl = 0u;
                bool cand = false;
                for (uint32_t e = 0; e < en; e++) {
                    const DTriLeaf tl = c.tri_leaf[e0 + e];
                    const f8 lh = ldg256(&c.leaf[tl.leaf]);  // lo, hi
                    if (!leaf_slab(lh.lo, lh.hi)) continue;
                    if (!cand) rf = tl.rank;
                    rl = tl.rank;
                    cand = true;
                    if (!WANT_T1) break;
                }
                if (cand) {
                    if (ANY) return true;
                    if (t < b0 || (t == b0 && rf < r0)) { b0 = t; r0 = rf; k0 = (int)ti; }
                    if (WANT_T1) { if (t > b1 || (t == b1 && rl >= r1)) { b1 = t; r1 = rl; k1 = (int)ti; } }
                }
            }
        } else {
            MRT_CHECK(cur < c.n_tbvh);
            float4 q0, q1, q2;
            uint2 ref;
            load_node(&c.tbvh[cur], &q0, &q1, &q2, &ref);
            float tl, tfl, tr, tfr;
            node_slabs(nr, q0, q1, q2, &tl, &tfl, &tr, &tfr);
            bool hl = !(tl > tfl || tfl < 0.0f), hr = !(tr > tfr || tfr < 0.0f);
            if (PRUNE) { hl = hl && tl <= b0; hr = hr && tr <= b0; }
            const uint32_t cl = ref.x, cr = ref.y;
            if (hl && hr) {
                const bool left_first = tl <= tr;
                MRT_CHECK(sp < 32);
                if (sp < 32) { stack[sp] = make_uint2(left_first ? cr : cl, __float_as_uint(left_first ? tr : tl)); sp++; }
                cur = left_first ? cl : cr;
                continue;
            }
            if (hl || hr) { cur = hl ? cl : cr; continue; }
        }
        // next subtree; one that starts behind the best entry found since it was pushed holds nothing closer
        bool more = false;
        while (sp > 0) {
            --sp;
            const uint2 e = stack[sp];
            if (!PRUNE || __uint_as_float(e.y) <= b0) { cur = e.x; more = true; break; }
        }
        if (!more) break;
    }
    if (ANY || k0 < 0) return false;
    *t0 = b0; *i0 = k0;
    if (WANT_T1) { *t1 = b1; *i1 = k1; } else { *t1 = b0; *i1 = k0; }
    return true;
}

// Mesh leg of Renderer::intersect, rt.rs:740-772, over the flattened depth-3 octree: the root box
// must be pierced (rt.rs:708-710), then every non-empty leaf the ray pierces contributes its triangle
// list (rt.rs:707-723); entry = first minimum t, exit = last maximum t.  o_rel = object-space origin
// minus instance pos.
//
// Two phases per chunk of 32 leaves, because the lanes of a warp pierce different leaves: (1) all
// lanes slab-test the chunk's leaves in lockstep and keep a bit per pierced leaf; (2) every lane
// walks ITS OWN pierced leaves and tests one triangle per iteration, so the warp runs for the
// longest lane's candidate list, not for the union of all lanes' leaves (ncu on Mesh.json: the
// leaf-major loop issued 2/3 of its instructions with <= 3 active lanes).  Candidate order per lane
// is unchanged (leaf order, then list order), so the first-min / last-max tie rules hold.
// (Only taken when the mesh has no triangle BVH: MRT_NO_MESH_BVH, the bit-identity tests.)
// root AABB of a mesh, centred on the instance (Mesh::gen_aabb, rt.rs:261-270; tested first, rt.rs:708-710)
__device__ __forceinline__ bool mesh_root_hit(const DMesh& mh, f3 m, f3 om) {
    const float ax = fabsf(m.x) * mh.half[0], ay = fabsf(m.y) * mh.half[1], az = fabsf(m.z) * mh.half[2];
    const float tn = fmaxf(fmaxf(-om.x - ax, -om.y - ay), -om.z - az);
    const float tf = fminf(fminf(-om.x + ax, -om.y + ay), -om.z + az);
    return !(tn > tf || tf < 0.0f);
}
template <bool ANY, bool WANT_T1>
__device__ __forceinline__ bool mesh_leaf_walk(const SceneCommon& c, const DMesh& mh, f3 o_rel, f3 d, f3 m, f3 om,
                                               float* t0, float* t1, int* i0, int* i1) {
    bool any = false;
    float b0 = 0.f, b1 = 0.f;
    int k0 = -1, k1 = -1;
    for (uint32_t base = 0; base < mh.n_leaf; base += 32u) {
        const uint32_t n = min(32u, mh.n_leaf - base);
        uint32_t mask = 0u;
        for (uint32_t l = 0; l < n; l++) {
            const float4 lo = __ldg(&c.leaf[mh.first_leaf + base + l].lo);
            const float4 hi = __ldg(&c.leaf[mh.first_leaf + base + l].hi);
            const float ax = fmaf(lo.x, m.x, -om.x), bx = fmaf(hi.x, m.x, -om.x);
            const float ay = fmaf(lo.y, m.y, -om.y), by = fmaf(hi.y, m.y, -om.y);
            const float az = fmaf(lo.z, m.z, -om.z), bz = fmaf(hi.z, m.z, -om.z);
            const float tn = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz));
            const float tf = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz));
            if (!(tn > tf || tf < 0.0f)) mask |= 1u << l;
        }
        uint32_t first = 0u, cnt = 0u, k = 0u;
        for (;;) {
            if (k == cnt) {  // next pierced leaf of this lane (leaves are never empty)
                if (mask == 0u) break;
                const uint32_t l = (uint32_t)__ffs((int)mask) - 1u;
                mask &= mask - 1u;
                first = __float_as_uint(__ldg(&c.leaf[mh.first_leaf + base + l].lo.w));
                cnt = __float_as_uint(__ldg(&c.leaf[mh.first_leaf + base + l].hi.w));
                k = 0u;
            }
            MRT_CHECK(first + k < c.n_leaf_idx);
            const uint32_t ti = __ldg(&c.leaf_idx[first + k]);
            MRT_CHECK(mh.first_tri + ti < c.n_tri);
            k++;
            const DTri* tp = &c.tri[mh.first_tri + ti];
            DTri tr;
            tr.v0 = __ldg(&tp->v0); tr.e0 = __ldg(&tp->e0); tr.e1 = __ldg(&tp->e1);
            float t;
            if (!tri_test(tr, o_rel, d, &t)) continue;
            if (!any) { b0 = b1 = t; k0 = k1 = (int)ti; any = true; continue; }
            if (t < b0) { b0 = t; k0 = (int)ti; }
            if (t >= b1) { b1 = t; k1 = (int)ti; }
        }
    }
    if (!any) return false;
    *t0 = b0; *t1 = b1; *i0 = k0; *i1 = k1;
    return true;
}
template <bool ANY, bool WANT_T1>
__device__ __forceinline__ bool mesh_test(const SceneCommon& c, uint32_t mesh_id, f3 o_rel, f3 d,
                                          float* t0, float* t1, int* i0, int* i1) {
    MRT_CHECK(mesh_id < c.n_mesh);
    const DMesh mh = c.mesh[mesh_id];
    const f3 m = rcp_fixed3(d);
    const f3 om = o_rel * m;
    if (!mesh_root_hit(mh, m, om)) return false;
    if (mh.bvh_root != 0xffffffffu) return mesh_test_bvh<ANY, WANT_T1>(c, mh, o_rel, d, m, om, t0, t1, i0, i1);
    return mesh_leaf_walk<ANY, WANT_T1>(c, mh, o_rel, d, m, om, t0, t1, i0, i1);
}

struct HitRec {
    float t0, t1;
    int inst;        // declaration-order instance index, -1 = miss
    int tri0, tri1;  // mesh triangle of entry / exit
};

// RayTracer::closest_hit, rt.rs:867-898 (without the normals): brute force over every
// instance; the first minimum of t0 wins (rt.rs:872).  Box t0 may be negative (rt.rs:327-331),
// sphere rejects t0 < 0 (rt.rs:353), plane needs t > 0 (rt.rs:407).
// ANY = occlusion query for shadow rays (rt.rs:1036: "is there any hit at all").
//
// Instances are grouped by kind, one loop per kind, declaration order inside a kind (a tie
// between two kinds would need bit-equal results of two different formulas — a rounding
// coincidence in the reference as well — so only the in-kind order is kept).  Each loop is a
// Duff's device: the first n % 8 entries run through a fall-through switch whose entries have
// COMPILE-TIME offsets — with the scene in the kernel-parameter constant bank (ParamView) they
// need no load, no index arithmetic and no loop control, the constants are FFMA operands — then
// chunks of 8 with a running base.  The switch runs its entries in descending order, hence its
// `<=`; the chunks ascend with `<`.
struct Best {
    float t0, t1;
    int bi, tr0, tr1;
    bool any;
};

// hit  <=>  t0 <= t1 and t1 >= 0 (Box) — folded with "closer than the best so far" into
// t0 < min(t1, best) and t1 >= 0: FMNMX + FSETP + FSETP.AND + 2 selects.  (t0 == t1, a ray grazing
// an edge exactly, counts as a miss here; measure zero.)
template <uint32_t F, bool ANY, bool WANT_T1, bool LE>
__device__ __forceinline__ void best_update_slab(Best& B, float t0, float t1, int idx) {
    if constexpr (ANY) {
        B.any |= (t0 <= t1) && (t1 >= 0.0f);
    } else if constexpr (WANT_T1 || (F & F_MESH) != 0) {
        const bool closer = LE ? (t0 <= B.t0) : (t0 < B.t0);
        if (fmaxf(t0, 0.0f) <= t1 && closer) {
            B.t0 = t0; B.bi = idx;
            if constexpr (WANT_T1) B.t1 = t1;
            if constexpr ((F & F_MESH) != 0) { B.tr0 = -1; B.tr1 = -1; }
        }
    } else {
        if constexpr (LE)
            asm("{\n\t.reg .pred p, q;\n\t.reg .f32 tm;\n\t"
                "min.ftz.f32 tm, %3, %0;\n\t"
                "setp.ge.ftz.f32 p, %3, 0f00000000;\n\t"
                "setp.le.and.ftz.f32 q, %2, tm, p;\n\t"
                "selp.f32 %0, %2, %0, q;\n\t"
                "selp.b32 %1, %4, %1, q;\n\t}"
                : "+f"(B.t0), "+r"(B.bi) : "f"(t0), "f"(t1), "r"(idx));
        else {
#ifndef MRT_T1_SIGN_ON_ALU_PIPE  // t1 >= 0 folded into the minimum through u = t1 * inf (+inf / -inf / NaN for t1 > / < / = 0;
            // min ignores a NaN): one FMUL on the FMA pipe + FMNMX3 instead of FMNMX + FSETP on the half-rate ALU pipe
            // (same predicate; headline +0.4 %: 13 969 -> 14 021 Mpaths/s, two runs each)
            asm("{\n\t.reg .pred q;\n\t.reg .f32 tm, u;\n\t"
                "mul.ftz.f32 u, %3, 0f7F800000;\n\t"
                "min.ftz.f32 tm, %3, %0;\n\t"
                "min.ftz.f32 tm, tm, u;\n\t"
                "setp.lt.ftz.f32 q, %2, tm;\n\t"
                "selp.f32 %0, %2, %0, q;\n\t"
                "selp.b32 %1, %4, %1, q;\n\t}"
                : "+f"(B.t0), "+r"(B.bi) : "f"(t0), "f"(t1), "r"(idx));
#else
            asm("{\n\t.reg .pred p, q;\n\t.reg .f32 tm;\n\t"
                "min.ftz.f32 tm, %3, %0;\n\t"
                "setp.ge.ftz.f32 p, %3, 0f00000000;\n\t"
                "setp.lt.and.ftz.f32 q, %2, tm, p;\n\t"
                "selp.f32 %0, %2, %0, q;\n\t"
                "selp.b32 %1, %4, %1, q;\n\t}"
                : "+f"(B.t0), "+r"(B.bi) : "f"(t0), "f"(t1), "r"(idx));
#endif
        }
    }
}

template <uint32_t F, bool ANY, bool WANT_T1, bool LE>
__device__ __forceinline__ void best_update(Best& B, bool hit, float t0, float t1, int idx, int tr0, int tr1) {
    if constexpr (ANY) {
        B.any |= hit;
    } else {
        const bool closer = LE ? (t0 <= B.t0) : (t0 < B.t0);
        if (hit && closer) {
            B.t0 = t0; B.bi = idx;
            if constexpr (WANT_T1) B.t1 = t1;
            if constexpr ((F & F_MESH) != 0) { B.tr0 = tr0; B.tr1 = tr1; }
        }
    }
}

struct RayPre { f3 o, d, m, nom, am, nam; NodeRay n; };  // m = 1/d (Box::intersect's fix-up applied), nom = -o*m, am = |m|, nam = -|m|; bm, bnom: true_rcp3 form for BVH nodes

// Box::intersect, rt.rs:299-333, centre/half form: n = (o - pos) m, k = half |m|,
// t0 = max(-n - k), t1 = min(-n + k); miss iff t0 > t1 or t1 < 0.  Two boxes per call, one in
// each f32x2 lane: 9 FFMA2 + 4 FMNMX3 for the pair.  idx = index of box A (B = idx + 1).
// c * m + n for a pair of centre coordinates; FOLD (specialised kernel: the operands are literals)
// drops the multiply when both are zero — the compiler may not (0 * inf), the scene guarantees it.
template <bool FOLD>
__device__ __forceinline__ f2 centre_term(float ca, float cb, float m, float n) {
    if constexpr (FOLD) { if (ca == 0.0f && cb == 0.0f) return bc2(n); }
    return fma2(pk2(ca, cb), bc2(m), bc2(n));
}
template <uint32_t F, bool ANY, bool WANT_T1, bool LE, bool FOLD = false>
__device__ __forceinline__ void test_box_pair(Best& B, const RayPre& r, const BoxPair e, int idx) {
    const f2 cx = centre_term<FOLD>(e.q0.x, e.q0.y, r.m.x, r.nom.x);
    const f2 cy = centre_term<FOLD>(e.q0.z, e.q0.w, r.m.y, r.nom.y);
    const f2 cz = centre_term<FOLD>(e.q1.x, e.q1.y, r.m.z, r.nom.z);
    const f2 hx = pk2(e.q1.z, e.q1.w), hy = pk2(e.q2.x, e.q2.y), hz = pk2(e.q2.z, e.q2.w);
    float lxa, lxb, lya, lyb, lza, lzb, hxa, hxb, hya, hyb, hza, hzb;
    if constexpr (FOLD) {  // literal half extents: negate them at compile time instead of -|m| at run time
        up2(fma2(pk2(-e.q1.z, -e.q1.w), bc2(r.am.x), cx), lxa, lxb);
        up2(fma2(pk2(-e.q2.x, -e.q2.y), bc2(r.am.y), cy), lya, lyb);
        up2(fma2(pk2(-e.q2.z, -e.q2.w), bc2(r.am.z), cz), lza, lzb);
    } else {
        up2(fma2(hx, bc2(r.nam.x), cx), lxa, lxb);
        up2(fma2(hy, bc2(r.nam.y), cy), lya, lyb);
        up2(fma2(hz, bc2(r.nam.z), cz), lza, lzb);
    }
    up2(fma2(hx, bc2(r.am.x), cx), hxa, hxb);
    up2(fma2(hy, bc2(r.am.y), cy), hya, hyb);
    up2(fma2(hz, bc2(r.am.z), cz), hza, hzb);
    const float t0a = fmaxf(fmaxf(lxa, lya), lza), t1a = fminf(fminf(hxa, hya), hza);
    const float t0b = fmaxf(fmaxf(lxb, lyb), lzb), t1b = fminf(fminf(hxb, hyb), hzb);
    if constexpr (LE) {  // descending traversal: B first, ties go to the lower index
        best_update_slab<F, ANY, WANT_T1, LE>(B, t0b, t1b, idx + 1);
        best_update_slab<F, ANY, WANT_T1, LE>(B, t0a, t1a, idx);
    } else {
        best_update_slab<F, ANY, WANT_T1, LE>(B, t0a, t1a, idx);
        best_update_slab<F, ANY, WANT_T1, LE>(B, t0b, t1b, idx + 1);
    }
}
// One box with literal operands (specialised kernel): used where packing two boxes would cost more
// uniform-register moves than it saves (lanes of the pair with different constants), and for the odd box.
template <uint32_t F, bool ANY, bool WANT_T1>
__device__ __forceinline__ void test_box_lit(Best& B, const RayPre& r, float px, float py, float pz, float hx, float hy, float hz, int idx) {
    const float cx = px == 0.0f ? r.nom.x : fmaf(px, r.m.x, r.nom.x);
    const float cy = py == 0.0f ? r.nom.y : fmaf(py, r.m.y, r.nom.y);
    const float cz = pz == 0.0f ? r.nom.z : fmaf(pz, r.m.z, r.nom.z);
    const float t0 = fmaxf(fmaxf(fmaf(-hx, r.am.x, cx), fmaf(-hy, r.am.y, cy)), fmaf(-hz, r.am.z, cz));
    const float t1 = fminf(fminf(fmaf(hx, r.am.x, cx), fmaf(hy, r.am.y, cy)), fmaf(hz, r.am.z, cz));
    best_update_slab<F, ANY, WANT_T1, false>(B, t0, t1, idx);
}
// Sphere::intersect, rt.rs:335-359, with a = d.d = 1 (directions are unit), half-b form
template <uint32_t F, bool ANY, bool WANT_T1, bool LE>
__device__ __forceinline__ void test_sphere(Best& B, const RayPre& r, const SlimInst e, int idx) {
    const f3 oc = r.o - xyz(e.a);
    const float hb = dot(oc, r.d);
    const float cc = fmaf(oc.z, oc.z, fmaf(oc.y, oc.y, fmaf(oc.x, oc.x, -e.b.x)));
    const float disc = fmaf(hb, hb, -cc);
    const float sq = sqrtf(fmaxf(disc, 0.0f));
    const float t0 = -hb - sq;
    best_update<F, ANY, WANT_T1, LE>(B, (disc >= 0.0f) && (t0 >= 0.0f), t0, sq - hb, idx, -1, -1);
}
// Plane::intersect, rt.rs:400-412, instance transform folded into n_w
template <uint32_t F, bool ANY, bool WANT_T1, bool LE>
__device__ __forceinline__ void test_plane(Best& B, const RayPre& r, const SlimInst e, int idx) {
    const float t0 = (e.b.x - dot(r.o, xyz(e.a))) * frcp(dot(r.d, xyz(e.a)));
    best_update<F, ANY, WANT_T1, LE>(B, t0 > 0.0f, t0, t0, idx, -1, -1);
}
// rotated box: ray into object space (rt.rs:726-733), slab test about the origin.  Origin and
// direction ride in the two f32x2 lanes: (o_l - pos, d_l)_i = M_i0 (o.x, d.x) + M_i1 (o.y, d.y)
// + M_i2 (o.z, d.z) + (-(M pos)_i, 0): 9 FFMA2 for both transforms.
struct RayPk { f2 x, y, z; };  // (o.x, d.x), (o.y, d.y), (o.z, d.z)
// acc + c * p; FOLD drops terms whose (literal) matrix entry is zero, e.g. the z row/column of a yaw-only instance
template <bool FOLD>
__device__ __forceinline__ f2 row_term(float c, f2 p, f2 acc) {
    if constexpr (FOLD) { if (c == 0.0f) return acc; }
    return fma2(bc2(c), p, acc);
}
template <uint32_t F, bool ANY, bool WANT_T1, bool LE, bool FOLD = false>
__device__ __forceinline__ void test_bxf(Best& B, const RayPk& p, const BxfInst e, int idx, float mz = 0.0f /* the ray's 1/d.z, FOLD only */) {
    float olx, dlx, oly, dly, olz, dlz;
    up2(row_term<FOLD>(e.r0.z, p.z, row_term<FOLD>(e.r0.y, p.y, row_term<FOLD>(e.r0.x, p.x, pk2(e.r0.w, 0.0f)))), olx, dlx);
    up2(row_term<FOLD>(e.r1.z, p.z, row_term<FOLD>(e.r1.y, p.y, row_term<FOLD>(e.r1.x, p.x, pk2(e.r1.w, 0.0f)))), oly, dly);
    bool z_through = false;  // yaw-only instance: the z row of M is (0, 0, 1), so d_l.z = d.z and 1/d_l.z is the ray's own
    if constexpr (FOLD) z_through = e.r2.x == 0.0f && e.r2.y == 0.0f && e.r2.z == 1.0f;
    if (z_through) { float oz, dz; up2(p.z, oz, dz); olz = oz + e.r2.w; dlz = dz; }
    else up2(row_term<FOLD>(e.r2.z, p.z, row_term<FOLD>(e.r2.y, p.y, row_term<FOLD>(e.r2.x, p.x, pk2(e.r2.w, 0.0f)))), olz, dlz);
    f3 ml;
    if (z_through) { ml.x = rcp_fixed(dlx); ml.y = rcp_fixed(dly); ml.z = mz; }
    else ml = rcp_fixed3(mk(dlx, dly, dlz));
    const float cx = -olx * ml.x, cy = -oly * ml.y, cz = -olz * ml.z;
    const float ax = fabsf(ml.x), ay = fabsf(ml.y), az = fabsf(ml.z);
    const float t0 = fmaxf(fmaxf(fmaf(-e.h.x, ax, cx), fmaf(-e.h.y, ay, cy)), fmaf(-e.h.z, az, cz));
    const float t1 = fminf(fminf(fmaf(e.h.x, ax, cx), fmaf(e.h.y, ay, cy)), fmaf(e.h.z, az, cz));
    best_update_slab<F, ANY, WANT_T1, LE>(B, t0, t1, idx);
}
template <uint32_t F, bool ANY, bool WANT_T1, bool LE>
__device__ __forceinline__ void test_mesh(Best& B, const SceneCommon& c, const RayPre& r, const SlimInst e, const Xf& x, int idx) {
    f3 ol = r.o - xyz(e.a), dl = r.d;
    if (__float_as_uint(e.b.x) != 0u) { ol = mulXf(x, ol); dl = mulXf(x, r.d); }
    float t0 = 0.0f, t1 = 0.0f;
    int tr0 = -1, tr1 = -1;
    const bool hit = mesh_test<ANY, WANT_T1>(c, __float_as_uint(e.b.y), ol, dl, &t0, &t1, &tr0, &tr1);
    best_update<F, ANY, WANT_T1, LE>(B, hit, t0, t1, idx, tr0, tr1);
}

// ---- scene-level BVH (GlobalView only).  LEX = true makes every update lexicographic in
// (t0, instance index), because the BVH visits candidates in no particular order.
template <uint32_t F, bool ANY, bool WANT_T1>
__device__ __forceinline__ void best_update_lex(Best& B, bool hit, float t0, float t1, int idx, int tr0, int tr1) {
    if constexpr (ANY) {
        B.any |= hit;
    } else {
#ifdef MRT_LEX_BRANCHY
        if (hit && (t0 < B.t0 || (t0 == B.t0 && idx < B.bi))) {
#else
        // ('&' / '|': with '&&' / '||' the compiler emits three divergent branches for this test)
        if (hit & ((t0 < B.t0) | ((t0 == B.t0) & (idx < B.bi)))) {
#endif
            B.t0 = t0; B.bi = idx;
            if constexpr (WANT_T1) B.t1 = t1;
            if constexpr ((F & F_MESH) != 0) { B.tr0 = tr0; B.tr1 = tr1; }
        }
    }
}
// The same interval with the primitive boxes' own arithmetic (r.m: the 1/E quirk included).  For a bracket
// around AXIS-ALIGNED BOXES ONLY this is exactly monotone — every box interval computed with the same
// formula lies inside its bracket's — so the specialised kernel's cluster brackets use this cheaper form.
__device__ __forceinline__ bool bracket_hit(const RayPre& r, float4 lo, float4 hi, float best) {
    const float ax = fmaf(lo.x, r.m.x, r.nom.x), bx = fmaf(hi.x, r.m.x, r.nom.x);
    const float ay = fmaf(lo.y, r.m.y, r.nom.y), by = fmaf(hi.y, r.m.y, r.nom.y);
    const float az = fmaf(lo.z, r.m.z, r.nom.z), bz = fmaf(hi.z, r.m.z, r.nom.z);
    const float tn = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz));
    const float tf = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz));
    return tn <= tf && tf >= 0.0f && tn <= best;
}
// What the specialised kernel of a BVH scene knows about the scene's shape (mrt_api.cu: MRT_JIT_BVH header):
// kinds that do not occur cost nothing in the leaves, plane and light loops have literal trip counts.
#if defined(MRT_JIT) && defined(MRT_JIT_BVH)
#define MRT_BVH_HAS_ABOX (MRT_JIT_N_ABOX > 0)
#define MRT_BVH_HAS_SPHERE (MRT_JIT_N_SPHERE > 0)
#define MRT_BVH_HAS_BXF (MRT_JIT_N_BXF > 0)
#define MRT_BVH_HAS_MESH (MRT_JIT_N_MESH > 0)
#define MRT_N_PLANES(c) ((uint32_t)MRT_JIT_N_PLANE)
#define MRT_N_LIGHTS(c) ((uint32_t)MRT_JIT_N_LIGHTS)
#define MRT_BVH_ALWAYS true   // no brute-force path in this kernel
#else
#define MRT_BVH_HAS_ABOX true
#define MRT_BVH_HAS_SPHERE true
#define MRT_BVH_HAS_BXF true
#define MRT_BVH_HAS_MESH true
#define MRT_N_PLANES(c) ((c).cnt[K_PLANE])
#define MRT_N_LIGHTS(c) ((c).n_lights)
#define MRT_BVH_ALWAYS false
#endif
template <uint32_t F, bool ANY, bool WANT_T1>
__device__ __forceinline__ void bvh_leaf(Best& B, const GlobalScene& s, const RayPre& r, const RayPk& rp, uint32_t ref) {
    const SceneCommon& c = s.c;
    // (table indices stay below 2^24 — the scene BVH is only built for fewer primitives, mrt_scene.cu — so the byte offset
    // k * sizeof(entry) is formed in 32 bits: one IMAD.WIDE instead of a 64-bit shift, 8 -> 3 instructions per leaf visit)
    const uint32_t kind = ref >> 28, k = ref & 0x00ffffffu;
    MRT_CHECK(kind < K_NKIND && kind != K_PLANE && k < c.cnt[kind]);
    float t0 = 0.f, t1 = 0.f;
    int tr0 = -1, tr1 = -1;
    bool hit;
    constexpr bool only_abox = MRT_BVH_HAS_ABOX && !MRT_BVH_HAS_SPHERE && !MRT_BVH_HAS_BXF && !MRT_BVH_HAS_MESH;
    constexpr bool only_sphere = MRT_BVH_HAS_SPHERE && !MRT_BVH_HAS_ABOX && !MRT_BVH_HAS_BXF && !MRT_BVH_HAS_MESH;
    if (MRT_BVH_HAS_ABOX && (only_abox || kind == K_BOX)) {  // one lane of a BoxPair, scalar form of test_box_pair
        const SlimInst e = ldg_slim(s.box + k);  // two 16-byte loads; same arithmetic as one lane of test_box_pair
        const float cx = fmaf(e.a.x, r.m.x, r.nom.x), cy = fmaf(e.a.y, r.m.y, r.nom.y), cz = fmaf(e.a.z, r.m.z, r.nom.z);
        const float hx = e.a.w, hy = e.b.x, hz = e.b.y;
        t0 = fmaxf(fmaxf(fmaf(hx, r.nam.x, cx), fmaf(hy, r.nam.y, cy)), fmaf(hz, r.nam.z, cz));
        t1 = fminf(fminf(fmaf(hx, r.am.x, cx), fmaf(hy, r.am.y, cy)), fmaf(hz, r.am.z, cz));
        hit = fmaxf(t0, 0.0f) <= t1 && t0 < t1;  // as best_update_slab: t0 == t1 (edge graze) is a miss
        best_update_lex<F, ANY, WANT_T1>(B, hit, t0, t1, (int)k, -1, -1);
    } else if (MRT_BVH_HAS_SPHERE && (only_sphere || kind == K_SPHERE)) {
        const SlimInst e = ldg_slim(s.sph + k);
        const f3 oc = r.o - xyz(e.a);
        const float hb = dot(oc, r.d);
        const float cc = fmaf(oc.z, oc.z, fmaf(oc.y, oc.y, fmaf(oc.x, oc.x, -e.b.x)));
        const float disc = fmaf(hb, hb, -cc);
        const float sq = sqrtf(fmaxf(disc, 0.0f));
        t0 = -hb - sq;
        best_update_lex<F, ANY, WANT_T1>(B, (disc >= 0.0f) && (t0 >= 0.0f), t0, sq - hb, (int)(c.first[K_SPHERE] + k), -1, -1);
    } else if (MRT_BVH_HAS_BXF && kind == K_BOX_XF) {
        Best L; L.t0 = __int_as_float(0x7f800000); L.t1 = 0.f; L.bi = -1; L.tr0 = L.tr1 = -1; L.any = false;
        test_bxf<F, false, true, false>(L, rp, {__ldg(&s.bxf[k].r0), __ldg(&s.bxf[k].r1), __ldg(&s.bxf[k].r2), __ldg(&s.bxf[k].h)}, 0);
        best_update_lex<F, ANY, WANT_T1>(B, L.bi == 0, L.t0, L.t1, (int)(c.first[K_BOX_XF] + k), -1, -1);
    } else if (MRT_BVH_HAS_MESH) {
        if constexpr ((F & F_MESH) != 0) {
            const SlimInst e = ldg_slim(s.mesh + k);
            f3 ol = r.o - xyz(e.a), dl = r.d;
            if (__float_as_uint(e.b.x) != 0u) { ol = mulXf(s.mesh_m[k], ol); dl = mulXf(s.mesh_m[k], r.d); }
            hit = mesh_test<ANY, WANT_T1>(c, __float_as_uint(e.b.y), ol, dl, &t0, &t1, &tr0, &tr1);
            best_update_lex<F, ANY, WANT_T1>(B, hit, t0, t1, (int)(c.first[K_MESH] + k), tr0, tr1);
        }
    }
}
template <uint32_t F, bool ANY, bool WANT_T1>
__device__ __forceinline__ void bvh_traverse(Best& B, const GlobalScene& s, const RayPre& r, const RayPk& rp) {
    // (reference, entry parameter of the pushed subtree's box) in one 8-byte word: one STL.64 per push, one LDL.64 per
    // popped entry (local memory; shared memory measured no faster)
    uint2 stack[32];
    int sp = 0;
    [[maybe_unused]] const uint32_t c_n_bvh = s.c.n_bvh;
    uint32_t cur = s.bvh_root;
    // next subtree; one that starts behind the best hit found since it was pushed holds nothing closer
    // ('<=': an equal t0 with a lower index must still be found)
    // (one step per turn: a while-while form — walk to a leaf, then test the leaves together — measured slower,
    // Instance.json 1 239 vs 1 326 Mpaths/s)
    auto pop = [&]() -> bool {
        while (sp > 0) {
            --sp;
            const uint2 e = stack[sp];
            if (ANY || __uint_as_float(e.y) <= B.t0) { cur = e.x; return true; }
        }
        return false;
    };
    for (;;) {
        if (cur & MRT_BVH_LEAF) {
            bvh_leaf<F, ANY, WANT_T1>(B, s, r, rp, cur & ~MRT_BVH_LEAF);
            if constexpr (ANY) { if (B.any) return; }
        } else {
            MRT_CHECK(cur < c_n_bvh);
            float4 q0, q1, q2;
            uint2 ref;
            load_node(&s.bvh[cur], &q0, &q1, &q2, &ref);
            float tl, tfl, tr, tfr;
            node_slabs(r.n, q0, q1, q2, &tl, &tfl, &tr, &tfr);
            // '<=': an equal t0 with a lower index must still be found
            const bool hl = tl <= tfl && tfl >= 0.0f && tl <= B.t0;
            const bool hr = tr <= tfr && tfr >= 0.0f && tr <= B.t0;
            const uint32_t cl = ref.x, cr = ref.y;
            if (hl && hr) {
                const bool left_first = tl <= tr;
                MRT_CHECK(sp < 32);
                if (sp < 32) { stack[sp] = make_uint2(left_first ? cr : cl, __float_as_uint(left_first ? tr : tl)); sp++; }
                cur = left_first ? cl : cr;
                continue;
            }
            if (hl || hr) { cur = hl ? cl : cr; continue; }
        }
        if (!pop()) return;
    }
}

// Duff's device over one kind: ENTRY(k, LE) tests entry k of the kind.
#define MRT_DUFF(n_expr, ENTRY)                                                  \
    {                                                                            \
        const uint32_t n__ = (n_expr);                                           \
        switch (n__ & 7u) {                                                      \
            case 7: ENTRY(6u, true) [[fallthrough]];                             \
            case 6: ENTRY(5u, true) [[fallthrough]];                             \
            case 5: ENTRY(4u, true) [[fallthrough]];                             \
            case 4: ENTRY(3u, true) [[fallthrough]];                             \
            case 3: ENTRY(2u, true) [[fallthrough]];                             \
            case 2: ENTRY(1u, true) [[fallthrough]];                             \
            case 1: ENTRY(0u, true) [[fallthrough]];                             \
            default: break;                                                      \
        }                                                                        \
        for (uint32_t i__ = n__ & 7u; i__ < n__; i__ += 8u) {                    \
            _Pragma("unroll") for (uint32_t u__ = 0; u__ < 8u; u__++) { ENTRY(i__ + u__, false) } \
        }                                                                        \
    }

template <class V, uint32_t F, bool ANY, bool WANT_T1>
__device__ __forceinline__ bool closest_hit(const V& sc, f3 o, f3 d, HitRec* out) {
    const SceneCommon& c = sc.c();
    RayPre r;
    r.o = o; r.d = d;
    r.m = rcp_fixed3(d);
    r.nom = mk(-o.x * r.m.x, -o.y * r.m.y, -o.z * r.m.z);
    r.am = mk(fabsf(r.m.x), fabsf(r.m.y), fabsf(r.m.z));
    r.nam = -r.am;
    r.n.bm = true_rcp3(d, r.m);  // only the BVH node tests read these
    r.n.bnom = mk(-o.x * r.n.bm.x, -o.y * r.n.bm.y, -o.z * r.n.bm.z);
    r.n.bam = mk(fabsf(r.n.bm.x), fabsf(r.n.bm.y), fabsf(r.n.bm.z));
    const RayPk rp = {pk2(o.x, d.x), pk2(o.y, d.y), pk2(o.z, d.z)};
    Best B;
    B.t0 = __int_as_float(0x7f800000); B.t1 = 0.0f; B.bi = -1; B.tr0 = B.tr1 = -1; B.any = false;

#ifdef MRT_JIT
    if constexpr (V::kJit) {
        // ascending declaration order, strict '<': the first minimum wins (rt.rs:872)
#define J_BOXP(k, a0, a1, a2, a3, a4, a5, a6, a7, a8, a9, a10, a11) \
    test_box_pair<F, ANY, WANT_T1, false, true>(B, r, BoxPair{make_float4(a0, a1, a2, a3), make_float4(a4, a5, a6, a7), make_float4(a8, a9, a10, a11)}, (int)(2 * (k)));
#define J_BOXS(k, a0, a1, a2, a3, a4, a5, a6, a7, a8, a9, a10, a11) /* the same pair, one box at a time */ \
    test_box_lit<F, ANY, WANT_T1>(B, r, a0, a2, a4, a6, a8, a10, (int)(2 * (k))); \
    test_box_lit<F, ANY, WANT_T1>(B, r, a1, a3, a5, a7, a9, a11, (int)(2 * (k) + 1));
#define J_BOX1(k, a0, a1, a2, a3, a4, a5, a6, a7, a8, a9, a10, a11) /* odd box count: lane B is padding */ \
    test_box_lit<F, ANY, WANT_T1>(B, r, a0, a2, a4, a6, a8, a10, (int)(2 * (k)));
#define J_SPH(k, cx, cy, cz, r2) \
    test_sphere<F, ANY, WANT_T1, false>(B, r, SlimInst{make_float4(cx, cy, cz, 0.0f), make_float4(r2, 0.0f, 0.0f, 0.0f)}, (int)(MRT_JIT_FIRST_SPHERE + (k)));
#define J_PLN(k, nx, ny, nz, off) \
    test_plane<F, ANY, WANT_T1, false>(B, r, SlimInst{make_float4(nx, ny, nz, 0.0f), make_float4(off, 0.0f, 0.0f, 0.0f)}, (int)(MRT_JIT_FIRST_PLANE + (k)));
#define J_BXF(k, a0, a1, a2, a3, a4, a5, a6, a7, a8, a9, a10, a11, hx, hy, hz) \
    test_bxf<F, ANY, WANT_T1, false, true>(B, rp, BxfInst{make_float4(a0, a1, a2, a3), make_float4(a4, a5, a6, a7), make_float4(a8, a9, a10, a11), make_float4(hx, hy, hz, 0.0f)}, (int)(MRT_JIT_FIRST_BXF + (k)), r.m.z);
#define J_MSH(k, px, py, pz, rot, mid, m0, m1, m2, m3, m4, m5, m6, m7, m8, m9, m10, m11) \
    if constexpr ((F & F_MESH) != 0) { const Xf x__ = {{m0, m1, m2, m3, m4, m5, m6, m7, m8, m9, m10, m11}}; \
        test_mesh<F, ANY, WANT_T1, false>(B, c, r, SlimInst{make_float4(px, py, pz, 0.0f), make_float4(__uint_as_float(rot), __uint_as_float(mid), 0.0f, 0.0f)}, x__, (int)(MRT_JIT_FIRST_MESH + (k))); }
        // J_CB / J_CE bracket a cluster of consecutive boxes with its bounding box (big scenes only): the
        // cluster is skipped when the ray misses the box or enters it behind the best hit so far
#define J_CB(lx, ly, lz, hx, hy, hz) { if (bracket_hit(r, make_float4(lx, ly, lz, 0.0f), make_float4(hx, hy, hz, 0.0f), ANY ? __int_as_float(0x7f800000) : B.t0)) {
#define J_CE }}
        MRT_JIT_BOXPAIRS(J_BOXP, J_BOXS, J_BOX1, J_CB, J_CE)
#undef J_CB
#undef J_CE
        MRT_JIT_SPHERES(J_SPH)
        MRT_JIT_PLANES(J_PLN)
        MRT_JIT_BXFS(J_BXF)
        MRT_JIT_MESHES(J_MSH)
#undef J_BOXP
#undef J_BOXS
#undef J_BOX1
#undef J_SPH
#undef J_PLN
#undef J_BXF
#undef J_MSH
        if constexpr (ANY) return B.any;
        out->t0 = B.t0; out->t1 = B.t1; out->inst = B.bi; out->tri0 = B.tr0; out->tri1 = B.tr1;
        return B.bi >= 0;
    }
#endif
#define E_BOX(k, LE) test_box_pair<F, ANY, WANT_T1, LE>(B, r, sc.boxp(k), (int)(2u * (k)));
#define E_SPH(k, LE) test_sphere<F, ANY, WANT_T1, LE>(B, r, sc.sph(k), (int)(c.first[K_SPHERE] + (k)));
#define E_PLN(k, LE) test_plane<F, ANY, WANT_T1, LE>(B, r, sc.pln(k), (int)(c.first[K_PLANE] + (k)));
#define E_BXF(k, LE) test_bxf<F, ANY, WANT_T1, LE>(B, rp, sc.bxf(k), (int)(c.first[K_BOX_XF] + (k)));
#define E_MSH(k, LE) test_mesh<F, ANY, WANT_T1, LE>(B, c, r, sc.mesh(k), sc.mesh_m(k), (int)(c.first[K_MESH] + (k)));
    if constexpr (V::kBvh) {
        if (MRT_BVH_ALWAYS || sc.s.bvh != nullptr) {  // warp-uniform: large scenes only
            bvh_traverse<F, ANY, WANT_T1>(B, sc.s, r, rp);
            for (uint32_t k = 0; k < MRT_N_PLANES(c); k++) {  // planes are infinite: brute force, same tie rule
                const SlimInst e = sc.pln(k);
                const float t0 = (e.b.x - dot(r.o, xyz(e.a))) * frcp(dot(r.d, xyz(e.a)));
                best_update_lex<F, ANY, WANT_T1>(B, t0 > 0.0f, t0, t0, (int)(c.first[K_PLANE] + k), -1, -1);
            }
            if constexpr (ANY) return B.any;
            out->t0 = B.t0; out->t1 = B.t1; out->inst = B.bi; out->tri0 = B.tr0; out->tri1 = B.tr1;
            return B.bi >= 0;
        }
    }
    if constexpr (!V::kJit && !(V::kBvh && MRT_BVH_ALWAYS)) {
        MRT_DUFF((c.cnt[K_BOX] + 1u) >> 1, E_BOX)
        MRT_DUFF(c.cnt[K_SPHERE], E_SPH)
        MRT_DUFF(c.cnt[K_PLANE], E_PLN)
        for (uint32_t k = 0; k < c.cnt[K_BOX_XF]; k++) { E_BXF(k, false) }
        if constexpr ((F & F_MESH) != 0) {
            for (uint32_t k = 0; k < c.cnt[K_MESH]; k++) { E_MSH(k, false) }
        }
    }
#undef E_BOX
#undef E_SPH
#undef E_PLN
#undef E_BXF
#undef E_MSH
    if constexpr (ANY) return B.any;
    out->t0 = B.t0; out->t1 = B.t1; out->inst = B.bi; out->tri0 = B.tr0; out->tri1 = B.tr1;
    return B.bi >= 0;
}

// ------------------------------------------------------------------ per-hit data of the winner
// What the specialised kernel knows about instance rotations (MRT_JIT_ROT): 0 = no instance is
// rotated, 1 = every rotated instance is yaw-only (M = [[a,b,0],[c,d,0],[0,0,1]]), 2 = anything.
#if defined(MRT_JIT) && defined(MRT_JIT_ROT)
#define MRT_ROT MRT_JIT_ROT
#else
#define MRT_ROT 2
#endif
struct Surf {
    float4 P, A, m0, m1, m2;
    __device__ __forceinline__ uint32_t flags() const { return __float_as_uint(P.w); }
    __device__ __forceinline__ uint32_t kind() const { return flags() & 0xffu; }
    __device__ __forceinline__ bool identity() const { return MRT_ROT == 0 || (flags() & FAT_IDENT) != 0u; }
    __device__ __forceinline__ bool textured() const { return (flags() & FAT_TEX) != 0u; }
    __device__ __forceinline__ bool normal_xf() const { return MRT_ROT != 0 && (flags() & FAT_NXF) != 0u; }
};
// M v with what MRT_ROT allows to skip
__device__ __forceinline__ f3 mulM_rot(const Surf& s, f3 v) {
#if MRT_ROT == 1
    return {fmaf(s.m0.y, v.y, s.m0.x * v.x), fmaf(s.m1.y, v.y, s.m1.x * v.x), v.z};
#else
    return {dot(xyz(s.m0), v), dot(xyz(s.m1), v), dot(xyz(s.m2), v)};
#endif
}
__device__ __forceinline__ void load_surf(const FatInst* f, Surf* s) {
    const f8 pa = ldg256(f);  // P, A
    s->P = pa.lo;
    s->A = pa.hi;
    // rows only when rotated or textured (the texture ids live in .w)
    const bool rows = (MRT_ROT == 0) ? (s->flags() & FAT_TEX) != 0u : (s->flags() & (FAT_IDENT | FAT_TEX)) != FAT_IDENT;
    if (rows) {
        const f8 m01 = ldg256(&f->m0);  // m0, m1
        s->m0 = m01.lo;
        s->m1 = m01.hi;
        s->m2 = __ldg(&f->m2);
    }
#ifndef MRT_JIT
    else {  // never read (every use is behind !identity() / textured()); initialised in the offline build
            // because ptxas allocates 8 registers fewer with it (64 vs 72: one more resident block per SM)
        s->m0 = make_float4(1.f, 0.f, 0.f, __uint_as_float(0xffffffffu));
        s->m1 = make_float4(0.f, 1.f, 0.f, __uint_as_float(0xffffffffu));
        s->m2 = make_float4(0.f, 0.f, 1.f, __uint_as_float(0xffffffffu));
    }
#endif
}
// object-space hit point minus instance pos: rot_y * (look * (hp - pos)), rt.rs:782
__device__ __forceinline__ f3 to_local(const Surf& s, f3 hp) {
    f3 r = hp - xyz(s.P);
    return s.identity() ? r : mulM_rot(s, r);
}
// The winner of a closest-hit search is a sphere: its entry / exit parameters once more, in the REFERENCE'S OWN
// ARITHMETIC — Renderer::intersect's ray (rt.rs:729-733: pos + M (o - pos)) and Sphere::intersect (rt.rs:335-359:
// o = orig - pos, a = d.d, b = 2 o.d, c = o.o - r^2, disc = b^2 - 4ac, (-b -+ sqrt(disc)) / 2a), every operation rounded
// on its own.  The search uses the cheaper half-b form with fused multiply-adds.  Both
// are ill-conditioned for a small sphere seen from afar (o.o - r^2 cancels), but their rounding noise differs — and
// that noise decides whether the NEXT ray, which starts E = 1e-4 off the computed hit point, begins inside the sphere
// (and leaves it unseen, rt.rs:353) or outside (and may hit it again: one more bounce, one more direct-light term).
// On Instance.json (1000 spheres of r = 0.2 seen from 5 - 9 units away) the search's own t made the image 0.85 % darker
// than the oracle's, 5 sigma at 64 spp (tests/test_gpu_statistics.py); an oracle with the half-b form shows the same
// shift.  Costs ~40 instructions per SPHERE HIT (not per test): nothing measurable.
__device__ __forceinline__ float dot_rn(f3 a, f3 b) {  // lin.rs:259-264: x*x + y*y + z*z, left to right, no contraction
    return __fadd_rn(__fadd_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)), __fmul_rn(a.z, b.z));
}
// Only scenes where it can matter pay for it (SceneCommon::refine_spheres, decided by mrt_set_scene): the noise of t is
// ~ ulp(|o - pos|^2) / (2 sqrt(disc)) ~ 1.2e-7 D^2 / r for rays of length D, and the next ray starts E = 1e-4 off the hit
// point, so where 1.2e-7 D^2 / r stays below 4e-5 for the scene's extent D (mrt_scene.cu) neither form's noise flips the
// side of the surface a ray starts on often enough to show (CornellBox2: 1.2e-5; Instance.json: 1.7e-4).
#if defined(MRT_JIT)
#define MRT_REFINE_SPHERES(c) (MRT_JIT_REFINE_SPHERES != 0)
#else
#define MRT_REFINE_SPHERES(c) ((c).refine_spheres != 0u)
#endif
__device__ __forceinline__ void refine_sphere_hit(const SceneCommon& c, f3 o, f3 d, HitRec* h) {
    const FatInst* f = c.fat + h->inst;
    const float4 P = __ldg(&f->P);
    const uint32_t flags = __float_as_uint(P.w);
    if ((flags & 0xffu) != K_SPHERE) return;
    const f3 pos = xyz(P);
    const float r = __ldg(&f->A).y;
    f3 rel = mk(__fadd_rn(o.x, -pos.x), __fadd_rn(o.y, -pos.y), __fadd_rn(o.z, -pos.z));
    f3 dl = d;
    if (MRT_ROT != 0 && (flags & FAT_IDENT) == 0u) {
        const float4 m0 = __ldg(&f->m0), m1 = __ldg(&f->m1), m2 = __ldg(&f->m2);
        rel = mk(dot_rn(xyz(m0), rel), dot_rn(xyz(m1), rel), dot_rn(xyz(m2), rel));
        dl = mk(dot_rn(xyz(m0), d), dot_rn(xyz(m1), d), dot_rn(xyz(m2), d));
    }
    const f3 no = mk(__fadd_rn(pos.x, rel.x), __fadd_rn(pos.y, rel.y), __fadd_rn(pos.z, rel.z));     // n_ray.orig
    const f3 oo = mk(__fadd_rn(no.x, -pos.x), __fadd_rn(no.y, -pos.y), __fadd_rn(no.z, -pos.z));     // Sphere::intersect's o
    const float a = dot_rn(dl, dl);
    const float b = __fmul_rn(2.0f, dot_rn(oo, dl));
    const float cc = __fadd_rn(dot_rn(oo, oo), -__fmul_rn(r, r));
    const float disc = __fadd_rn(__fmul_rn(b, b), -__fmul_rn(__fmul_rn(4.0f, a), cc));
    if (!(disc >= 0.0f)) return;  // the search saw a hit the literal form does not: keep the search's numbers
    // (the square root and the division go through the fast units: their 1 - 2 ulp are nothing against the rounding of
    // b^2 - 4ac, whose cancellation is what shapes the noise; IEEE versions cost 200 instructions and 8 registers)
    const float sq = sqrtf(disc), inv = frcp(__fmul_rn(2.0f, a));
    const float t0 = __fmul_rn(__fadd_rn(-b, -sq), inv), t1 = __fmul_rn(__fadd_rn(-b, sq), inv);
    if (!(t0 >= 0.0f)) return;
    h->t0 = t0; h->t1 = t1;
}

// Box::normal, rt.rs:414-445, on p = local point * 2/size: windows |p_i| in 1 +- E, checked
// x, -x, y, -y, then — the missing `else` at :435 — an independent z test that overrides.
// (The reference's windows are half-open, [1-E, 1+E); the open form used here differs only when
// |p_i| equals a window end exactly.)  When no window matches the reference normalises a zero
// vector (NaN normal, measured 1e-7 of hits); the nearest face is taken instead.
__device__ __forceinline__ f3 box_face(f3 p) {
    // windowed axes get a score that encodes the reference's priority (z over x over y); the rest
    // score e = |p_i| - 1 (<= 0 on the surface), so the largest score also picks the nearest face
    // when no window matches.  Branch-free.  (Testing the three windows directly and keeping the nearest-face rule
    // for the ~1e-7 of hits without one is fewer operations on paper, but the compiler turns it into a chain of
    // divergent branches: 331 instead of 317 instructions per loop iteration of the headline kernel.)
    const float ex = fabsf(p.x) - 1.0f, ey = fabsf(p.y) - 1.0f, ez = fabsf(p.z) - 1.0f;
    const float sz = fabsf(ez) < MRT_E ? 3e30f : ez;
    const float sx = fabsf(ex) < MRT_E ? 2e30f : ex;
    const float sy = fabsf(ey) < MRT_E ? 1e30f : ey;
    const bool fz = sz >= sx && sz >= sy;
    const bool fx = !fz && sx >= sy;
    const bool fy = !fz && !fx;
    return mk(fx ? copysignf(1.0f, p.x) : 0.0f, fy ? copysignf(1.0f, p.y) : 0.0f, fz ? copysignf(1.0f, p.z) : 0.0f);
}
// Renderer::normal, rt.rs:776-793: kind normal of the object-space hit point, pushed through
// the FORWARD transform again (rt.rs:792; unless MRT_NORMAL_OBJECT is selected) and normalised.  Unit inputs through an orthonormal M
// stay unit to 1e-7, so only mesh normals need the rsqrt.
// The specialised kernel knows which kinds the scene holds (MRT_JIT_HAS_*): absent kinds cost nothing.
#ifdef MRT_JIT
#define MRT_HAS_PLANE (MRT_JIT_N_PLANE > 0)
#define MRT_HAS_SPHERE (MRT_JIT_N_SPHERE > 0)
#define MRT_HAS_BOX (MRT_JIT_N_BOX > 0)
#else
#define MRT_HAS_PLANE true
#define MRT_HAS_SPHERE true
#define MRT_HAS_BOX true
#endif
template <uint32_t F>
__device__ __forceinline__ f3 surf_normal(const SceneCommon& c, const Surf& s, f3 pl, int tri) {
    const uint32_t k = s.kind();
    if (MRT_HAS_PLANE && k == K_PLANE) return xyz(s.A);  // precomputed norm(M n)
    f3 n;
    bool mesh = false;
    if constexpr ((F & F_MESH) != 0) mesh = k == K_MESH;
    if (MRT_HAS_SPHERE && (k == K_SPHERE || (!MRT_HAS_BOX && !mesh))) n = pl * s.A.x;  // (hit - pos) / r
    else if (mesh) {
        MRT_CHECK(tri >= 0 && __float_as_uint(s.A.x) + (uint32_t)tri < c.n_tri);
        const DTri* tp = &c.tri[__float_as_uint(s.A.x) + (uint32_t)tri];
        n = cross(xyz(__ldg(&tp->e0)), xyz(__ldg(&tp->e1)));  // rt.rs:459-466
    } else n = box_face(pl * xyz(s.A));
    if (s.normal_xf()) n = mulM_rot(s, n);  // MRT_OPT_NORMAL_SPACE
    if (mesh) n = normalize(n);
    return n;
}

// UV impls, rt.rs:468-548, on the object-space point (pl = point - pos; plane uses the absolute
// object-space point, rt.rs:529-541).
__device__ __forceinline__ float2 surf_uv(const Surf& s, f3 pl) {
    const uint32_t k = s.kind();
    if (k == K_SPHERE) {
        f3 v = normalize(pl);
        return make_float2(0.5f + 0.5f * atan2f(v.x, -v.y) * 0.31830988618379067154f, 0.5f - 0.5f * v.z);
    }
    if (k == K_PLANE) {
        f3 h = pl + xyz(s.P);
        float x = h.x + 0.5f; x = x - truncf(x); if (x < 0.0f) x = 1.0f + x;
        float y = h.y + 0.5f; y = y - truncf(y); if (y < 0.0f) y = 1.0f + y;
        return make_float2(x, y);
    }
    if (k == K_BOX || k == K_BOX_XF) {
        f3 p = pl * xyz(s.A);
        const float pl_ = 1.0f - MRT_E, ph = 1.0f + MRT_E, nl = -1.0f - MRT_E, nh = -1.0f + MRT_E;
        const float th = 1.0f / 3.0f;
        // same chain and priorities as rt.rs:475-514 (x and y faces return before z is looked at)
        if (p.x >= pl_ && p.x < ph) return make_float2((0.5f + 0.5f * p.y) * 0.25f + 0.5f, (0.5f - 0.5f * p.z) * th + th);
        if (p.x >= nl && p.x < nh) return make_float2((0.5f - 0.5f * p.y) * 0.25f, (0.5f - 0.5f * p.z) * th + th);
        if (p.y >= pl_ && p.y < ph) return make_float2((0.5f - 0.5f * p.x) * 0.25f + 0.75f, (0.5f - 0.5f * p.z) * th + th);
        if (p.y >= nl && p.y < nh) return make_float2((0.5f + 0.5f * p.x) * 0.25f + 0.25f, (0.5f - 0.5f * p.z) * th + th);
        if (p.z >= pl_ && p.z < ph) return make_float2((0.5f + 0.5f * p.x) * 0.25f + 0.25f, (0.5f - 0.5f * p.y) * th);
        if (p.z >= nl && p.z < nh) return make_float2((0.5f + 0.5f * p.x) * 0.25f + 0.25f, (0.5f + 0.5f * p.y) * th + 2.0f * th);
        return make_float2(0.0f, 0.0f);
    }
    return make_float2(0.0f, 0.0f);
}

// Texture::get_color, rt.rs:618-628: nearest texel, truncating casts, linear index x + y*w
// (u == 1 runs into the next row, as in the reference); the index is clamped to the last
// texel where the reference would panic.
__device__ __forceinline__ float4 tex_fetch(const SceneCommon& c, uint32_t id, float2 uv) {
    MRT_CHECK(id < c.n_tex);
    const DTex t = c.tex[id];
    if (!t.has_dat) return make_float4(0.f, 0.f, 0.f, 0.f);
    float fx = uv.x * (float)t.w, fy = uv.y * (float)t.h;
    // Rust `as usize`: NaN/negative -> 0, saturating
    unsigned long long x = (fx > 0.0f) ? (fx >= 1.8446744e19f ? ~0ull : (unsigned long long)fx) : 0ull;
    unsigned long long y = (fy > 0.0f) ? (fy >= 1.8446744e19f ? ~0ull : (unsigned long long)fy) : 0ull;
    const unsigned long long n = (unsigned long long)t.w * t.h;
    unsigned long long idx = (y >= n || x >= n) ? n - 1 : y * t.w + x;
    if (idx >= n) idx = n - 1;
    MRT_CHECK(t.first + (uint32_t)idx < c.n_texels);
    return __ldg(&c.texels[t.first + (uint32_t)idx]);
}

struct Mat {
    f3 color;
    float rough, metal, glass, opacity, emit;
    float metal_raw;  // Ray::reflect tests the raw field (rt.rs:564), not the mmap
};
// material getters, rt.rs:811-863, all maps fetched at the same uv
template <uint32_t F>
__device__ __forceinline__ void load_mat(const SceneCommon& c, const FatInst* f, const Surf& s, f3 pl, Mat* m) {
#if defined(MRT_JIT) && defined(MRT_JIT_UNIFORM_R)
    const float4 ae = __ldg(&f->C);
    const float4 r = make_float4(MRT_JIT_UNIFORM_R);  // every material of the scene has these rough/metal/glass/opacity: literals
#else
    const f8 cr = ldg256(&f->C);  // C, R
    const float4 ae = cr.lo, r = cr.hi;
#endif
    m->color = xyz(ae); m->emit = ae.w;
    m->rough = r.x; m->metal = r.y; m->glass = r.z; m->opacity = r.w; m->metal_raw = r.y;
    if constexpr ((F & F_TEX) != 0) {
        if (s.textured()) {
            const uint32_t t0 = __float_as_uint(s.m0.w), t1 = __float_as_uint(s.m1.w), t2 = __float_as_uint(s.m2.w);
            const float2 uv = surf_uv(s, pl);
            if ((t0 & 0xffffu) != 0xffffu) m->color = m->color * xyz(tex_fetch(c, t0 & 0xffffu, uv));
            if ((t0 >> 16) != 0xffffu) m->rough = tex_fetch(c, t0 >> 16, uv).x;
            if ((t1 & 0xffffu) != 0xffffu) m->metal = tex_fetch(c, t1 & 0xffffu, uv).x;
            if ((t1 >> 16) != 0xffffu) m->glass = tex_fetch(c, t1 >> 16, uv).x;
            if ((t2 & 0xffffu) != 0xffffu) m->opacity = tex_fetch(c, t2 & 0xffffu, uv).x;
            if ((t2 >> 16) != 0xffffu) m->emit = tex_fetch(c, t2 >> 16, uv).x;
        }
    }
}

// RayTracer::rand, rt.rs:996-1007: n + r * (uniform point on the unit sphere), normalised.
// th = acos(1-2u1) => cos th = 1-2u1, sin th = sqrt(1 - cos^2): same distribution, no acos.
__device__ __forceinline__ f3 rand_normal(f3 n, float r, float u1, float u2) {
    float z = 1.0f - 2.0f * u1;
    float st = sqrtf(fmaxf(0.0f, fmaf(-z, z, 1.0f)));
    float sp, cp;
#ifdef MRT_PRECISE
    sincosf(u2 * 6.283185307179586f, &sp, &cp);
#else
    __sincosf(u2 * 6.283185307179586f, &sp, &cp);
#endif
    f3 v = mk(st * cp, st * sp, z);
    return normalize(fma3(v, r, n));
}
// the same from the raw words: the conversions fold into the first multiply of each use
__device__ __forceinline__ f3 rand_normal_w(f3 n, float r, uint32_t w1, uint32_t w2) {
    const float z = fmaf((float)w1, -2.0f * MRT_U32_TO_UNIT, 1.0f);
    const float st = sqrtf(fmaxf(0.0f, fmaf(-z, z, 1.0f)));
    float sp, cp;
#ifdef MRT_PRECISE
    sincosf((float)w2 * (6.283185307179586f * MRT_U32_TO_UNIT), &sp, &cp);
#else
    __sincosf((float)w2 * (6.283185307179586f * MRT_U32_TO_UNIT), &sp, &cp);
#endif
    return normalize(fma3(mk(st * cp, st * sp, z), r, n));
}
__device__ __forceinline__ f3 reflect3(f3 v, f3 n) { return fma3(n, -2.0f * dot(v, n), v); }  // lin.rs:68-70
