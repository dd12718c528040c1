// mrt_path.cuh — the body of the path-tracing megakernel (camera ray generation + bounce loop),
// shared by the offline-compiled generic kernels (mrt_kernels.cu) and the scene-specialised
// kernel that mrt_jit.cu compiles at run time with NVRTC.  Reference: rt.rs:900-1066.
#pragma once
#include "mrt_device.cuh"

// RayTracer::iter + cast, rt.rs:900-954, split into the per-pixel part (q = focus point
// minus camera position) and the per-sample lens jitter.
__device__ __forceinline__ f3 pixel_focus_vec(const FilmParams& fp, uint32_t px, uint32_t py) {
    const float w = fp.fw, h = fp.fh;
    const float aspect = w / h;
    const float ux = aspect * ((float)px - 0.5f * w) / w;
    const float uy = ((float)py - 0.5f * h) / h;
    const f3 dir = normalize(mk(ux, fp.fy, -uy));
    // p = (cam.pos + dir*E) + dir*foc ; q = p - cam.pos
    return fma3(dir, fp.foc, dir * MRT_E);
}
__device__ __forceinline__ void camera_ray(const FilmParams& fp, f3 q, float u1, float u2, f3* o, f3* d) {
    const float jx = (u1 - 0.5f) * fp.aprt, jz = (u2 - 0.5f) * fp.aprt;
    const f3 nd = normalize(mk(q.x - jx, q.y, q.z - jz));
    const f3 dd = fp.cam_identity ? nd : mk(fmaf(fp.cam_M[2], nd.z, fmaf(fp.cam_M[1], nd.y, fp.cam_M[0] * nd.x)),
                     fmaf(fp.cam_M[5], nd.z, fmaf(fp.cam_M[4], nd.y, fp.cam_M[3] * nd.x)),
                     fmaf(fp.cam_M[8], nd.z, fmaf(fp.cam_M[7], nd.y, fp.cam_M[6] * nd.x)));
    const f3 pos = mk(fp.cam_pos[0] + jx, fp.cam_pos[1], fp.cam_pos[2] + jz);
    *o = fma3(dd, MRT_E, pos);  // Ray::cast_default, rt.rs:555-557
    *d = dd;
}

#ifndef MRT_POOL_BLOCK
#define MRT_POOL_BLOCK 128  // threads per block of the pooled kernels (= MRT_PATH_BLOCK)
#endif
struct PathState {
    f3 o, d;          // current ray
    f3 T;             // throughput of the path so far: prod (0.5 + color_i) pwr_i   (rt.rs:990-992, forward form)
    float pwr;        // ray.pwr: (1 - loss)^bounce   (rt.rs:571)
    uint32_t bounce;  // MRT_NEED_PATH: the lane needs a new camera path
};
#define MRT_NEED_PATH 0xffffffffu

// unit vector from a hit point towards light li: norm(pos - hit) or -norm(dir) (rt.rs:1029-1032, 975-978)
__device__ __forceinline__ f3 light_vec(const SceneCommon& c, uint32_t li, f3 hp) {
    const float4 lv = c.light[li].v_kind;
    return (__float_as_uint(lv.w) == 0u) ? normalize(xyz(lv) - hp) : xyz(lv);
}
// a path that left the scene: primary miss = sky.color (rt.rs:958); later miss = the fold's seed sky.color * sky.pwr (rt.rs:964)
__device__ __forceinline__ void path_miss(const SceneCommon& c, const PathState& p, f3& acc) {
#if !(defined(MRT_JIT) && defined(MRT_JIT_SKY_BLACK))  // a black sky adds nothing: compiled out when the scene says so
    if (p.bounce == 0) acc = acc + mk(c.sky[0], c.sky[1], c.sky[2]);
    else acc = acc + p.T * mk(c.sky_tail[0], c.sky_tail[1], c.sky_tail[2]);
#else
    (void)c; (void)p; (void)acc;
#endif
}

// Everything of a path segment that follows the searches: the hit `h` of the segment's ray and the lights `vis` that
// are visible from its entry point are known.  Ray::reflect / refract (rt.rs:559-589), the transmission lottery
// (rt.rs:1051-1059) and the step of RayTracer::reduce_light (rt.rs:956-994, evaluated forward).  Adds the segment's
// radiance to `acc` and advances `p` to the next ray; returns true when the path ended here (emission, bounce limit).
template <uint32_t F>
__device__ __forceinline__ bool shade_hit(const SceneCommon& c, const FilmParams& fp, uint32_t pix, uint32_t sample, PathState& p, f3& acc,
                                          const HitRec& h, uint32_t vis) {
    const f3 o = p.o, d = p.d;
    const FatInst* fat = c.fat + h.inst;
    Surf s;
    load_surf(fat, &s);
    f3 hp = fma3(d, h.t0, o);
    f3 pl = to_local(s, hp);
    f3 n = surf_normal<F>(c, s, pl, h.tri0);
    Mat m;
    load_mat<F>(c, fat, s, pl, &m);

    const uint4 uw = rng_words(pix, sample, 1u + 2u * p.bounce, fp.key);  // x: reflect lottery, y z: direction, w: emission draw

    // Ray::reflect from the entry hit, rt.rs:559-572
    f3 nd, no;
    {
        float rough = m.rough;
        if (m.metal_raw == 0.0f && m.opacity != 0.0f && uw.x < MRT_LOTTERY_80) rough = 1.0f;  // u < 0.8
        const f3 nn = rand_normal_w(n, rough, uw.y, uw.z);
        nd = reflect3(d, nn);  // unit d about unit nn stays unit (the reference's .norm() is a no-op to 1e-7)
        no = fma3(nd, MRT_E, hp);
    }
    // 15 % chance to keep reflecting for transparent material, rt.rs:1051-1059
    if constexpr ((F & F_TRANSMIT) != 0) {
        const float pt = fminf(1.0f - m.opacity, 0.85f);
        if (pt > 0.0f) {
            const float4 ub = rng_block(pix, sample, 2u + 2u * p.bounce, fp.key);
            if (ub.x < pt) {
                // exit hit: point, normal and material at t1 (rt.rs:886-894)
                const f3 hp1 = fma3(d, h.t1, o);
                const f3 pl1 = to_local(s, hp1);
                const f3 n1 = surf_normal<F>(c, s, pl1, h.tri1);
                Mat m1;
                load_mat<F>(c, fat, s, pl1, &m1);
                // Ray::refract, rt.rs:574-589 ; Vec3f::refract, lin.rs:96-105
                float rough = m1.rough;
                if (m1.metal_raw == 0.0f && m1.opacity != 0.0f && ub.y < 0.80f) rough = 1.0f;
                const f3 nn = rand_normal(n1, rough, ub.z, ub.w);
                const float eta = 1.0f + 0.5f * m1.glass;
                const float cs = -dot(nn, d);
                const float k = 1.0f - eta * eta * (1.0f - cs * cs);
                if (k >= 0.0f) {
                    const f3 rd = normalize(fma3(nn, cs * eta + sqrtf(k), d * eta));
                    nd = rd;
                    no = fma3(rd, MRT_E, hp1);
                    hp = hp1; n = n1; m = m1;  // the recorded hit is the exit hit
                }
            }
        }
    }

    // ---- RayTracer::reduce_light, rt.rs:956-994, evaluated forward
#if defined(MRT_JIT) && defined(MRT_JIT_EMIT_BINARY)
    // every emit of the scene is 0 or 1 (no emap): gen_bool(1) is always true, gen_bool(0) never — no draw needed
    if (m.emit != 0.0f) {
#else
    if ((float)uw.w * MRT_U32_TO_UNIT < m.emit) {  // emission draw, rt.rs:966-970
#endif
        acc = acc + p.T * m.color;
        return true;
    }
    if constexpr ((F & F_LIGHTS) != 0) {
        f3 lc = mk(0.f, 0.f, 0.f);
        for (uint32_t li = 0; li < MRT_N_LIGHTS(c); li++) {
            if (!((vis >> li) & 1u)) continue;
            const float4 cp = c.light[li].color_pwr;
            const f3 l = light_vec(c, li, hp);
            const float diff = fmaxf(dot(l, n), 0.0f);
            float sp = fmaxf(dot(d, reflect3(l, n)), 0.0f);
            sp *= sp; sp *= sp; sp *= sp; sp *= sp; sp *= sp;  // powi(32), rt.rs:981
            sp *= (1.0f - m.rough);
            const f3 oc = m.color * ((1.0f - m.metal) * diff);
            lc = lc + mk(fmaf(oc.x, cp.x, sp), fmaf(oc.y, cp.y, sp), fmaf(oc.z, cp.z, sp)) * cp.w;
        }
        acc = acc + p.T * (lc * p.pwr);
    }
    p.T = p.T * mk((0.5f + m.color.x) * p.pwr, (0.5f + m.color.y) * p.pwr, (0.5f + m.color.z) * p.pwr);  // rt.rs:990-992
    p.o = no; p.d = nd;
    p.pwr *= fp.keep;
    p.bounce++;
    // a NaN direction (|n + r v| = 0) would poison the lane: end the path as a miss
    const bool bad = !(fabsf(nd.x) <= 2.0f);
    if (p.bounce > fp.max_bounce || bad) {  // rt.rs:1018
#if !(defined(MRT_JIT) && defined(MRT_JIT_SKY_BLACK))
        acc = acc + p.T * mk(c.sky_tail[0], c.sky_tail[1], c.sky_tail[2]);
#endif
        return true;
    }
    return false;
}

// One path segment: RaytraceIterator::next (rt.rs:1014-1066) — closest hit of the ray, light visibility from the entry
// point (shadow rays without a distance limit, rt.rs:1027-1045) — then shade_hit.  Returns true when the path ended.
template <class V, uint32_t F>
__device__ __forceinline__ bool path_segment(const V& sc, const FilmParams& fp, uint32_t pix, uint32_t sample, PathState& p, f3& acc) {
    const SceneCommon& c = sc.c();
    HitRec h;
    if (!closest_hit<V, F, false, (F & F_TRANSMIT) != 0>(sc, p.o, p.d, &h)) {
        path_miss(c, p, acc);
        return true;
    }
    uint32_t vis = 0;
    if constexpr ((F & F_LIGHTS) != 0) {
        const f3 hp = fma3(p.d, h.t0, p.o);
        for (uint32_t li = 0; li < MRT_N_LIGHTS(c); li++) {
            const f3 l = light_vec(c, li, hp);
            HitRec dummy;
            if (!closest_hit<V, F, true, false>(sc, fma3(l, MRT_E, hp), l, &dummy)) vis |= 1u << li;
        }
    }
    return shade_hit<F>(c, fp, pix, sample, p, acc, h, vis);
}

// Pixel of this thread.  Tiled: a warp renders an 8x4 tile and a block 16x8, so the lanes of a warp look at
// neighbouring surfaces (paths of similar length, similar BVH traversals); the RNG and the accumulator are
// keyed by the PIXEL, so the image does not depend on the mapping.
__device__ __forceinline__ void thread_pixel(const FilmParams& fp, uint32_t* px, uint32_t* py) {
    if (fp.tiles_x) {
        const uint32_t by = blockIdx.x / fp.tiles_x, bx = blockIdx.x - by * fp.tiles_x;
        const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
        *px = bx * 16u + (w & 1u) * 8u + (lane & 7u);
        *py = by * 8u + (w >> 1) * 4u + (lane >> 3);
    } else {
        const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
        *py = i / fp.nw; *px = i - *py * fp.nw;
    }
}

// The megakernel body, one thread per supersampled pixel: the lane renders its pixel's samples back to back.
template <class V, uint32_t F>
__device__ __forceinline__ void path_body(const V sc, const FilmParams& fp) {
    uint32_t px, py;
    thread_pixel(fp, &px, &py);
    if (px >= fp.nw || py >= fp.nh || fp.n_samples == 0u) return;
    const uint32_t pix = py * fp.nw + px;
    const f3 q = pixel_focus_vec(fp, px, py);
    const uint32_t cam_seed = cam_hash_seed(pix, fp.key);

    f3 acc = mk(0.f, 0.f, 0.f);
    PathState p;
    p.o = mk(0.f, 0.f, 0.f); p.d = mk(0.f, 1.f, 0.f); p.T = mk(1.f, 1.f, 1.f);
    p.pwr = 1.0f;
    p.bounce = MRT_NEED_PATH;
    uint32_t j = 0;

    // One iteration = one path segment.  A lane whose path ended (miss, emission, bounce limit)
    // starts its next camera sample at the top of the next iteration and rejoins the warp before
    // the closest-hit search, so a warp only idles once a lane has rendered all its samples.
    // Measured alternatives (profiles/): regeneration at the loop tail makes the compiler peel it
    // into an outer loop (lanes then wait for each other's paths to end); folding the rare miss /
    // bounce-limit exits into predicated straight-line code costs more issue slots than the
    // divergent blocks it removes.
    for (;;) {
        if (p.bounce == MRT_NEED_PATH) {
            if (j >= fp.n_samples) break;
            const float2 u = rng_cam(cam_seed, fp.sample0 + j * fp.sample_stride);
            camera_ray(fp, q, u.x, u.y, &p.o, &p.d);
            p.T = mk(1.f, 1.f, 1.f);
            p.pwr = 1.0f;
            p.bounce = 0;
        }
        const uint32_t sample = fp.sample0 + j * fp.sample_stride;
        if (path_segment<V, F>(sc, fp, pix, sample, p, acc)) { p.bounce = MRT_NEED_PATH; j++; }
    }
    float4 a = fp.accum[pix];
    a.x += acc.x; a.y += acc.y; a.z += acc.z;
    fp.accum[pix] = a;
}

// The same kernel with the lanes UNBOUND from the pixels: a warp owns the (pixel, sample) items of its tile — 32
// pixels x n_samples — and a lane that finishes a path takes the next item from the warp's counter, whichever pixel
// it belongs to.  In scenes searched through a BVH the cost of a path varies by orders of magnitude between the
// pixels of a tile (sky next to a mesh silhouette): bound to its pixel a lane on a cheap pixel runs out of samples
// and idles for the rest of the launch (ncu, round 1: 5.8 - 10.9 of 32 lanes active); unbound, every lane works until
// the tile's pool is dry.  The RNG is keyed by (pixel, global sample), so the paths are the same paths; only the
// order in which a pixel's samples are summed changes.  Per warp in shared memory: the pixels' focus vectors and lens
// seeds, their radiance sums (float atomics: at most a handful of lanes finish in the same iteration), the counter.
template <class V, uint32_t F>
__device__ __forceinline__ void path_body_pool(const V sc, const FilmParams& fp) {
    __shared__ float4 s_q[MRT_POOL_BLOCK];        // focus vector of the pixel (xyz) + its film index (w, bits)
    __shared__ float s_acc[3][MRT_POOL_BLOCK];
    __shared__ uint32_t s_next[MRT_POOL_BLOCK / 32];
    uint32_t px, py;
    thread_pixel(fp, &px, &py);
    const bool inside = px < fp.nw && py < fp.nh;
    const uint32_t lane = threadIdx.x & 31u, wbase = threadIdx.x & ~31u, w = threadIdx.x >> 5;
    const uint32_t valid = __ballot_sync(0xffffffffu, inside);
    if (valid == 0u || fp.n_samples == 0u) return;  // warp-uniform
    const uint32_t pix = py * fp.nw + px;
    {
        const f3 q = inside ? pixel_focus_vec(fp, px, py) : mk(0.f, 1.f, 0.f);
        s_q[threadIdx.x] = make_float4(q.x, q.y, q.z, __uint_as_float(pix));
        s_acc[0][threadIdx.x] = 0.0f; s_acc[1][threadIdx.x] = 0.0f; s_acc[2][threadIdx.x] = 0.0f;
        if (lane == 0u) s_next[w] = 0u;
    }
    __syncwarp();
    const uint32_t nvalid = (uint32_t)__popc(valid);
    const uint32_t total = nvalid * fp.n_samples;
    uint32_t tl = 0, sample = 0, tpix = 0;  // the item this lane is working on: tile lane, global sample, film index
    f3 acc = mk(0.f, 0.f, 0.f);
    PathState p;
    p.o = mk(0.f, 0.f, 0.f); p.d = mk(0.f, 1.f, 0.f); p.T = mk(1.f, 1.f, 1.f);
    p.pwr = 1.0f;
    p.bounce = MRT_NEED_PATH;
    bool have = false;
    for (;;) {
        if (p.bounce == MRT_NEED_PATH) {
            if (have) {  // hand the finished path's radiance to its pixel
                atomicAdd(&s_acc[0][wbase + tl], acc.x);
                atomicAdd(&s_acc[1][wbase + tl], acc.y);
                atomicAdd(&s_acc[2][wbase + tl], acc.z);
                have = false;
            }
            const uint32_t item = atomicAdd(&s_next[w], 1u);
            if (item >= total) break;
            // items run sample-major: all pixels of the tile for one sample, then the next sample
            const uint32_t jj = item / nvalid, k = item - jj * nvalid;
            tl = valid == 0xffffffffu ? k : (uint32_t)__fns(valid, 0u, (int)k + 1);
            sample = fp.sample0 + jj * fp.sample_stride;
            const float4 qs = s_q[wbase + tl];
            tpix = __float_as_uint(qs.w);
            const float2 u = rng_cam(cam_hash_seed(tpix, fp.key), sample);
            camera_ray(fp, xyz(qs), u.x, u.y, &p.o, &p.d);
            p.T = mk(1.f, 1.f, 1.f);
            p.pwr = 1.0f;
            p.bounce = 0;
            acc = mk(0.f, 0.f, 0.f);
            have = true;
        }
        if (path_segment<V, F>(sc, fp, tpix, sample, p, acc)) p.bounce = MRT_NEED_PATH;
    }
    __syncwarp();
    if (inside) {
        float4 a = fp.accum[pix];
        a.x += s_acc[0][threadIdx.x]; a.y += s_acc[1][threadIdx.x]; a.z += s_acc[2][threadIdx.x];
        fp.accum[pix] = a;
    }
}

// =====================================================================================================================
// The WAVE kernel: the megakernel's work re-cut for scenes that are searched through a BVH.
//
// ncu on the megakernel (profiles/r2_*): on Mesh.json 5.8 - 6.4 of 32 lanes are active per issued instruction, on
// Instance.json 9.1 - 9.7, on Minecraft.json 10.9 - 12.7 — and unbinding the lanes from their pixels (path_body_pool)
// does not change that.  The lanes are lost INSIDE the searches: every lane starts its ray's walk together, the walks
// last anything from one step (the ray misses everything) to a hundred, and a lane that is done waits for the
// longest walk of its warp, segment after segment.  So here a lane is bound neither to a pixel nor to a path:
//   * a warp keeps MRT_WAVE_SLOTS paths in flight, their state (ray, throughput, radiance, hit record) in shared
//     memory, and two queues of slot numbers: paths whose ray still has to be searched, paths whose search is done;
//   * SEARCH: a lane without a job takes the next slot from the first queue and walks its ray through the BVH one
//     step per turn of the warp's loop — scene nodes, primitives, and, inside a mesh instance, the triangle BVH, all in
//     the same loop; when the walk ends it goes on with the segment's shadow rays (one walk per light), then hands the
//     slot to the second queue and takes the next job.  Lanes finish at different turns and never wait for each other;
//   * SHADE: whenever 32 searched paths wait (or the search queue runs dry), every lane shades one of them
//     (shade_hit: the megakernel's own code), which ends the path — its radiance goes to its pixel, the slot gets the
//     warp's next (pixel, sample) item — or gives the slot its next ray; either way the slot returns to the first queue.
// Same rays, same tests, same tie rules, same RNG keys as the megakernel: a pixel's samples are merely summed in
// another order.  Everything is per warp: no block-level synchronisation, no global-memory queues.
// =====================================================================================================================
#ifndef MRT_WAVE_SLOTS
#define MRT_WAVE_SLOTS 64
#endif
#ifndef MRT_WAVE_STARVE
#define MRT_WAVE_STARVE 8   // lanes without work (and nothing queued for them) at which the warp turns to shading
#endif
#define MRT_WAVE_IDLE 0xffffffffu

// One lane's walk of one ray (registers).  The stack lives in local memory next to it.
struct Walk {
    uint32_t slot;    // MRT_WAVE_IDLE: no job
    uint32_t phase;   // 0: closest hit of the path's ray; 1 + li: shadow ray towards light li (any hit)
    f3 o, d;          // the ray in the space being searched: world, or the object space of the mesh instance entered
    f3 m, nom;        // 1/d with Box::intersect's fix-up (rt.rs:303-316) and -o * m: primitive tests, octree-leaf candidacy
    f3 bm, bnom;      // the geometrically true reciprocal (true_rcp3) and -o * bm: BVH node slabs
    Best B;           // scene-level best so far
    uint32_t cur;     // node / leaf to visit next, MRT_WALK_EMPTY: pop
    int sp;
    // the mesh instance being walked
    bool in_mesh;
    uint32_t m_first_tri, m_inst;
    float b0, b1;
    uint32_t r0, r1;
    int k0, k1;
};

__device__ __forceinline__ void walk_set_ray(Walk& w, f3 o, f3 d) {
    w.o = o; w.d = d;
    w.m = rcp_fixed3(d);
    w.nom = mk(-o.x * w.m.x, -o.y * w.m.y, -o.z * w.m.z);
    w.bm = true_rcp3(d, w.m);
    w.bnom = mk(-o.x * w.bm.x, -o.y * w.bm.y, -o.z * w.bm.z);
}

// Start the walk of (o, d) through the scene: planes first (infinite, not in the BVH), then the root.
// Returns true when the walk is already over (any-hit query answered by a plane).
template <uint32_t F>
__device__ __forceinline__ bool walk_begin(Walk& w, const GlobalScene& s, f3 o, f3 d) {
    constexpr bool WANT_T1 = (F & F_TRANSMIT) != 0;
    const SceneCommon& c = s.c;
    walk_set_ray(w, o, d);
    w.B.t0 = __int_as_float(0x7f800000); w.B.t1 = 0.0f; w.B.bi = -1; w.B.tr0 = w.B.tr1 = -1; w.B.any = false;
    w.in_mesh = false;
    w.sp = 0;
    w.cur = s.bvh_root;
    for (uint32_t k = 0; k < MRT_N_PLANES(c); k++) {  // same arithmetic and tie rule as closest_hit's plane loop
        const SlimInst e = ldg_slim(s.pln + k);
        const float t0 = (e.b.x - dot(o, xyz(e.a))) * frcp(dot(d, xyz(e.a)));
        best_update_lex<F, false, WANT_T1>(w.B, t0 > 0.0f, t0, t0, (int)(c.first[K_PLANE] + k), -1, -1);
    }
    return w.phase != 0u && w.B.bi >= 0;
}

// One step of a walk.  Returns true when the walk is over (w.B holds the answer; for a shadow ray: B.bi >= 0 = occluded).
// `wo`, `wd`: the walk's world-space ray (needed again when a mesh instance is left).
template <uint32_t F>
__device__ __forceinline__ bool walk_turn(Walk& w, uint32_t* stack, float* stack_t, const GlobalScene& s, f3 wo, f3 wd) {
    constexpr bool MESH = (F & F_MESH) != 0 && MRT_BVH_HAS_MESH;
    constexpr bool WANT_T1 = (F & F_TRANSMIT) != 0;
    const SceneCommon& c = s.c;
    const float INF = __int_as_float(0x7f800000);
    const bool any = w.phase != 0u;
    // behind this nothing can win ('<=': an equal t0 with a lower index / rank must still be found); inside a mesh whose
    // exit candidate is wanted nothing may be pruned
    auto far_bound = [&]() -> float {
        if (any) return INF;
        if (MESH && w.in_mesh) return WANT_T1 ? INF : fminf(w.b0, w.B.t0);
        return w.B.t0;
    };
    if (w.cur & MRT_BVH_LEAF) {
        const uint32_t ref = w.cur & ~MRT_BVH_LEAF;
        w.cur = MRT_WALK_EMPTY;
        if (MESH && w.in_mesh) {  // a triangle: hit, then candidacy (is one of the octree leaves that list it pierced?)
            const DTri* tp = &c.tri[w.m_first_tri + ref];
            DTri tr;
            tr.v0 = __ldg(&tp->v0); tr.e0 = __ldg(&tp->e0); tr.e1 = __ldg(&tp->e1);
            float t;
            if (tri_test(tr, w.o, w.d, &t) && (any || WANT_T1 || t <= w.b0)) {
                uint32_t rf = 0xffffffffu, rl = 0u;
                const f3 om = mk(-w.nom.x, -w.nom.y, -w.nom.z);
                if (tri_candidate<WANT_T1>(c, tr, w.m, om, &rf, &rl)) {
                    if (any) { w.B.bi = 0; return true; }
                    if (t < w.b0 || (t == w.b0 && rf < w.r0)) { w.b0 = t; w.r0 = rf; w.k0 = (int)ref; }
                    if constexpr (WANT_T1) { if (t > w.b1 || (t == w.b1 && rl >= w.r1)) { w.b1 = t; w.r1 = rl; w.k1 = (int)ref; } }
                }
            }
        } else if (MESH && (ref >> 28) == K_MESH) {  // a mesh instance: ray into object space, root AABB, then its triangle BVH
            const uint32_t k = ref & 0x0fffffffu;
            const SlimInst e = ldg_slim(s.mesh + k);
            f3 ol = w.o - xyz(e.a), dl = w.d;
            if (__float_as_uint(e.b.x) != 0u) { ol = mulXf(s.mesh_m[k], ol); dl = mulXf(s.mesh_m[k], w.d); }
            const DMesh mh = c.mesh[__float_as_uint(e.b.y)];
            const f3 m = rcp_fixed3(dl);
            const f3 om = ol * m;
            if (mesh_root_hit(mh, m, om)) {
                if (mh.bvh_root == 0xffffffffu) {  // no triangle BVH (MRT_NO_MESH_BVH): the sequential leaf walk, as one step
                    float t0 = 0.f, t1 = 0.f;
                    int tr0 = -1, tr1 = -1;
                    const bool hit = mesh_leaf_walk<false, WANT_T1>(c, mh, ol, dl, m, om, &t0, &t1, &tr0, &tr1);
                    best_update_lex<F, false, WANT_T1>(w.B, hit, t0, t1, (int)(c.first[K_MESH] + k), tr0, tr1);
                } else {
                    stack[w.sp] = MRT_WALK_MARKER; stack_t[w.sp] = 0.0f; w.sp++;
                    walk_set_ray(w, ol, dl);
                    w.in_mesh = true;
                    w.m_first_tri = mh.first_tri; w.m_inst = k;
                    w.b0 = INF; w.b1 = -INF; w.r0 = 0xffffffffu; w.r1 = 0u; w.k0 = -1; w.k1 = -1;
                    w.cur = mh.bvh_root;
                }
            }
        } else {  // any other primitive
            RayPre r;
            r.o = w.o; r.d = w.d; r.m = w.m; r.nom = w.nom;
            r.am = mk(fabsf(w.m.x), fabsf(w.m.y), fabsf(w.m.z));
            r.nam = -r.am;
            const RayPk rp = {pk2(w.o.x, w.d.x), pk2(w.o.y, w.d.y), pk2(w.o.z, w.d.z)};
            bvh_leaf<F, false, WANT_T1>(w.B, s, r, rp, ref);
        }
        if (any && w.B.bi >= 0) return true;
    } else if (w.cur != MRT_WALK_EMPTY) {
        const BvhNode* nd = ((MESH && w.in_mesh) ? c.tbvh : s.bvh) + w.cur;
        const float4 q0 = __ldg(&nd->q0), q1 = __ldg(&nd->q1), q2 = __ldg(&nd->q2);
        const uint2 ref = __ldg(reinterpret_cast<const uint2*>(&nd->ref));
        NodeRay nr;
        nr.bm = w.bm; nr.bnom = w.bnom;
        nr.bam = mk(fabsf(w.bm.x), fabsf(w.bm.y), fabsf(w.bm.z));
        float tl, tfl, tr, tfr;
        node_slabs(nr, q0, q1, q2, &tl, &tfl, &tr, &tfr);
        const float bound = far_bound();
        const bool hl = tl <= tfl && tfl >= 0.0f && tl <= bound;
        const bool hr = tr <= tfr && tfr >= 0.0f && tr <= bound;
        if (hl && hr) {
            const bool left_first = tl <= tr;
            if (w.sp < MRT_WALK_STACK) { stack[w.sp] = left_first ? ref.y : ref.x; stack_t[w.sp] = left_first ? tr : tl; w.sp++; }
            w.cur = left_first ? ref.x : ref.y;
        } else w.cur = hl ? ref.x : (hr ? ref.y : MRT_WALK_EMPTY);
    }
    while (w.cur == MRT_WALK_EMPTY) {  // next subtree; one that starts behind the best found since it was pushed holds nothing closer
        if (w.sp == 0) return true;
        --w.sp;
        const uint32_t ref = stack[w.sp];
        if (MESH && ref == MRT_WALK_MARKER) {  // the mesh instance is done: fold its candidates into the scene-level best
            best_update_lex<F, false, WANT_T1>(w.B, w.k0 >= 0, w.b0, WANT_T1 ? w.b1 : w.b0, (int)(c.first[K_MESH] + w.m_inst), w.k0, WANT_T1 ? w.k1 : w.k0);
            w.in_mesh = false;
            walk_set_ray(w, wo, wd);
        } else if (stack_t[w.sp] <= far_bound()) w.cur = ref;
    }
    return false;
}

template <class V, uint32_t F>
__device__ __forceinline__ void path_body_wave(const V sc, const FilmParams& fp) {
    constexpr uint32_t S = MRT_WAVE_SLOTS, NW = MRT_POOL_BLOCK / 32u;
    static_assert(S >= 32u && S <= 256u && (S & (S - 1u)) == 0u, "slots per warp: a power of two, 32 .. 256");
    constexpr bool WANT_T1 = (F & F_TRANSMIT) != 0;
    constexpr bool MESH = (F & F_MESH) != 0 && MRT_BVH_HAS_MESH;
    // per pixel of the block's tile
    __shared__ float4 s_q[MRT_POOL_BLOCK];        // focus vector (xyz) + film index (w, bits)
    __shared__ float s_acc[3][MRT_POOL_BLOCK];
    // per path slot (structure of arrays: lane i touches word i of a row)
    __shared__ float s_ray[6][NW * S];            // o, d
    __shared__ float s_T[3][NW * S];              // throughput
    __shared__ float s_L[3][NW * S];              // radiance of the path so far
    __shared__ float s_pwr[NW * S];
    __shared__ uint32_t s_bounce[NW * S], s_tl[NW * S], s_sample[NW * S];
    __shared__ float s_t0[NW * S], s_t1[NW * S];  // hit record of the path's current ray
    __shared__ int s_inst[NW * S], s_tri0[NW * S], s_tri1[NW * S];
    __shared__ uint32_t s_vis[NW * S];
    __shared__ uint8_t s_qtrav[NW][S], s_qshade[NW][S];

    const GlobalScene& gs = sc.s;
    const SceneCommon& c = sc.c();
    uint32_t px, py;
    thread_pixel(fp, &px, &py);
    const bool inside = px < fp.nw && py < fp.nh;
    const uint32_t lane = threadIdx.x & 31u, wbase = threadIdx.x & ~31u, wi = threadIdx.x >> 5, sbase = wi * S;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint32_t valid = __ballot_sync(0xffffffffu, inside);
    if (valid == 0u || fp.n_samples == 0u) return;  // warp-uniform
    const uint32_t pix = py * fp.nw + px;
    {
        const f3 q = inside ? pixel_focus_vec(fp, px, py) : mk(0.f, 1.f, 0.f);
        s_q[threadIdx.x] = make_float4(q.x, q.y, q.z, __uint_as_float(pix));
        s_acc[0][threadIdx.x] = 0.0f; s_acc[1][threadIdx.x] = 0.0f; s_acc[2][threadIdx.x] = 0.0f;
    }
    __syncwarp();
    const uint32_t nvalid = (uint32_t)__popc(valid);
    const uint32_t total = nvalid * fp.n_samples;  // the warp's (pixel, sample) items, sample-major

    // a new camera path for `slot` from item `item`
    auto start_path = [&](uint32_t slot, uint32_t item) {
        const uint32_t jj = item / nvalid, k = item - jj * nvalid;
        const uint32_t tl = valid == 0xffffffffu ? k : (uint32_t)__fns(valid, 0u, (int)k + 1);
        const uint32_t sample = fp.sample0 + jj * fp.sample_stride;
        const float4 qs = s_q[wbase + tl];
        const float2 u = rng_cam(cam_hash_seed(__float_as_uint(qs.w), fp.key), sample);
        f3 o, d;
        camera_ray(fp, xyz(qs), u.x, u.y, &o, &d);
        const uint32_t i = sbase + slot;
        s_ray[0][i] = o.x; s_ray[1][i] = o.y; s_ray[2][i] = o.z; s_ray[3][i] = d.x; s_ray[4][i] = d.y; s_ray[5][i] = d.z;
        s_T[0][i] = 1.0f; s_T[1][i] = 1.0f; s_T[2][i] = 1.0f;
        s_L[0][i] = 0.0f; s_L[1][i] = 0.0f; s_L[2][i] = 0.0f;
        s_pwr[i] = 1.0f; s_bounce[i] = 0u; s_tl[i] = tl; s_sample[i] = sample;
    };
    // world-space ray of job (slot, phase): the path's ray, or the shadow ray from its entry point towards light phase-1
    auto job_ray = [&](uint32_t slot, uint32_t phase, f3* o, f3* d) {
        const uint32_t i = sbase + slot;
        const f3 po = mk(s_ray[0][i], s_ray[1][i], s_ray[2][i]), pd = mk(s_ray[3][i], s_ray[4][i], s_ray[5][i]);
        if (phase == 0u) { *o = po; *d = pd; return; }
        const f3 hp = fma3(pd, s_t0[i], po);
        const f3 l = light_vec(c, phase - 1u, hp);
        *o = fma3(l, MRT_E, hp);  // Ray::cast_default, rt.rs:555-557
        *d = l;
    };

    // uniform across the warp: the pool of items and the two rings of slot numbers
    uint32_t pool_next = 0u, trav_head = 0u, trav_n = 0u, shade_head = 0u, shade_n = 0u;
    for (uint32_t base = 0u; base < S && pool_next < total; base += 32u) {  // the first paths
        const uint32_t slot = base + lane, item = pool_next + lane;
        if (item < total) { start_path(slot, item); s_qtrav[wi][(trav_n + lane) & (S - 1u)] = (uint8_t)slot; }
        const uint32_t n = min(32u, total - pool_next);
        pool_next += n; trav_n += n;
    }
    __syncwarp();

    uint32_t stack[MRT_WALK_STACK];
    float stack_t[MRT_WALK_STACK];
    Walk w;
    w.slot = MRT_WAVE_IDLE; w.phase = 0u; w.cur = MRT_WALK_EMPTY; w.sp = 0; w.in_mesh = false;
    w.o = w.d = w.m = w.nom = w.bm = w.bnom = mk(0.f, 0.f, 0.f);
    w.B.t0 = 0.f; w.B.t1 = 0.f; w.B.bi = -1; w.B.tr0 = w.B.tr1 = -1; w.B.any = false;
    w.m_first_tri = w.m_inst = 0u; w.b0 = w.b1 = 0.f; w.r0 = w.r1 = 0u; w.k0 = w.k1 = -1;

    // a walk is over: record its answer; go on with the segment's next search if there is one.  Returns true when the
    // slot has no search left (it then goes to the shade queue).
    auto walk_over = [&]() -> bool {
        for (;;) {
            const uint32_t i = sbase + w.slot;
            if (w.phase == 0u) {
                s_t0[i] = w.B.t0; s_t1[i] = w.B.t1; s_inst[i] = w.B.bi; s_tri0[i] = w.B.tr0; s_tri1[i] = w.B.tr1;
                s_vis[i] = 0u;
                if (w.B.bi < 0 || (F & F_LIGHTS) == 0 || MRT_N_LIGHTS(c) == 0u) return true;
            } else {
                if (w.B.bi < 0) s_vis[i] |= 1u << (w.phase - 1u);  // nothing in the way (no distance limit, rt.rs:1034-1038)
                if (w.phase >= MRT_N_LIGHTS(c)) return true;
            }
            w.phase++;
            f3 o, d;
            job_ray(w.slot, w.phase, &o, &d);
            if (!walk_begin<F>(w, gs, o, d)) return false;
        }
    };

    for (;;) {
        // ---------------------------------------------------------------- search
        for (;;) {
            const uint32_t idle = __ballot_sync(0xffffffffu, w.slot == MRT_WAVE_IDLE);
            bool done = false;
            if (idle != 0u && trav_n != 0u) {  // hand the queued slots to the lanes without a job
                const uint32_t rank = (uint32_t)__popc(idle & lt_mask);
                if (w.slot == MRT_WAVE_IDLE && rank < trav_n) {
                    w.slot = s_qtrav[wi][(trav_head + rank) & (S - 1u)];
                    w.phase = 0u;
                    f3 o, d;
                    job_ray(w.slot, 0u, &o, &d);
                    if (walk_begin<F>(w, gs, o, d)) done = walk_over();
                }
                const uint32_t take = min((uint32_t)__popc(idle), trav_n);
                trav_head = (trav_head + take) & (S - 1u); trav_n -= take;
            }
            const uint32_t busy = __ballot_sync(0xffffffffu, w.slot != MRT_WAVE_IDLE);
            if (busy == 0u) break;
            if (w.slot != MRT_WAVE_IDLE && !done) {
                f3 wo = w.o, wd = w.d;
                if (MESH && w.in_mesh) job_ray(w.slot, w.phase, &wo, &wd);
                if (walk_turn<F>(w, stack, stack_t, gs, wo, wd)) done = walk_over();
            }
            const uint32_t dm = __ballot_sync(0xffffffffu, done);
            if (dm != 0u) {  // searched: on to the shade queue
                if (done) {
                    s_qshade[wi][(shade_head + shade_n + (uint32_t)__popc(dm & lt_mask)) & (S - 1u)] = (uint8_t)w.slot;
                    w.slot = MRT_WAVE_IDLE;
                }
                shade_n += (uint32_t)__popc(dm);
            }
            if (shade_n >= 32u) break;
            if (trav_n == 0u && shade_n != 0u && 32u - (uint32_t)__popc(busy & ~dm) >= MRT_WAVE_STARVE) break;
        }
        __syncwarp();
        if (shade_n == 0u) {
            if (trav_n == 0u && __ballot_sync(0xffffffffu, w.slot != MRT_WAVE_IDLE) == 0u) break;  // everything rendered
            continue;
        }
        // ---------------------------------------------------------------- shade: one searched path per lane
        const uint32_t n = min(32u, shade_n);
        const bool mine = lane < n;
        uint32_t slot = 0u;
        bool ended = false;
        if (mine) {
            slot = s_qshade[wi][(shade_head + lane) & (S - 1u)];
            const uint32_t i = sbase + slot;
            PathState p;
            p.o = mk(s_ray[0][i], s_ray[1][i], s_ray[2][i]); p.d = mk(s_ray[3][i], s_ray[4][i], s_ray[5][i]);
            p.T = mk(s_T[0][i], s_T[1][i], s_T[2][i]);
            p.pwr = s_pwr[i]; p.bounce = s_bounce[i];
            f3 L = mk(s_L[0][i], s_L[1][i], s_L[2][i]);
            const uint32_t tl = s_tl[i];
            HitRec h;
            h.t0 = s_t0[i]; h.t1 = s_t1[i]; h.inst = s_inst[i]; h.tri0 = s_tri0[i]; h.tri1 = s_tri1[i];
            if (h.inst < 0) { path_miss(c, p, L); ended = true; }
            else ended = shade_hit<F>(c, fp, __float_as_uint(s_q[wbase + tl].w), s_sample[i], p, L, h, s_vis[i]);
            if (ended) {  // the path's radiance goes to its pixel
                atomicAdd(&s_acc[0][wbase + tl], L.x); atomicAdd(&s_acc[1][wbase + tl], L.y); atomicAdd(&s_acc[2][wbase + tl], L.z);
            } else {
                s_ray[0][i] = p.o.x; s_ray[1][i] = p.o.y; s_ray[2][i] = p.o.z; s_ray[3][i] = p.d.x; s_ray[4][i] = p.d.y; s_ray[5][i] = p.d.z;
                s_T[0][i] = p.T.x; s_T[1][i] = p.T.y; s_T[2][i] = p.T.z;
                s_L[0][i] = L.x; s_L[1][i] = L.y; s_L[2][i] = L.z;
                s_pwr[i] = p.pwr; s_bounce[i] = p.bounce;
            }
        }
        shade_head = (shade_head + n) & (S - 1u); shade_n -= n;
        // a slot whose path ended takes the warp's next item; slots with a ray to search go back to the search queue
        const uint32_t em = __ballot_sync(0xffffffffu, mine && ended);
        bool queue = mine && !ended;
        if (mine && ended) {
            const uint32_t item = pool_next + (uint32_t)__popc(em & lt_mask);
            if (item < total) { start_path(slot, item); queue = true; }
        }
        pool_next = min(total, pool_next + (uint32_t)__popc(em));
        const uint32_t qm = __ballot_sync(0xffffffffu, queue);
        if (queue) s_qtrav[wi][(trav_head + trav_n + (uint32_t)__popc(qm & lt_mask)) & (S - 1u)] = (uint8_t)slot;
        trav_n += (uint32_t)__popc(qm);
        __syncwarp();
    }
    __syncwarp();
    if (inside) {
        float4 a = fp.accum[pix];
        a.x += s_acc[0][threadIdx.x]; a.y += s_acc[1][threadIdx.x]; a.z += s_acc[2][threadIdx.x];
        fp.accum[pix] = a;
    }
}
