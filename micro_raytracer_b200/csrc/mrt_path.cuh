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
// `hp` = the entry point p.o + p.d * h.t0 (the caller's: the pinhole loop carries it instead of p.o and h.t0; p.o is only
// read for the exit point of a transmission).
template <uint32_t F>
__device__ __forceinline__ bool shade_hit(const SceneCommon& c, const FilmParams& fp, uint32_t pix, uint32_t sample, PathState& p, f3& acc,
                                          const HitRec& h, uint32_t vis, f3 hp) {
    const f3 o = p.o, d = p.d;
    MRT_CHECK(h.inst >= 0 && (uint32_t)h.inst < c.n_inst);
    const FatInst* fat = c.fat + h.inst;
    Surf s;
    load_surf(fat, &s);
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

// Light visibility from a hit's entry point: shadow rays without a distance limit, rt.rs:1027-1045.
template <class V, uint32_t F>
__device__ __forceinline__ uint32_t light_visibility(const V& sc, f3 o, f3 d, const HitRec& h) {
    uint32_t vis = 0;
    if constexpr ((F & F_LIGHTS) != 0) {
        const SceneCommon& c = sc.c();
        const f3 hp = fma3(d, h.t0, o);
        for (uint32_t li = 0; li < MRT_N_LIGHTS(c); li++) {
            const f3 l = light_vec(c, li, hp);
            HitRec dummy;
            if (!closest_hit<V, F, true, false>(sc, fma3(l, MRT_E, hp), l, &dummy)) vis |= 1u << li;
        }
    }
    return vis;
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
    if (MRT_HAS_SPHERE && MRT_REFINE_SPHERES(c)) refine_sphere_hit(c, p.o, p.d, &h);
    const uint32_t vis = light_visibility<V, F>(sc, p.o, p.d, h);
    return shade_hit<F>(c, fp, pix, sample, p, acc, h, vis, fma3(p.d, h.t0, p.o));
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

// The megakernel body for a PINHOLE camera (aperture 0; the reference's default is 0.001, so only frames that ask for it).
// The lens jitter is (u - 0.5) * 0, so every sample of a pixel starts with the same ray and finds the same first hit (and
// the same lights visible from it).  Ray, hit and visibility are computed once per pixel by the same arithmetic as
// camera_ray / path_segment (the same paths; images equal to the last bits, tests/test_gpu_parity.py), and the loop is
// rotated: an iteration is shade -> search, a new path starts at the cached hit.  Against the thin-lens loop below this removes
//   * the ~50-instruction path start (lens hash, normalisation, camera rotation) that ~4 lanes of a warp ran in 96 % of
//     the iterations (ncu, profiles/r2_path_kernel_jit_ncu_summary.txt: 12.5 % of the issued instructions at 6.7 lanes), and
//   * one closest-hit search per path: a path of k hits takes k iterations, whether it ends on a light, at the bounce
//     limit or by leaving the scene (before: k + 1 when it left the scene; CornellBox2: 6.71 -> 6.18 iterations per path).
template <class V, uint32_t F>
__device__ __forceinline__ void path_body_pinhole(const V sc, const FilmParams& fp) {
    // what a new path starts from, per thread, in shared memory (read by the few lanes that start a path in an iteration;
    // in registers it would cost the kernel a resident block per SM): ray direction + t0 of the first hit, instance,
    // visible lights, and for transmitting scenes / meshes the exit parameter and the triangles
    __shared__ float4 s_dt[128];
    __shared__ int s_inst[128];
    __shared__ uint32_t s_vis[(F & F_LIGHTS) != 0 ? 128 : 1];
    __shared__ float s_t1[(F & F_TRANSMIT) != 0 ? 128 : 1];
    __shared__ int2 s_tri[(F & F_MESH) != 0 ? 128 : 1];
    constexpr bool T1 = (F & F_TRANSMIT) != 0;
    uint32_t px, py;
    thread_pixel(fp, &px, &py);
    if (px >= fp.nw || py >= fp.nh || fp.n_samples == 0u) return;
    const SceneCommon& c = sc.c();
    const uint32_t pix = py * fp.nw + px;
    f3 acc = mk(0.f, 0.f, 0.f);
    PathState p;
    HitRec h;
    camera_ray(fp, pixel_focus_vec(fp, px, py), 0.5f, 0.5f, &p.o, &p.d);
    p.T = mk(1.f, 1.f, 1.f); p.pwr = 1.0f; p.bounce = 0;
    if (!closest_hit<V, F, false, T1>(sc, p.o, p.d, &h)) {
        // the pixel looks past the scene: every sample is sky.color (rt.rs:958), added one by one like the other loop does
        for (uint32_t j = 0; j < fp.n_samples; j++) path_miss(c, p, acc);
    } else {
        if (MRT_HAS_SPHERE && MRT_REFINE_SPHERES(c)) refine_sphere_hit(c, p.o, p.d, &h);
        uint32_t vis = light_visibility<V, F>(sc, p.o, p.d, h);
        s_dt[threadIdx.x] = make_float4(p.d.x, p.d.y, p.d.z, h.t0);
        s_inst[threadIdx.x] = h.inst;
        if constexpr ((F & F_LIGHTS) != 0) s_vis[threadIdx.x] = vis;
        if constexpr (T1) s_t1[threadIdx.x] = h.t1;
        if constexpr ((F & F_MESH) != 0) s_tri[threadIdx.x] = make_int2(h.tri0, h.tri1);
        f3 hp = mk(0.f, 0.f, 0.f);
        p.bounce = MRT_NEED_PATH;
        uint32_t j = 0;
        // (a path starts at the TOP of the iteration after the one it ended in, behind a state flag, as in the thin-lens
        // loop: written at the tail, where the path ends, the compiler nests the loops — an inner one per path — and the
        // lanes of a warp wait for each other's paths to end)
        for (;;) {
            if (p.bounce == MRT_NEED_PATH) {
                if (j >= fp.n_samples) break;
                uint32_t tid;  // re-read here: kept in a register across the loop it would be the 49th (one resident block per SM less)
                asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid));
                const float4 dt = s_dt[tid];
                p.d = mk(dt.x, dt.y, dt.z);
                p.o = fma3(p.d, MRT_E, mk(fp.cam_pos[0], fp.cam_pos[1], fp.cam_pos[2]));  // Ray::cast_default, rt.rs:555-557
                hp = fma3(p.d, dt.w, p.o);
                p.T = mk(1.f, 1.f, 1.f); p.pwr = 1.0f; p.bounce = 0;
                h.t0 = dt.w; h.inst = s_inst[tid];
                if constexpr ((F & F_LIGHTS) != 0) vis = s_vis[tid];
                if constexpr (T1) h.t1 = s_t1[tid];
                if constexpr ((F & F_MESH) != 0) { const int2 tr = s_tri[tid]; h.tri0 = tr.x; h.tri1 = tr.y; }
            }
            bool ended = shade_hit<F>(c, fp, pix, fp.sample0 + j * fp.sample_stride, p, acc, h, vis, hp);
            if (!ended) {
                if (!closest_hit<V, F, false, T1>(sc, p.o, p.d, &h)) {
                    path_miss(c, p, acc);
                    ended = true;
                } else {
                    if (MRT_HAS_SPHERE && MRT_REFINE_SPHERES(c)) refine_sphere_hit(c, p.o, p.d, &h);
                    vis = light_visibility<V, F>(sc, p.o, p.d, h);
                    hp = fma3(p.d, h.t0, p.o);
                }
            }
            if (ended) { p.bounce = MRT_NEED_PATH; j++; }
        }
    }
    MRT_CHECK(pix < fp.nw * fp.nh);
    float4 a = fp.accum[pix];
    a.x += acc.x; a.y += acc.y; a.z += acc.z;
    fp.accum[pix] = a;
}

// The megakernel body, one thread per supersampled pixel: the lane renders its pixel's samples back to back.
// LENS: how camera rays start — LENS_PINHOLE / LENS_THIN compile one case in (the specialised kernel has an entry point
// for each), LENS_ANY decides at run time (the offline-built generic kernels).
enum : int { LENS_PINHOLE = 0, LENS_THIN = 1, LENS_ANY = 2 };
template <class V, uint32_t F, int LENS = LENS_ANY>
__device__ __forceinline__ void path_body(const V sc, const FilmParams& fp) {
    if constexpr (LENS == LENS_PINHOLE) { path_body_pinhole<V, F>(sc, fp); return; }
    uint32_t px, py;
    thread_pixel(fp, &px, &py);
    if (px >= fp.nw || py >= fp.nh || fp.n_samples == 0u) return;
    const uint32_t pix = py * fp.nw + px;
    f3 q = pixel_focus_vec(fp, px, py);
    const uint32_t cam_seed = cam_hash_seed(pix, fp.key);
    // Pinhole camera (aperture 0 — not the reference's default, which is 0.001, parser.rs:206): the lens jitter is
    // (u - 0.5) * 0, so every sample of a pixel starts with the same ray.  The generic kernels compute its direction once,
    // here, by the very same arithmetic (camera_ray with zero jitter), and starting a path shrinks from ~50 instructions
    // — lens hash, normalisation, camera rotation; run by ~5 lanes of the warp in 96 % of the loop's iterations — to 11.
    // (The specialised kernel goes further for such frames: path_body_pinhole.)
    const bool pinhole = LENS == LENS_PINHOLE || (LENS == LENS_ANY && fp.aprt == 0.0f);  // warp-uniform
    if (pinhole) { f3 o0; camera_ray(fp, q, 0.5f, 0.5f, &o0, &q); }  // q := the pixel's ray direction

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
            if (pinhole) {
                // (the empty asm keeps the origin's three FFMA here: hoisted out of the loop it would hold three more
                // registers for the whole kernel, 56 instead of 48 = one resident block per SM less)
                float dx = q.x, dy = q.y, dz = q.z;
                asm volatile("" : "+f"(dx), "+f"(dy), "+f"(dz));
                p.d = mk(dx, dy, dz);
                p.o = fma3(p.d, MRT_E, mk(fp.cam_pos[0], fp.cam_pos[1], fp.cam_pos[2]));  // Ray::cast_default, rt.rs:555-557
            } else {
                const float2 u = rng_cam(cam_seed, fp.sample0 + j * fp.sample_stride);
                camera_ray(fp, q, u.x, u.y, &p.o, &p.d);
            }
            p.T = mk(1.f, 1.f, 1.f);
            p.pwr = 1.0f;
            p.bounce = 0;
        }
        const uint32_t sample = fp.sample0 + j * fp.sample_stride;
        if (path_segment<V, F>(sc, fp, pix, sample, p, acc)) { p.bounce = MRT_NEED_PATH; j++; }
    }
    MRT_CHECK(pix < fp.nw * fp.nh);
    float4 a = fp.accum[pix];
    a.x += acc.x; a.y += acc.y; a.z += acc.z;
    fp.accum[pix] = a;
}
