// mrt_kernels.cu — the sm_100a kernels of the path-tracing hot path.
//
//   path_kernel<View, F>   the megakernel: one thread per supersampled pixel; each lane runs
//                          its launch's samples back to back and starts its next camera path
//                          the moment the current one ends (per-lane path regeneration), so a
//                          warp only idles at the very end of the launch; radiance is summed
//                          in registers and written with one float4 read-modify-write per
//                          pixel per launch.  Replaces Sampler::execute + RayTracer::iter +
//                          reduce_light + RaytraceIterator::next (sampler.rs:28-78,
//                          rt.rs:937-994, 1014-1066).
//   primary_kernel<View>   deterministic probe of closest_hit on the primary rays.
//   film kernels           Sampler::img (sampler.rs:80-99): tonemap to u8, Lanczos3 resize.
//   fp32_peak_kernel       FFMA microbenchmark = live roofline denominator.
#include "mrt_path.cuh"
#include "mrt_kernels.h"

namespace {

template <uint32_t F>
__global__ void MRT_PATH_BOUNDS path_kernel_param(const __grid_constant__ ParamScene scene,
                                                                    const __grid_constant__ FilmParams fp) {
    path_body<ParamView, F>(ParamView{scene}, fp);
}
template <uint32_t F>
__global__ void MRT_PATH_BOUNDS path_kernel_global(const __grid_constant__ GlobalScene scene,
                                                                     const __grid_constant__ FilmParams fp) {
    path_body<GlobalView, F>(GlobalView{scene}, fp);
}

// ------------------------------------------------------------------ deterministic probe
__global__ void __launch_bounds__(128) primary_kernel(const __grid_constant__ GlobalScene scene,
                                                      const __grid_constant__ FilmParams fp,
                                                      mrt_hit* __restrict__ out, const uint32_t* __restrict__ obj_inst) {
    const uint32_t pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= fp.nw * fp.nh) return;
    const uint32_t py = pix / fp.nw, px = pix - py * fp.nw;
    GlobalView sc{scene};
    const SceneCommon& c = sc.c();
    const f3 q = pixel_focus_vec(fp, px, py);
    f3 o, d;
    camera_ray(fp, q, 0.5f, 0.5f, &o, &d);
    mrt_hit r;
    r.orig[0] = o.x; r.orig[1] = o.y; r.orig[2] = o.z;
    r.dir[0] = d.x; r.dir[1] = d.y; r.dir[2] = d.z;
    r.uv[0] = r.uv[1] = 0.0f;
    HitRec h;
    if (closest_hit<GlobalView, F_ALL, false, true>(sc, o, d, &h)) {
        if (c.refine_spheres) refine_sphere_hit(c, o, d, &h);
        const FatInst* fat = c.fat + h.inst;
        Surf s;
        load_surf(fat, &s);
        const f3 p0 = to_local(s, fma3(d, h.t0, o));
        const f3 p1 = to_local(s, fma3(d, h.t1, o));
        const f3 n0 = normalize(surf_normal<F_ALL>(c, s, p0, h.tri0));
        const f3 n1 = normalize(surf_normal<F_ALL>(c, s, p1, h.tri1));
        r.t0 = h.t0; r.t1 = h.t1;
        const uint32_t oi = obj_inst[h.inst];
        r.obj = (int)(oi & 0xffffu); r.inst = (int)(oi >> 16);
        r.tri0 = h.tri0; r.tri1 = h.tri1;
        r.n0[0] = n0.x; r.n0[1] = n0.y; r.n0[2] = n0.z;
        r.n1[0] = n1.x; r.n1[1] = n1.y; r.n1[2] = n1.z;
        if (s.kind() != K_MESH) {
            const float2 uv = surf_uv(s, p0);
            r.uv[0] = uv.x; r.uv[1] = uv.y;
        }
    } else {
        r.t0 = r.t1 = -1.0f;
        r.obj = r.inst = r.tri0 = r.tri1 = -1;
        r.n0[0] = r.n0[1] = r.n0[2] = r.n1[0] = r.n1[1] = r.n1[2] = 0.0f;
    }
    out[pix] = r;
}

// ------------------------------------------------------------------ film, sampler.rs:80-99
// `(255.0 * v) as u8`: saturating, NaN -> 0
__device__ __forceinline__ uint8_t to_u8(float v) {
    v *= 255.0f;
    if (!(v > 0.0f)) return 0;
    if (v >= 255.0f) return 255;
    return (uint8_t)v;
}
__device__ __forceinline__ uint8_t tonemap1(float sum, float inv_n, float gamma, float e2) {
    const float g = powf(sum * inv_n, gamma);                   // sampler.rs:85-88
    const float t = __fdiv_rn(g * (1.0f + __fdiv_rn(g, e2)), 1.0f + g);  // sampler.rs:91 (IEEE division kept under -prec-div=false)
    return to_u8(t);
}
__global__ void __launch_bounds__(256) tonemap_kernel(const float4* __restrict__ accum, uint8_t* __restrict__ out,
                                                      uint32_t npix, float inv_n, float gamma, float e2) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    const float4 a = accum[i];
    out[3 * i + 0] = tonemap1(a.x, inv_n, gamma, e2);
    out[3 * i + 1] = tonemap1(a.y, inv_n, gamma, e2);
    out[3 * i + 2] = tonemap1(a.z, inv_n, gamma, e2);
}

// Device groups (mrt_create_group): the film of a group is the SUM of its members' accumulators.  Each member
// tone-maps one band of pixels: it reads that band from every member's accumulator — its own from HBM, the others'
// over NVLink peer mappings — adds them in member order (so the result does not depend on which device runs the
// band) and writes the u8 pixels straight into the first member's supersampled image.  This is the group's only
// exchange step: a gather fused into the film kernel instead of a reduce of the whole 16-byte-per-pixel buffer
// followed by a tonemap.  Four pixels per thread: 4 x 16-byte loads per member, 12 bytes out as three 32-bit stores.
__global__ void __launch_bounds__(256) tonemap_peers_kernel(const PeerAccums acc, uint8_t* __restrict__ out, uint32_t first, uint32_t count,
                                                            float inv_n, float gamma, float e2) {
    const uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) * 4u;
    if (q >= count) return;
    const uint32_t n = min(4u, count - q);
    const uint32_t pix = first + q;
    uint8_t b[12];
    for (uint32_t k = 0; k < 4u; k++) {
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k < n)
            for (uint32_t m = 0; m < acc.n; m++) {
                const float4 a = acc.p[m][pix + k];
                s.x += a.x; s.y += a.y; s.z += a.z;
            }
        b[3 * k + 0] = tonemap1(s.x, inv_n, gamma, e2);
        b[3 * k + 1] = tonemap1(s.y, inv_n, gamma, e2);
        b[3 * k + 2] = tonemap1(s.z, inv_n, gamma, e2);
    }
    uint8_t* o = out + 3 * (size_t)pix;
    if (n == 4u && (reinterpret_cast<uintptr_t>(o) & 3u) == 0u) {
        uint32_t* o32 = reinterpret_cast<uint32_t*>(o);
        for (int k = 0; k < 3; k++) o32[k] = (uint32_t)b[4 * k] | (uint32_t)b[4 * k + 1] << 8 | (uint32_t)b[4 * k + 2] << 16 | (uint32_t)b[4 * k + 3] << 24;
    } else {
        for (uint32_t k = 0; k < 3u * n; k++) o[k] = b[k];
    }
}
// sum of the members' accumulators as packed RGB f32 (mrt_accum of a group)
__global__ void __launch_bounds__(256) unpack_peers_kernel(const PeerAccums acc, float* __restrict__ out, uint32_t npix) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (uint32_t m = 0; m < acc.n; m++) {
        const float4 a = acc.p[m][i];
        s.x += a.x; s.y += a.y; s.z += a.z;
    }
    out[3 * i] = s.x; out[3 * i + 1] = s.y; out[3 * i + 2] = s.z;
}
// dst += src (src: a peer's accumulator through its P2P mapping, or a staged copy of it)
__global__ void __launch_bounds__(256) accum_add_kernel(float4* __restrict__ dst, const float4* __restrict__ src, uint32_t npix) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    float4 d = dst[i];
    const float4 a = src[i];
    d.x += a.x; d.y += a.y; d.z += a.z;
    dst[i] = d;
}

// image 0.24 imageops::resize(.., Lanczos3) (sampler.rs:98): per output index the window
// [left, right) and its normalised weights.  One thread per output index; table row stride = max_taps.
__device__ __forceinline__ float sinc_(float t) {
    const float a = t * 3.14159265358979323846f;
    return t == 0.0f ? 1.0f : __fdiv_rn(sinf(a), a);
}
__device__ __forceinline__ float lanczos3_(float x) { return fabsf(x) < 3.0f ? sinc_(x) * sinc_(__fdiv_rn(x, 3.0f)) : 0.0f; }
__global__ void lanczos_weights_kernel(uint32_t in_n, uint32_t out_n, uint32_t max_taps,
                                       int32_t* __restrict__ left_out, int32_t* __restrict__ cnt_out, float* __restrict__ w_out) {
    const uint32_t o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= out_n) return;
    const float ratio = __fdiv_rn((float)in_n, (float)out_n);
    const float sratio = ratio < 1.0f ? 1.0f : ratio;
    const float support = 3.0f * sratio;
    float input = ((float)o + 0.5f) * ratio;
    long long left = (long long)floorf(input - support);
    left = left < 0 ? 0 : (left > (long long)in_n - 1 ? (long long)in_n - 1 : left);
    long long right = (long long)ceilf(input + support);
    right = right < left + 1 ? left + 1 : (right > (long long)in_n ? (long long)in_n : right);
    input -= 0.5f;
    int cnt = (int)(right - left);
    if (cnt > (int)max_taps) cnt = (int)max_taps;
    float* w = w_out + (size_t)o * max_taps;
    float sum = 0.0f;
    for (int i = 0; i < cnt; i++) {
        const float v = lanczos3_(__fdiv_rn((float)(left + i) - input, sratio));
        w[i] = v;
        sum = __fadd_rn(sum, v);
    }
    for (int i = 0; i < cnt; i++) w[i] = __fdiv_rn(w[i], sum);
    left_out[o] = (int32_t)left;
    cnt_out[o] = cnt;
}
// vertical_sample: u8 (w x h) -> f32 (w x nh), RGB
__global__ void __launch_bounds__(256) lanczos_vertical_kernel(const uint8_t* __restrict__ src, float* __restrict__ tmp,
                                                               uint32_t w, uint32_t nh, uint32_t max_taps,
                                                               const int32_t* __restrict__ left, const int32_t* __restrict__ cnt,
                                                               const float* __restrict__ wt) {
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t oy = blockIdx.y;
    if (x >= w || oy >= nh) return;
    const int l = left[oy], n = cnt[oy];
    const float* ws = wt + (size_t)oy * max_taps;
    float t0 = 0.f, t1 = 0.f, t2 = 0.f;
    for (int i = 0; i < n; i++) {
        const uint8_t* p = src + ((size_t)(l + i) * w + x) * 3;
        const float wi = ws[i];
        t0 = __fadd_rn(t0, __fmul_rn((float)p[0], wi));
        t1 = __fadd_rn(t1, __fmul_rn((float)p[1], wi));
        t2 = __fadd_rn(t2, __fmul_rn((float)p[2], wi));
    }
    float* q = tmp + ((size_t)oy * w + x) * 3;
    q[0] = t0; q[1] = t1; q[2] = t2;
}
// horizontal_sample: f32 (w x nh) -> u8 (nw x nh), clamp to [0,255] and round half away from zero
__global__ void __launch_bounds__(256) lanczos_horizontal_kernel(const float* __restrict__ tmp, uint8_t* __restrict__ dst,
                                                                 uint32_t w, uint32_t nw, uint32_t nh, uint32_t max_taps,
                                                                 const int32_t* __restrict__ left, const int32_t* __restrict__ cnt,
                                                                 const float* __restrict__ wt) {
    const uint32_t ox = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t y = blockIdx.y;
    if (ox >= nw || y >= nh) return;
    const int l = left[ox], n = cnt[ox];
    const float* ws = wt + (size_t)ox * max_taps;
    float t0 = 0.f, t1 = 0.f, t2 = 0.f;
    for (int i = 0; i < n; i++) {
        const float* p = tmp + ((size_t)y * w + (l + i)) * 3;
        const float wi = ws[i];
        t0 = __fadd_rn(t0, __fmul_rn(p[0], wi));
        t1 = __fadd_rn(t1, __fmul_rn(p[1], wi));
        t2 = __fadd_rn(t2, __fmul_rn(p[2], wi));
    }
    uint8_t* q = dst + ((size_t)y * nw + ox) * 3;
    q[0] = (uint8_t)roundf(fminf(fmaxf(t0, 0.0f), 255.0f));
    q[1] = (uint8_t)roundf(fminf(fmaxf(t1, 0.0f), 255.0f));
    q[2] = (uint8_t)roundf(fminf(fmaxf(t2, 0.0f), 255.0f));
}

// accumulator float4 -> packed RGB f32 (for mrt_accum)
__global__ void __launch_bounds__(256) unpack_accum_kernel(const float4* __restrict__ accum, float* __restrict__ out, uint32_t npix) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    const float4 a = accum[i];
    out[3 * i] = a.x; out[3 * i + 1] = a.y; out[3 * i + 2] = a.z;
}

// ------------------------------------------------------------------ FP32 peak microbenchmark
// 16 independent FFMA chains per thread, register operands only.
__global__ void __launch_bounds__(256) fp32_peak_kernel(float* out, int iters, float a, float b) {
    float v[16];
#pragma unroll
    for (int k = 0; k < 16; k++) v[k] = (float)(threadIdx.x + k);
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int k = 0; k < 16; k++) v[k] = fmaf(v[k], a, b);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 16; k++) s += v[k];
    if (s == 12345.678f) out[0] = s;  // never true; keeps the chains alive
}

template <uint32_t F>
cudaError_t launch_path_f(bool in_param, const ParamScene* ps, const GlobalScene* gs, const FilmParams& fp, cudaStream_t st) {
    static_assert(MRT_PATH_BLOCK == 128, "the tiled pixel mapping assumes blocks of four warps");
    const dim3 grid(path_grid_blocks(fp)), block(MRT_PATH_BLOCK);
    if (in_param) path_kernel_param<F><<<grid, block, 0, st>>>(*ps, fp);
    else path_kernel_global<F><<<grid, block, 0, st>>>(*gs, fp);
    return cudaGetLastError();
}

}  // namespace

cudaError_t mrt_launch_path(uint32_t features, bool in_param, const ParamScene* ps, const GlobalScene* gs,
                            const FilmParams& fp, cudaStream_t st) {
    switch (features & F_ALL) {
#define MRT_CASE(f) case f: return launch_path_f<f>(in_param, ps, gs, fp, st);
        MRT_CASE(0) MRT_CASE(1) MRT_CASE(2) MRT_CASE(3) MRT_CASE(4) MRT_CASE(5) MRT_CASE(6) MRT_CASE(7)
        MRT_CASE(8) MRT_CASE(9) MRT_CASE(10) MRT_CASE(11) MRT_CASE(12) MRT_CASE(13) MRT_CASE(14) MRT_CASE(15)
#undef MRT_CASE
    }
    return cudaErrorInvalidValue;
}

cudaError_t mrt_launch_primary(const GlobalScene& gs, const FilmParams& fp, mrt_hit* out, const uint32_t* obj_inst, cudaStream_t st) {
    const uint32_t npix = fp.nw * fp.nh;
    primary_kernel<<<(npix + 127) / 128, 128, 0, st>>>(gs, fp, out, obj_inst);
    return cudaGetLastError();
}

cudaError_t mrt_launch_tonemap(const float4* accum, uint8_t* out, uint32_t npix, float inv_n, float gamma, float exp, cudaStream_t st) {
    const float d = 1.0f - exp;
    tonemap_kernel<<<(npix + 255) / 256, 256, 0, st>>>(accum, out, npix, inv_n, gamma, d * d);
    return cudaGetLastError();
}

cudaError_t mrt_launch_tonemap_peers(const PeerAccums& acc, uint8_t* out, uint32_t first, uint32_t count, float inv_n, float gamma, float exp, cudaStream_t st) {
    if (!count) return cudaSuccess;
    const float d = 1.0f - exp;
    const uint32_t threads = (count + 3u) / 4u;
    tonemap_peers_kernel<<<(threads + 255) / 256, 256, 0, st>>>(acc, out, first, count, inv_n, gamma, d * d);
    return cudaGetLastError();
}
cudaError_t mrt_launch_unpack_peers(const PeerAccums& acc, float* out, uint32_t npix, cudaStream_t st) {
    unpack_peers_kernel<<<(npix + 255) / 256, 256, 0, st>>>(acc, out, npix);
    return cudaGetLastError();
}
cudaError_t mrt_launch_accum_add(float4* dst, const float4* src, uint32_t npix, cudaStream_t st) {
    accum_add_kernel<<<(npix + 255) / 256, 256, 0, st>>>(dst, src, npix);
    return cudaGetLastError();
}

cudaError_t mrt_launch_unpack(const float4* accum, float* out, uint32_t npix, cudaStream_t st) {
    unpack_accum_kernel<<<(npix + 255) / 256, 256, 0, st>>>(accum, out, npix);
    return cudaGetLastError();
}

cudaError_t mrt_launch_lanczos_weights(uint32_t in_n, uint32_t out_n, uint32_t max_taps, int32_t* left, int32_t* cnt, float* w, cudaStream_t st) {
    lanczos_weights_kernel<<<(out_n + 127) / 128, 128, 0, st>>>(in_n, out_n, max_taps, left, cnt, w);
    return cudaGetLastError();
}
cudaError_t mrt_launch_lanczos_vertical(const uint8_t* src, float* tmp, uint32_t w, uint32_t nh, uint32_t max_taps,
                                        const int32_t* left, const int32_t* cnt, const float* wt, cudaStream_t st) {
    lanczos_vertical_kernel<<<dim3((w + 255) / 256, nh), 256, 0, st>>>(src, tmp, w, nh, max_taps, left, cnt, wt);
    return cudaGetLastError();
}
cudaError_t mrt_launch_lanczos_horizontal(const float* tmp, uint8_t* dst, uint32_t w, uint32_t nw, uint32_t nh, uint32_t max_taps,
                                          const int32_t* left, const int32_t* cnt, const float* wt, cudaStream_t st) {
    lanczos_horizontal_kernel<<<dim3((nw + 255) / 256, nh), 256, 0, st>>>(tmp, dst, w, nw, nh, max_taps, left, cnt, wt);
    return cudaGetLastError();
}
cudaError_t mrt_launch_fp32_peak(float* out, int blocks, int iters, cudaStream_t st) {
    fp32_peak_kernel<<<blocks, 256, 0, st>>>(out, iters, 0.999f, 0.001f);
    return cudaGetLastError();
}
