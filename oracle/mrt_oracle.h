/* mrt_oracle.h — entry points of the CPU parity oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a literal CPU restatement of the reference's
 * per-pixel path-tracing path (/root/reference/src/{lin,rt,sampler}.rs).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` leg may load
 * it, and only as the checker / the timed CPU baseline.  The product library
 * (micro_raytracer_b200/libmrt.so) never links, loads or calls anything in oracle/.
 *
 * Parity pinning: the Rust reference cannot be built in this environment (no cargo/rustc),
 * and it ships no tests; the oracle is pinned against the reference's own rendered images
 * doc/out0.png … out4.png (tests/test_oracle_golden.py).  Scene features those images do
 * not reach (box-atlas textures, o/e/r/m/g maps, dir lights, instances, meshes) are
 * "parity unpinned": only the literal restatement anchors them.
 *
 * Scene/frame structs are the boundary types of include/mrt.h.
 */
#ifndef MRT_ORACLE_H
#define MRT_ORACLE_H

#include "../include/mrt.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mrt_cpu_ctx mrt_cpu_ctx;

enum {
    MRT_CPU_LITERAL = 0, /* reduce_light exactly as rt.rs:956-994: count-trace + re-trace + reverse fold */
    MRT_CPU_FORWARD = 1  /* single trace, forward accumulation, stop at the first passing emission draw */
};

typedef struct mrt_cpu_stats {
    uint64_t paths;        /* reduce_light evaluations */
    uint64_t segments;     /* closest_hit calls for path segments (not shadow rays) */
    uint64_t hits;         /* segments that hit */
    uint64_t shadow_rays;  /* closest_hit calls for light visibility */
    uint64_t nan_normals;  /* hits whose normal came out non-finite (Box::normal no-match, rt.rs:421-443) */
    uint64_t hit_hist[34]; /* histogram of hits per path (forward mode: up to termination) */
} mrt_cpu_stats;

int mrt_cpu_create(mrt_cpu_ctx** out, uint32_t workers, uint32_t n_dim);
void mrt_cpu_destroy(mrt_cpu_ctx* ctx);
const char* mrt_cpu_last_error(const mrt_cpu_ctx* ctx);

int mrt_cpu_set_scene(mrt_cpu_ctx* ctx, const mrt_scene* scene);
int mrt_cpu_set_frame(mrt_cpu_ctx* ctx, const mrt_frame* frame);
int mrt_cpu_set_rt(mrt_cpu_ctx* ctx, uint32_t bounce, float loss, uint64_t seed);
int mrt_cpu_set_partition(mrt_cpu_ctx* ctx, uint32_t rank, uint32_t world);
int mrt_cpu_set_mode(mrt_cpu_ctx* ctx, int mode);
int mrt_cpu_set_option(mrt_cpu_ctx* ctx, uint32_t option, uint32_t value);

int mrt_cpu_execute(mrt_cpu_ctx* ctx, uint32_t n_passes, double* seconds);
int mrt_cpu_reset(mrt_cpu_ctx* ctx);
int mrt_cpu_film_size(mrt_cpu_ctx* ctx, uint32_t* nw, uint32_t* nh, uint32_t* passes);
int mrt_cpu_accum(mrt_cpu_ctx* ctx, float* rgb, uint32_t* passes);
int mrt_cpu_img(mrt_cpu_ctx* ctx, uint8_t* rgb);
int mrt_cpu_img_ss(mrt_cpu_ctx* ctx, uint8_t* rgb);
int mrt_cpu_trace_primary(mrt_cpu_ctx* ctx, mrt_hit* out);
int mrt_cpu_get_stats(mrt_cpu_ctx* ctx, mrt_cpu_stats* out);

/* One path, for unit tests: radiance of (pixel x,y ; global sample index). */
int mrt_cpu_path(mrt_cpu_ctx* ctx, uint32_t x, uint32_t y, uint32_t sample, float rgb[3]);

/* image 0.24 `imageops::resize(.., FilterType::Lanczos3)` on RGB8 (sampler.rs:98). */
int mrt_cpu_resize_lanczos3(const uint8_t* src, uint32_t w, uint32_t h,
                            uint8_t* dst, uint32_t nw, uint32_t nh);
/* sampler.rs:85-94 on one linear value triple. */
void mrt_cpu_tonemap(const float* rgb_mean, float gamma, float exp, uint8_t* out, size_t n_values);

/* The counter-based RNG both implementations share: 4 uniforms of block `block`
 * for (pixel, sample). */
void mrt_cpu_rng_block(uint32_t pixel, uint32_t sample, uint32_t block, uint64_t seed, float out[4]);

#ifdef __cplusplus
}
#endif
#endif
