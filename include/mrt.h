/* mrt.h — C ABI of the B200-native path-tracing hot path of micro-raytracer.
 *
 * This header is the drop-in boundary.  It replaces the reference's `Sampler`
 * struct API (the narrowest seam the per-pixel path-tracing loop sits behind):
 *
 *   Sampler::new(workers, n_dim)            /root/reference/src/sampler.rs:19
 *   Sampler::execute(&scene, &frame, &rt)   /root/reference/src/sampler.rs:28
 *   Sampler::img(&frame)                    /root/reference/src/sampler.rs:80
 *
 * called from CLI::raytrace (src/cli.rs:155-177) and HttpServer::raytrace
 * (src/http.rs:136-148).  The descriptive structs below carry exactly the
 * fields of the reference's Render/Scene/Frame types (src/rt.rs:10-190), as
 * plain pointers and sizes.  All pointers are caller-owned HOST memory that is
 * copied during the call.  Every entry point returns 0 on success and a
 * non-zero mrt_status on error (message via mrt_last_error), never aborts —
 * this is the C form of the reference's `Result<_, String>` convention
 * (src/parser.rs:12-14).  One context = one caller thread; distinct contexts
 * are independent (the reference makes one Sampler per HTTP connection thread,
 * src/http.rs:138,155).
 *
 * The reference's call pattern is the first-class one: Sampler::new once, then
 * `for _ in 0..rt.sample { sampler.execute(&scene, &frame, &rt) }` and img()
 * (src/cli.rs:157-174, src/http.rs:138-147).  A host that maps it literally —
 * mrt_create_group(all devices) once; per pass mrt_update_scene, mrt_update_frame,
 * mrt_set_rt, mrt_execute(ctx, 1, &t); mrt_img at the end or after any pass —
 * gets every GPU of the box and full-length kernel launches: one-pass calls are
 * queued and rendered in launches of up to spp_per_launch passes per device
 * (MRT_OPT_COALESCE), the devices of a group split the samples, and mrt_img
 * gathers their films over NVLink peer mappings.
 *
 * There is NO CPU fallback behind this ABI: every entry point that computes
 * runs hand-written sm_100a CUDA kernels and fails with MRT_ERR_CUDA when no
 * device is usable.
 */
#ifndef MRT_H
#define MRT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MRT_ABI_VERSION 2

typedef enum mrt_status {
    MRT_OK = 0,
    MRT_ERR_INVALID = 1,   /* bad argument / scene the reference would panic on */
    MRT_ERR_CUDA = 2,      /* CUDA runtime error, or no usable device            */
    MRT_ERR_STATE = 3,     /* call order (e.g. execute before set_scene)         */
    MRT_ERR_NOMEM = 4
} mrt_status;

/* RendererKind, src/rt.rs:137-144 */
typedef enum mrt_kind {
    MRT_SPHERE = 0,   /* param[0] = r                          rt.rs:120 */
    MRT_PLANE = 1,    /* param[0..3] = n (not normalised)      rt.rs:123 */
    MRT_BOX = 2,      /* param[0..3] = sizes (full extents)    rt.rs:126 */
    MRT_TRIANGLE = 3, /* param[0..9] = v0,v1,v2                rt.rs:129 */
    MRT_MESH = 4      /* mesh = index into mrt_scene.meshes    rt.rs:132 */
} mrt_kind;

/* Material, src/rt.rs:88-103.  Texture handles index mrt_scene.textures, -1 = None. */
typedef struct mrt_material {
    float albedo[3];
    float rough, metal, glass, opacity, emit;
    int32_t tex, rmap, mmap, gmap, omap, emap;
} mrt_material;

/* Renderer, src/rt.rs:152-158 (the never-read `aabb` field is dropped). */
typedef struct mrt_object {
    uint32_t kind;        /* mrt_kind */
    uint32_t mesh;        /* MRT_MESH only */
    float param[9];
    uint32_t first_inst;  /* instances [first_inst, first_inst + n_inst) */
    uint32_t n_inst;
    mrt_material mat;
} mrt_object;

/* RendererInstance, src/rt.rs:146-150.  dir is (w, x, y, z) as serialised by
 * src/lin.rs:428-443: xyz = facing vector, w = sin(roll). */
typedef struct mrt_instance {
    float pos[3];
    float dir[4];
} mrt_instance;

/* Texture, src/rt.rs:81-86: row-major RGB f32 texels at texels[3*first_texel ...]. */
typedef struct mrt_texture {
    uint32_t w, h;
    uint64_t first_texel;
    uint32_t has_dat;     /* 0 ⇒ dat == None ⇒ every fetch returns zero (rt.rs:626) */
    uint32_t _pad;
} mrt_texture;

/* Mesh, src/rt.rs:131-135: triangles [first_tri, first_tri + n_tri), 9 floats each.
 * The depth-3 octree (src/parser.rs:805-824, src/rt.rs:630-703) is rebuilt by the library. */
typedef struct mrt_mesh {
    uint32_t first_tri;
    uint32_t n_tri;
} mrt_mesh;

/* Light, src/rt.rs:160-175 */
typedef enum mrt_light_kind { MRT_LIGHT_POINT = 0, MRT_LIGHT_DIR = 1 } mrt_light_kind;
typedef struct mrt_light {
    uint32_t kind;
    float v[3];          /* Point: pos, Dir: dir */
    float pwr;
    float color[3];
} mrt_light;

/* Scene, src/rt.rs:183-190 (renderer_bvh is always None, src/parser.rs:922). */
typedef struct mrt_scene {
    const mrt_object* objects;     uint32_t n_objects;
    const mrt_instance* instances; uint32_t n_instances;
    const mrt_texture* textures;   uint32_t n_textures;
    const float* texels;           uint64_t n_texels;     /* RGB triples */
    const mrt_mesh* meshes;        uint32_t n_meshes;
    const float* triangles;        uint32_t n_triangles;  /* 9 floats each */
    const mrt_light* lights;       uint32_t n_lights;
    float sky_color[3];
    float sky_pwr;
} mrt_scene;

/* Frame + Camera, src/rt.rs:63-79 */
typedef struct mrt_frame {
    uint16_t res[2];
    float ssaa;
    float cam_pos[3];
    float cam_dir[4];     /* (w, x, y, z) */
    float fov, gamma, exp, aprt, foc;
} mrt_frame;

/* Deterministic per-ray probe record (closest_hit of the primary ray, rt.rs:867-898). */
typedef struct mrt_hit {
    float t0, t1;         /* entry / exit parameter; NaN-free; t0 = -1 on miss */
    int32_t obj, inst;    /* object index, instance index within the object; -1 on miss */
    int32_t tri0, tri1;   /* mesh triangle index of entry / exit hit, -1 if none */
    float n0[3];          /* unit normal at the entry hit (rt.rs:776-793) */
    float n1[3];          /* unit normal at the exit hit */
    float uv[2];          /* to_uv of the entry hit point (rt.rs:795-809); 0 for meshes */
    float orig[3];        /* primary ray origin and direction (rt.rs:900-931) */
    float dir[3];
} mrt_hit;

typedef struct mrt_ctx mrt_ctx;

/* ≙ Sampler::new(workers, n_dim), sampler.rs:19.  `workers`/`n_dim` (--worker/--dim) are
 * accepted for signature parity and ignored: the CUDA grid replaces the tile pool. */
int mrt_create(mrt_ctx** out, int device, uint32_t workers, uint32_t n_dim);
/* The same Sampler over SEVERAL devices of the box (devices == NULL or n_devices <= 0: all of them; at most 16).
 * Every entry point below takes the group context like a plain one.  The group splits the passes of
 * mrt_execute over its devices (pass k goes to device k mod G; the RNG is keyed by pixel and global sample
 * index, so the image equals the one-device image up to the f32 summation order) and read-out (mrt_img,
 * mrt_img_ss, mrt_accum) sums the devices' films: each device tone-maps one band of pixels, reading that band
 * of every film over peer mappings (NVLink) and writing u8 pixels into the first device's image — no NCCL,
 * no host staging.  Devices without peer access fall back to staged device-to-device copies.  One device in
 * the list gives a plain context.  mrt_trace_primary, mrt_fp32_peak and mrt_set_stream act on the first device. */
int mrt_create_group(mrt_ctx** out, const int* devices, int n_devices, uint32_t workers, uint32_t n_dim);
/* Devices behind the context (1 for a plain one) and whether they map each other's memory. */
int mrt_group_info(mrt_ctx* ctx, uint32_t* n_devices, uint32_t* peer_access);
void mrt_destroy(mrt_ctx* ctx);
const char* mrt_last_error(const mrt_ctx* ctx);   /* ctx may be NULL: last create error */
int mrt_abi_version(void);
/* Number of usable CUDA devices (0 and MRT_ERR_CUDA without a driver): what a multi-GPU host
 * (one context per device + mrt_set_partition, see below) may pass to mrt_create. */
int mrt_device_count(int* n);

/* The three borrows of Sampler::execute (sampler.rs:28).  Data is validated, packed
 * into the device layout and uploaded.  Changing scene or frame drops the accumulated
 * passes (the reference would have to build a new Sampler). */
int mrt_set_scene(mrt_ctx* ctx, const mrt_scene* scene);
int mrt_set_frame(mrt_ctx* ctx, const mrt_frame* frame);
/* For hosts that, like Sampler::execute (sampler.rs:28), are handed scene and frame on EVERY pass: a
 * description identical to the one the context holds (64-bit content hash of every array / memcmp of the
 * frame) is a no-op that keeps the accumulated passes; anything else is mrt_set_scene / mrt_set_frame.
 * (Deviation: the reference keeps adding into the same buffer when the scene changes between passes,
 * sampler.rs:60-70 — here a changed scene or frame starts a new film.) */
int mrt_update_scene(mrt_ctx* ctx, const mrt_scene* scene);
int mrt_update_frame(mrt_ctx* ctx, const mrt_frame* frame);
/* RayTracer{bounce, loss} (rt.rs:16-22).  `seed` keys the counter-based RNG that stands
 * in for rand::thread_rng (unseedable in the reference). */
int mrt_set_rt(mrt_ctx* ctx, uint32_t bounce, float loss, uint64_t seed);

/* Behaviour switches.  An option takes effect at the next mrt_set_scene.
 *   MRT_OPT_NORMAL_SPACE   how Renderer::normal (rt.rs:776-793) returns the kind normal n of a
 *                          ROTATED instance (identity instances are unaffected):
 *     MRT_NORMAL_FORWARD_XF  norm(rot_y * (look * n)) — rt.rs:792 as written at HEAD (default)
 *     MRT_NORMAL_OBJECT      norm(n) — what the revision that rendered doc/out3.png (README's
 *                            CornellBox2 image) did; kept so that golden image stays reproducible.
 *   MRT_OPT_JIT            scene-specialised path kernel: for small scenes (those not searched through the BVH) the library can
 *                          compile the instance tables INTO the kernel at run time (NVRTC, ~0.15 s; cached per
 *                          scene in the process and as a cubin under $MRT_JIT_CACHE | ~/.cache/mrt_b200).
 *                          Same arithmetic, same results to rounding; only the instruction stream differs
 *                          (no table loads, no loop control): +17 % on the headline scene.
 *     MRT_JIT_AUTO   (default) the first mrt_execute after mrt_set_scene starts the compile on a background
 *                    thread and keeps rendering with the generic kernel; launches switch over once the
 *                    specialised kernel is ready.  A call of >= 2^33 paths waits for it.  Falls back to
 *                    the generic kernel if NVRTC is unavailable.
 *     MRT_JIT_OFF    never          MRT_JIT_FORCE   always, waiting for the compile (error if it fails).
 *                    Takes effect at the next mrt_execute.
 *   MRT_OPT_COALESCE       1 (default): mrt_execute(ctx, 1, ..) — the reference's one-pass call — only queues
 *                          the pass; see mrt_execute.  0: every call launches and waits (env MRT_COALESCE). */
typedef enum mrt_option { MRT_OPT_NORMAL_SPACE = 1, MRT_OPT_JIT = 2, MRT_OPT_COALESCE = 3 } mrt_option;
enum { MRT_NORMAL_FORWARD_XF = 0, MRT_NORMAL_OBJECT = 1 };
enum { MRT_JIT_OFF = 0, MRT_JIT_AUTO = 1, MRT_JIT_FORCE = 2 };
int mrt_set_option(mrt_ctx* ctx, uint32_t option, uint32_t value);

/* Sample split across PROCESSES (one context or group per process, e.g. one rank per GPU under torchrun, or one
 * group per node): this context renders global sample indices rank, rank + world, rank + 2*world, ...  Default
 * rank 0 of world 1.  The devices of one process need no partition: use mrt_create_group. */
int mrt_set_partition(mrt_ctx* ctx, uint32_t rank, uint32_t world);

/* ≙ n_passes × Sampler::execute (sampler.rs:28-78): adds n_passes paths per supersampled pixel to the film.
 *   n_passes == 1 (the reference's call, cli.rs:163 / http.rs:142; MRT_OPT_COALESCE on): the pass is QUEUED and the
 *     call returns at once.  Queued passes are rendered spp_per_launch (per device) at a time, and whenever something
 *     needs the film (mrt_img, mrt_img_ss, mrt_accum, mrt_accum_device, mrt_sync) or the settings they were asked
 *     under change (mrt_set_rt, mrt_set_partition) — so the loop `for _ in 0..1024 { execute }; img()` runs the same
 *     launches, bit for bit, as one mrt_execute(ctx, 1024, ..).  *seconds receives the AMORTISED device time: the
 *     CUDA-event time of the launches that finished since the last report (0 while nothing finished), so the
 *     per-pass log lines of cli.rs:164 still add up to the render time.
 *   n_passes != 1: queued passes + these are launched and the call blocks until the device is done; *seconds
 *     receives the device time not reported yet (for a group: the slowest device's).
 * mrt_film_size counts queued passes as rendered. */
int mrt_execute(mrt_ctx* ctx, uint32_t n_passes, double* seconds);
/* Launch queued passes + n_passes without waiting: the launches are queued on the context's stream(s). */
int mrt_execute_async(mrt_ctx* ctx, uint32_t n_passes);
/* Launch what is queued and wait for the device(s). */
int mrt_sync(mrt_ctx* ctx);
/* Device seconds (CUDA events) of all path launches of this context that have finished, since mrt_create. */
int mrt_device_seconds(mrt_ctx* ctx, double* total);

/* Drop the accumulated passes (≙ a fresh Sampler). */
int mrt_reset(mrt_ctx* ctx);

/* Geometry of the supersampled film: nw = (res.0 * ssaa) as usize (sampler.rs:29-30). */
int mrt_film_size(mrt_ctx* ctx, uint32_t* nw, uint32_t* nh, uint32_t* passes);

/* Linear accumulated sums, RGB f32, nw*nh*3 (what Sampler.colors holds, sampler.rs:14). */
int mrt_accum(mrt_ctx* ctx, float* rgb, uint32_t* passes);
/* Device view of the accumulator: nw*nh float4 (rgb + unused w), for an NCCL reduce by the
 * host (one process per GPU).  mrt_set_passes records the pass count the summed buffer holds.
 * A group first sums its devices' films onto its first device and hands that one out. */
int mrt_accum_device(mrt_ctx* ctx, void** dptr, size_t* n_floats, void** cuda_stream);
int mrt_set_passes(mrt_ctx* ctx, uint32_t passes);
/* Run this context's work on a caller-owned CUDA stream (a cudaStream_t / CUstream handle, e.g.
 * torch.cuda.current_stream().cuda_stream) so that the host can order it with its own events
 * and collectives.  NULL restores the context's private stream. */
int mrt_set_stream(mrt_ctx* ctx, void* cuda_stream);

/* Film gather ACROSS PROCESSES of one box (one context per process and GPU, e.g. one rank per GPU under torchrun; the
 * devices of ONE process use mrt_create_group instead).  Rather than reducing the whole 16-byte-per-pixel accumulator
 * onto one rank (NCCL) and tone-mapping there, every rank tone-maps one band of pixels: it reads that band of EVERY
 * rank's accumulator through CUDA IPC peer mappings (NVLink), sums in rank order and writes the u8 pixels straight into
 * the film rank's supersampled image.  Protocol (micro_raytracer_b200/distributed.py: gather_film):
 *   once per frame size   mrt_ipc_export on every rank -> exchange the 64-byte handles (any host channel) ->
 *                         mrt_ipc_attach(rank, world, all accumulator handles, the film rank's image handle)
 *   per read-out          [device-side barrier: every rank's passes are rendered] -> mrt_ipc_tonemap_band(total passes)
 *                         on every rank -> [device-side barrier: every band is written] -> mrt_img_gathered on the film rank.
 * The barriers are the caller's (e.g. a one-element NCCL all-reduce on the stream given to mrt_set_stream).
 * mrt_set_frame invalidates the mappings (export / attach again). */
#define MRT_IPC_HANDLE_BYTES 64
int mrt_ipc_export(mrt_ctx* ctx, uint8_t accum_handle[MRT_IPC_HANDLE_BYTES], uint8_t image_handle[MRT_IPC_HANDLE_BYTES]);
int mrt_ipc_attach(mrt_ctx* ctx, uint32_t rank, uint32_t world, const uint8_t* accum_handles /* world x 64 bytes, by rank */,
                   const uint8_t* film_image_handle /* rank 0's image_handle */);
int mrt_ipc_tonemap_band(mrt_ctx* ctx, uint32_t total_passes);
/* Film rank (rank 0), after every band has been written: Lanczos3 resize of the gathered image + copy to the host. */
int mrt_img_gathered(mrt_ctx* ctx, uint8_t* rgb);

/* ≙ Sampler::img (sampler.rs:80-99): ÷passes, powf(gamma), extended-Reinhard, `as u8`,
 * Lanczos3 resize nw×nh → res.  rgb = res.0*res.1*3 bytes. */
int mrt_img(mrt_ctx* ctx, uint8_t* rgb);
/* The u8 supersampled image before the resize (sampler.rs:84-96), nw*nh*3 bytes. */
int mrt_img_ss(mrt_ctx* ctx, uint8_t* rgb);

/* Deterministic probe: closest_hit of every primary ray with the lens jitter replaced by
 * the lens centre (u = 0.5).  out = nw*nh records, row-major. */
int mrt_trace_primary(mrt_ctx* ctx, mrt_hit* out);

/* Launch granularity of mrt_execute: at most `spp` passes are rendered by one kernel launch
 * (each launch reads and writes the accumulator once).  Results do not depend on it (the RNG is
 * keyed by the global sample index); default 1024, env MRT_SPP_PER_LAUNCH.  spp = 0 only queries. */
int mrt_spp_per_launch(mrt_ctx* ctx, uint32_t spp, uint32_t* current);

/* State of the scene-specialised kernel: eligible (scene small enough), compiled (ready and in use),
 * launches that used it, seconds the NVRTC compile took (negative: seconds to load the cubin from the
 * on-disk cache instead).  A compile error text is left in mrt_last_error. */
int mrt_jit_status(mrt_ctx* ctx, uint32_t* eligible, uint32_t* compiled, uint64_t* launches, double* compile_seconds);

/* How the library renders the scene it holds: through a scene-level BVH (scenes above ~60 box-equivalents; smaller
 * ones are unrolled into the kernel), with the run-time specialised kernel already in use, and the feature mask of the
 * kernel (1 lights, 2 textures, 4 transmission, 8 meshes). */
int mrt_scene_info(mrt_ctx* ctx, uint32_t* scene_bvh, uint32_t* specialised, uint32_t* features);

/* Counters of the kernels this context launched (bench.py's gpu_launches). */
int mrt_launch_count(mrt_ctx* ctx, uint64_t* n);
/* Measured FP32 FMA peak of the device this context lives on, in TFLOP/s
 * (dependent-free FFMA microbenchmark, CUDA events); used as a live roofline denominator. */
int mrt_fp32_peak(mrt_ctx* ctx, double* tflops, double* seconds);

#ifdef __cplusplus
}
#endif
#endif /* MRT_H */
