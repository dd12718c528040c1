// mrt_api.cu — the C ABI of include/mrt.h: contexts and device groups, launch scheduling (pass coalescing,
// sample split over the devices of a group), film read-out.  Scene packing lives in mrt_scene.cu.
// Host code only; every pixel is computed by the kernels in mrt_kernels.cu / the run-time specialised kernel.
// There is no CPU fallback: without a usable CUDA device every compute entry point fails.
//
// A context renders through one device (mrt_create) or through several (mrt_create_group: one member context per
// device, the reference's single `Sampler` spanning all of them).  The code below is written once for both: a
// plain context is "its own only member".
#include "mrt_ctx.h"

#include <thread>

namespace {

thread_local std::string g_create_err;

inline bool is_group(const mrt_ctx* c) { return !c->members.empty(); }
inline size_t n_render(const mrt_ctx* c) { return is_group(c) ? c->members.size() : 1; }
inline mrt_ctx* render_ctx(mrt_ctx* c, size_t i) { return is_group(c) ? c->members[i] : c; }
inline mrt_ctx* film_ctx(mrt_ctx* c) { return render_ctx(c, 0); }  // the device the film is read out on
// run `call` (an expression in `m`) on every member of group `c`, handing a member's error up
#define MEMBERS(call) do { for (mrt_ctx* m : c->members) { const int rc__ = (call); if (rc__) { c->err = m->err; return rc__; } } } while (0)
#define FWD(call) do { mrt_ctx* m = film_ctx(c); const int rc__ = (call); if (rc__) c->err = m->err; return rc__; } while (0)

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
    fp.tiles_x = c->knobs.tiled ? (c->nw + 15u) / 16u : 0u;
    return fp;
}

// ---------------------------------------------------------------- run-time specialised kernel: state per (member) context
// Looks the kernel of c->jit_header up (starting its NVRTC compile on a background thread at the first call).
// wait_ms < 0 blocks until the compile is over.
int jit_poll(mrt_ctx* c, int wait_ms) {
    const bool want = !c->jit_header.empty() && c->jit_mode != MRT_JIT_OFF;
    if (!want || c->jit_kernel || c->jit_failed) return MRT_OK;
    MrtJitInfo info;
    c->jit_kernel = mrt_jit_kernel(c->jit_header, wait_ms, &info);
    c->jit_requested = true;
    if (!info.pending) {
        c->jit_seconds = info.seconds;
        c->jit_from_disk = info.from_disk;
        c->jit_err = info.err;
        c->jit_failed = !c->jit_kernel;
        if (c->jit_failed && c->jit_mode == MRT_JIT_FORCE) return fail(c, MRT_ERR_CUDA, "scene specialisation failed: " + c->jit_err);
    }
    return MRT_OK;
}
inline bool jit_pending(const mrt_ctx* c) {
    return c->jit_mode == MRT_JIT_AUTO && !c->jit_header.empty() && !c->jit_kernel && !c->jit_failed;
}

// Queue the launches that render `count` samples sample0, sample0 + stride, ... of every pixel on the
// (single-device) context `c`: at most spp_per_launch samples per launch, each launch one read-modify-write of the
// accumulator.  MRT_JIT_AUTO never stalls a small call: the first call after set_scene starts the NVRTC compile on a
// background thread (or finds the cubin in the process / on-disk cache) and launches switch over as soon as it is
// ready; a call big enough to amortise the ~0.15 s compile (>= 2^33 paths) waits for it.  MRT_JIT_FORCE waits.
int launch_samples(mrt_ctx* c, uint32_t sample0, uint32_t stride, uint32_t count) {
    CK(cudaSetDevice(c->device));
    FilmParams fp = make_film_params(c);
    const uint64_t paths = (uint64_t)c->nw * c->nh * count;
    const bool force = c->jit_mode == MRT_JIT_FORCE || paths >= (1ull << 33);
    if (int rc = jit_poll(c, force ? -1 : (c->jit_requested ? 0 : 3))) return rc;  // 3 ms: a cubin on disk is ready by then
    const bool use_jit = c->jit_kernel && c->jit_mode != MRT_JIT_OFF && !c->jit_header.empty();
    uint32_t left = count, s0 = sample0;
    while (left) {
        const uint32_t n = std::min(left, c->spp_per_launch);
        fp.sample0 = s0;
        fp.sample_stride = stride;
        fp.n_samples = n;
        cudaError_t e = use_jit ? (c->gscene.bvh ? mrt_jit_launch_bvh(c->jit_kernel, c->gscene, fp, c->stream, c->knobs.pinhole)
                                                 : mrt_jit_launch(c->jit_kernel, c->gscene.c, fp, c->stream, c->knobs.pinhole))
                                : mrt_launch_path(c->features, c->in_param, c->pscene, &c->gscene, fp, c->stream);
        if (e != cudaSuccess) return cuda_fail(c, e, "path kernel launch");
        if (use_jit) c->jit_launches++;
        c->launches++;
        s0 += n * stride;
        left -= n;
    }
    return MRT_OK;
}

cudaEvent_t take_event(mrt_ctx* m) {  // current device = m->device
    if (!m->event_pool.empty()) { cudaEvent_t e = m->event_pool.back(); m->event_pool.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    if (cudaEventCreate(&e) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return e;
}

// Render the next `count` passes of context / group `c`.  Pass k of `c` is global sample rank + k * world
// (mrt_set_partition); inside a group pass k belongs to member k mod G, so every member renders a strided subset and
// the image does not depend on G beyond the f32 summation order (the RNG is keyed by pixel and global sample index).
// While the scene-specialised kernel is still compiling (MRT_JIT_AUTO) and the caller may block, the generic kernel
// is fed in slices of ~2^27 paths per device and the compile is looked at again after each, so a one-shot render
// (the CLI case) switches over after ~0.15 s instead of finishing on the slower kernel.
int run_passes(mrt_ctx* c, uint32_t count, bool may_block) {
    if (!count) return MRT_OK;
    const uint32_t G = (uint32_t)n_render(c);
    mrt_ctx::Round round;
    round.ev.assign(G, {nullptr, nullptr});
    for (uint32_t i = 0; i < G; i++) {
        mrt_ctx* m = render_ctx(c, i);
        CK(cudaSetDevice(m->device));
        round.ev[i].first = take_event(m);
        if (round.ev[i].first) CK(cudaEventRecord(round.ev[i].first, m->stream));
    }
    uint32_t k = c->passes, left = count;
    while (left) {
        bool pend = false;
        for (uint32_t i = 0; i < G; i++) pend |= jit_pending(render_ctx(c, i));
        uint32_t n = left;
        if (pend && may_block) {
            const uint64_t npix = std::max<uint64_t>(1, (uint64_t)film_ctx(c)->nw * film_ctx(c)->nh);
            n = (uint32_t)std::min<uint64_t>(left, std::max<uint64_t>(1, ((uint64_t)G << 27) / npix));
        }
        for (uint32_t i = 0; i < G; i++) {
            const uint32_t kk = k + (i + G - k % G) % G;  // first pass >= k that belongs to member i
            if (kk >= k + n) continue;
            const uint32_t cnt = (k + n - kk + G - 1u) / G;
            mrt_ctx* m = render_ctx(c, i);
            const int rc = launch_samples(m, c->rank + kk * c->world, G * c->world, cnt);
            if (rc) { if (m != c) c->err = m->err; return rc; }
        }
        k += n; left -= n;
        if (pend && may_block && left)
            for (uint32_t i = 0; i < G; i++) { mrt_ctx* m = render_ctx(c, i); CK(cudaSetDevice(m->device)); CK(cudaStreamSynchronize(m->stream)); }
    }
    for (uint32_t i = 0; i < G; i++) {
        mrt_ctx* m = render_ctx(c, i);
        CK(cudaSetDevice(m->device));
        round.ev[i].second = take_event(m);
        if (round.ev[i].second) CK(cudaEventRecord(round.ev[i].second, m->stream));
    }
    c->timing.push_back(std::move(round));
    c->passes += count;
    c->passes_total += count;
    return MRT_OK;
}

int flush(mrt_ctx* c, bool may_block) {
    const uint32_t n = c->pending;
    c->pending = 0;
    return run_passes(c, n, may_block);
}

// Read back the device time of finished rounds (wait: of all rounds).  A round's time is the slowest member's.
int harvest(mrt_ctx* c, bool wait) {
    while (!c->timing.empty()) {
        mrt_ctx::Round& r = c->timing.front();
        if (!wait)
            for (auto& ev : r.ev) {
                if (!ev.second) continue;
                const cudaError_t q = cudaEventQuery(ev.second);
                if (q == cudaErrorNotReady) { cudaGetLastError(); return MRT_OK; }
                if (q != cudaSuccess) return cuda_fail(c, q, "cudaEventQuery");
            }
        double t = 0.0;
        for (size_t i = 0; i < r.ev.size(); i++) {
            mrt_ctx* m = render_ctx(c, i);
            if (r.ev[i].first && r.ev[i].second) {
                CK(cudaSetDevice(m->device));
                CK(cudaEventSynchronize(r.ev[i].second));
                float ms = 0.0f;
                CK(cudaEventElapsedTime(&ms, r.ev[i].first, r.ev[i].second));
                t = std::max(t, (double)ms * 1e-3);
            }
            if (r.ev[i].first) m->event_pool.push_back(r.ev[i].first);
            if (r.ev[i].second) m->event_pool.push_back(r.ev[i].second);
        }
        c->unreported_s += t;
        c->total_s += t;
        c->timing.pop_front();
    }
    return MRT_OK;
}

int sync_streams(mrt_ctx* c) {
    for (size_t i = 0; i < n_render(c); i++) {
        mrt_ctx* m = render_ctx(c, i);
        CK(cudaSetDevice(m->device));
        CK(cudaStreamSynchronize(m->stream));
    }
    return MRT_OK;
}

// ---------------------------------------------------------------- device groups: the exchange step
// Order every member's stream after everything queued so far on every other member's stream (`ev` picks the event).
int group_fence(mrt_ctx* c, cudaEvent_t mrt_ctx::*ev) {
    for (mrt_ctx* m : c->members) { CK(cudaSetDevice(m->device)); CK(cudaEventRecord(m->*ev, m->stream)); }
    for (mrt_ctx* m : c->members) {
        CK(cudaSetDevice(m->device));
        for (mrt_ctx* o : c->members) if (o != m) CK(cudaStreamWaitEvent(m->stream, o->*ev, 0));
    }
    return MRT_OK;
}
PeerAccums peer_accums(const mrt_ctx* c) {
    PeerAccums pa{};
    pa.n = (uint32_t)c->members.size();
    for (uint32_t i = 0; i < pa.n; i++) pa.p[i] = c->members[i]->d_accum.p;
    return pa;
}
// Sum every member's accumulator into the first member's and zero the others: the group's film then lives on one
// device (mrt_accum_device for an outer reduce; groups whose devices cannot map each other's memory).
int group_collapse(mrt_ctx* c) {
    mrt_ctx* f = film_ctx(c);
    const uint32_t npix = f->nw * f->nh;
    if (int rc = group_fence(c, &mrt_ctx::ev_sync)) return rc;
    CK(cudaSetDevice(f->device));
    for (size_t i = 1; i < c->members.size(); i++) {
        mrt_ctx* m = c->members[i];
        const float4* src = m->d_accum.p;
        if (!c->p2p) {
            CK(c->d_stage.alloc(npix));
            CK(cudaMemcpyPeerAsync(c->d_stage.p, f->device, m->d_accum.p, m->device, (size_t)npix * sizeof(float4), f->stream));
            src = c->d_stage.p;
        }
        CK(mrt_launch_accum_add(f->d_accum.p, src, npix, f->stream));
        f->launches++;
    }
    CK(cudaEventRecord(f->ev_band, f->stream));
    for (size_t i = 1; i < c->members.size(); i++) {
        mrt_ctx* m = c->members[i];
        CK(cudaSetDevice(m->device));
        CK(cudaStreamWaitEvent(m->stream, f->ev_band, 0));
        CK(cudaMemsetAsync(m->d_accum.p, 0, m->d_accum.n * sizeof(float4), m->stream));
    }
    return MRT_OK;
}

void ipc_detach(mrt_ctx* c) {
    for (void* p : c->ipc_mapped) cudaIpcCloseMemHandle(p);
    cudaGetLastError();
    c->ipc_mapped.clear();
    c->ipc_world = 0;
    c->ipc_image = nullptr;
}

int ready_to_render(mrt_ctx* c, const char* what) {
    if (!c->have_scene || !c->have_frame) return fail(c, MRT_ERR_STATE, std::string(what) + " before set_scene/set_frame");
    return MRT_OK;
}

// u8 supersampled image of the accumulated film into film_ctx(c)->d_ss (sampler.rs:84-96)
int tonemap_ss(mrt_ctx* c) {
    if (!c->have_frame) return fail(c, MRT_ERR_STATE, "img before set_frame");
    if (c->pending && c->have_scene) { if (int rc = flush(c, true)) return rc; }
    if (c->passes_total == 0) return fail(c, MRT_ERR_STATE, "img before any pass");
    mrt_ctx* f = film_ctx(c);
    const uint32_t npix = f->nw * f->nh;
    const float inv_n = 1.0f / (float)c->passes_total;  // Vec3f / f32 = v * (1/n), lin.rs:296-302
    CK(cudaSetDevice(f->device));
    CK(f->d_ss.alloc((size_t)npix * 3));
    if (!is_group(c) || !c->p2p) {
        if (is_group(c)) { if (int rc = group_collapse(c)) return rc; }
        CK(cudaSetDevice(f->device));
        CK(mrt_launch_tonemap(f->d_accum.p, f->d_ss.p, npix, inv_n, f->frame.gamma, f->frame.exp, f->stream));
        f->launches++;
        return MRT_OK;
    }
    // gather fused into the tonemap: member i sums pixel band i of every accumulator over the peer mappings and
    // writes the u8 pixels into the film device's image (bands start at multiples of four pixels)
    if (int rc = group_fence(c, &mrt_ctx::ev_sync)) return rc;
    const PeerAccums pa = peer_accums(c);
    const uint32_t G = (uint32_t)c->members.size();
    const uint32_t band = ((npix + G - 1u) / G + 3u) & ~3u;
    for (uint32_t i = 0; i < G; i++) {
        mrt_ctx* m = c->members[i];
        const uint32_t first = std::min(npix, i * band), count = std::min(npix - first, band);
        CK(cudaSetDevice(m->device));
        CK(mrt_launch_tonemap_peers(pa, f->d_ss.p, first, count, inv_n, f->frame.gamma, f->frame.exp, m->stream));
        if (count) m->launches++;
    }
    // the film device continues once every band is written; nobody overwrites an accumulator a band still reads
    return group_fence(c, &mrt_ctx::ev_band);
}

// Sampler::img's resize (sampler.rs:98) of f->d_ss on the (single-device) film context, then the copy to the host
int resize_and_copy(mrt_ctx* c, uint8_t* rgb) {
    CK(cudaSetDevice(c->device));
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
    *out = nullptr;
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
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_sync, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_band, cudaEventDisableTiming);
    if (e != cudaSuccess) {
        g_create_err = std::string("mrt_create: ") + cudaGetErrorString(e);
        mrt_destroy(c);
        return MRT_ERR_CUDA;
    }
    c->knobs.read();
    if (const char* s = std::getenv("MRT_SPP_PER_LAUNCH")) {
        const int v = std::atoi(s);
        if (v > 0) c->spp_per_launch = (uint32_t)v;
    }
    if (const char* s = std::getenv("MRT_JIT")) {  // default MRT_OPT_JIT of new contexts (experiments, CI)
        const int v = std::atoi(s);
        if (v >= 0 && v <= (int)MRT_JIT_FORCE) c->jit_mode = (uint32_t)v;
    }
    if (const char* s = std::getenv("MRT_COALESCE")) c->coalesce = std::atoi(s) != 0;
    c->pscene = new ParamScene();
    *out = c;
    return MRT_OK;
}

int mrt_create_group(mrt_ctx** out, const int* devices, int n_devices, uint32_t workers, uint32_t n_dim) {
    if (!out) { g_create_err = "mrt_create_group: null out"; return MRT_ERR_INVALID; }
    *out = nullptr;
    int avail = 0;
    const cudaError_t e = cudaGetDeviceCount(&avail);
    if (e != cudaSuccess || avail == 0) {
        g_create_err = std::string("mrt_create_group: no CUDA device (") + (e != cudaSuccess ? cudaGetErrorString(e) : "count 0") + "); there is no CPU fallback";
        return MRT_ERR_CUDA;
    }
    std::vector<int> devs;
    if (!devices || n_devices <= 0) for (int d = 0; d < avail; d++) devs.push_back(d);  // every device of the box
    else devs.assign(devices, devices + n_devices);
    if (devs.size() > MRT_MAX_GROUP) { g_create_err = "mrt_create_group: more than 16 devices"; return MRT_ERR_INVALID; }
    for (size_t i = 0; i < devs.size(); i++)
        for (size_t j = 0; j < i; j++)
            if (devs[i] == devs[j]) { g_create_err = "mrt_create_group: a device is listed twice"; return MRT_ERR_INVALID; }
    if (devs.size() == 1) return mrt_create(out, devs[0], workers, n_dim);
    // The first touch of a device creates its primary context (0.3 s each on a B200 box) and each peer mapping costs
    // milliseconds: one thread per device does both, so a one-shot render on 8 GPUs waits for the slowest device, not
    // for the sum (measured: Sampler::new over 8 GPUs 2.5 s done one after the other).
    const bool want_p2p = !std::getenv("MRT_NO_P2P");
    std::vector<int> peer_ok(devs.size(), 1);
    {
        std::vector<std::thread> th;
        for (size_t i = 0; i < devs.size(); i++)
            th.emplace_back([&, i] {
                if (cudaSetDevice(devs[i]) != cudaSuccess || cudaFree(nullptr) != cudaSuccess) { cudaGetLastError(); peer_ok[i] = 0; return; }
                for (size_t j = 0; j < devs.size() && want_p2p; j++) {
                    if (i == j) continue;
                    int can = 0;
                    if (cudaDeviceCanAccessPeer(&can, devs[i], devs[j]) != cudaSuccess || !can) { peer_ok[i] = 0; break; }
                    const cudaError_t pe = cudaDeviceEnablePeerAccess(devs[j], 0);
                    if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) { peer_ok[i] = 0; break; }
                }
                cudaGetLastError();
            });
        for (auto& t : th) t.join();
    }
    mrt_ctx* g = new mrt_ctx();
    for (int d : devs) {
        mrt_ctx* m = nullptr;
        const int rc = mrt_create(&m, d, workers, n_dim);
        if (rc) { mrt_destroy(g); return rc; }
        g->members.push_back(m);
    }
    g->device = devs[0];
    g->knobs = g->members[0]->knobs;
    g->coalesce = g->members[0]->coalesce;
    g->spp_per_launch = g->members[0]->spp_per_launch;
    g->jit_mode = g->members[0]->jit_mode;
    // peer mappings, both directions of every pair (NVLink on an NVSwitch box; enabled by the threads above); without
    // them the exchange step falls back to staged copies
    g->p2p = want_p2p;
    for (int ok : peer_ok) g->p2p = g->p2p && ok != 0;
    *out = g;
    return MRT_OK;
}

int mrt_group_info(mrt_ctx* c, uint32_t* n_devices, uint32_t* peer_access) {
    if (!c) return MRT_ERR_INVALID;
    if (n_devices) *n_devices = (uint32_t)n_render(c);
    if (peer_access) *peer_access = is_group(c) ? (c->p2p ? 1u : 0u) : 1u;
    return MRT_OK;
}

void mrt_destroy(mrt_ctx* c) {
    if (!c) return;
    if (is_group(c)) {
        cudaSetDevice(c->device);
        c->d_stage.release();
        for (mrt_ctx* m : c->members) mrt_destroy(m);
        delete c;
        return;
    }
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->jit_requested && !c->jit_header.empty()) mrt_jit_wait(c->jit_header);
    for (auto& b : c->d_slim) b.release();
    c->d_boxp.release(); c->d_bvh.release(); c->d_bxf.release(); c->d_mesh_m.release(); c->d_fat.release(); c->d_tex.release(); c->d_texels.release();
    c->d_mesh.release(); c->d_leaf.release(); c->d_leaf_idx.release(); c->d_tri.release(); c->d_tbvh.release(); c->d_tri_leaf.release(); c->d_obj_inst.release();
    c->d_accum.release(); c->d_ss.release(); c->d_out.release(); c->d_tmp.release(); c->d_rgb.release();
    c->d_wv.release(); c->d_wh.release(); c->d_lv.release(); c->d_cv.release(); c->d_lh.release(); c->d_ch.release();
    c->d_hits.release(); c->d_stage.release();
    ipc_detach(c);
    for (auto& r : c->timing) for (auto& ev : r.ev) { if (ev.first) cudaEventDestroy(ev.first); if (ev.second) cudaEventDestroy(ev.second); }
    for (cudaEvent_t ev : c->event_pool) cudaEventDestroy(ev);
    for (cudaEvent_t ev : {c->ev0, c->ev1, c->ev_sync, c->ev_band}) if (ev) cudaEventDestroy(ev);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
}

int mrt_reset(mrt_ctx* c) {
    if (!c) return MRT_ERR_INVALID;
    c->passes = 0;
    c->passes_total = 0;
    c->pending = 0;
    if (is_group(c)) { MEMBERS(mrt_reset(m)); return MRT_OK; }
    CK(cudaSetDevice(c->device));
    if (c->d_accum.p) CK(cudaMemsetAsync(c->d_accum.p, 0, c->d_accum.n * sizeof(float4), c->stream));
    return MRT_OK;
}

int mrt_set_scene(mrt_ctx* c, const mrt_scene* s) {
    if (!c || !s) return MRT_ERR_INVALID;
    if (is_group(c)) {
        c->have_scene = false;
        for (mrt_ctx* m : c->members) m->normal_space = c->normal_space;
        {   // every member packs and uploads on its own thread (independent contexts: packing is host work, ~0.1 - 1 ms each)
            std::vector<int> rcs(c->members.size(), MRT_OK);
            std::vector<std::thread> th;
            for (size_t i = 0; i < c->members.size(); i++) th.emplace_back([&, i] { rcs[i] = mrt_scene_upload(c->members[i], s); });
            for (auto& t : th) t.join();
            for (size_t i = 0; i < rcs.size(); i++) if (rcs[i]) { c->err = c->members[i]->err; return rcs[i]; }
        }
        c->scene_hash = c->members[0]->scene_hash;
        c->have_scene = true;
        return mrt_reset(c);
    }
    if (int rc = mrt_scene_upload(c, s)) return rc;
    return mrt_reset(c);
}

int mrt_update_scene(mrt_ctx* c, const mrt_scene* s) {
    if (!c || !s) return MRT_ERR_INVALID;
    if (c->have_scene && c->scene_hash == mrt_scene_hash(s, c->normal_space)) return MRT_OK;
    return mrt_set_scene(c, s);
}

int mrt_set_frame(mrt_ctx* c, const mrt_frame* f) {
    if (!c || !f) return MRT_ERR_INVALID;
    uint32_t nw, nh;
    film_dims(*f, &nw, &nh);
    if (nw == 0 || nh == 0 || f->res[0] == 0 || f->res[1] == 0) return fail(c, MRT_ERR_INVALID, "empty film");
    if ((uint64_t)nw * nh > 0x7fffffffull) return fail(c, MRT_ERR_INVALID, "film too large");
    if (is_group(c)) {
        c->have_frame = false;
        MEMBERS(mrt_set_frame(m, f));
        c->frame = *f;
        c->nw = nw; c->nh = nh;
        c->have_frame = true;
        return mrt_reset(c);
    }
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    c->have_frame = false;  // stays false if the allocation fails: no launch on a missing accumulator
    c->pending = 0;
    ipc_detach(c);          // the buffers other processes mapped may move
    cudaError_t e = c->d_accum.alloc((size_t)nw * nh);
    if (e != cudaSuccess) return cuda_fail(c, e, "accumulator allocation");
    c->frame = *f;
    c->nw = nw; c->nh = nh;
    c->weights_ready = false;
    c->have_frame = true;
    return mrt_reset(c);
}

int mrt_update_frame(mrt_ctx* c, const mrt_frame* f) {
    if (!c || !f) return MRT_ERR_INVALID;
    if (c->have_frame && std::memcmp(&c->frame, f, sizeof *f) == 0) return MRT_OK;
    return mrt_set_frame(c, f);
}

int mrt_set_rt(mrt_ctx* c, uint32_t bounce, float loss, uint64_t seed) {
    if (!c) return MRT_ERR_INVALID;
    if (bounce > 0x3fffffffu) return fail(c, MRT_ERR_INVALID, "bounce too large");
    if (c->bounce == bounce && c->seed == seed && std::memcmp(&c->loss, &loss, sizeof loss) == 0) return MRT_OK;
    if (c->pending && c->have_scene && c->have_frame) { if (int rc = flush(c, true)) return rc; }  // queued passes were asked for under the old settings
    c->bounce = bounce; c->loss = loss; c->seed = seed;
    if (is_group(c)) MEMBERS(mrt_set_rt(m, bounce, loss, seed));
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
        if (is_group(c)) MEMBERS(mrt_set_option(m, option, value));
        return MRT_OK;
    }
    if (option == MRT_OPT_COALESCE && value <= 1u) {
        if (!value && c->pending && c->have_scene && c->have_frame) { if (int rc = flush(c, true)) return rc; }
        c->coalesce = value != 0u;
        return MRT_OK;
    }
    return fail(c, MRT_ERR_INVALID, "unknown option or value");
}

int mrt_set_partition(mrt_ctx* c, uint32_t rank, uint32_t world) {
    if (!c) return MRT_ERR_INVALID;
    if (world == 0 || rank >= world) return fail(c, MRT_ERR_INVALID, "bad partition");
    if (c->rank == rank && c->world == world) return MRT_OK;
    if (c->pending && c->have_scene && c->have_frame) { if (int rc = flush(c, true)) return rc; }
    c->rank = rank; c->world = world;
    return MRT_OK;
}

int mrt_execute_async(mrt_ctx* c, uint32_t n_passes) {
    if (!c) return MRT_ERR_INVALID;
    if (int rc = ready_to_render(c, "execute")) return rc;
    const uint32_t n = c->pending + n_passes;
    c->pending = 0;
    return run_passes(c, n, false);
}

int mrt_sync(mrt_ctx* c) {
    if (!c) return MRT_ERR_INVALID;
    if (c->pending && c->have_scene && c->have_frame) { if (int rc = flush(c, true)) return rc; }
    if (int rc = sync_streams(c)) return rc;
    return harvest(c, true);
}

int mrt_execute(mrt_ctx* c, uint32_t n_passes, double* seconds) {
    if (!c) return MRT_ERR_INVALID;
    if (int rc = ready_to_render(c, "execute")) return rc;
    if (n_passes == 1 && c->coalesce) {
        // ≙ one Sampler::execute.  The pass is queued; launches go out spp_per_launch passes (per device) at a time,
        // or when something needs the film.  The first call already starts the background scene specialisation.
        for (size_t i = 0; i < n_render(c); i++) {
            mrt_ctx* m = render_ctx(c, i);
            if (m->jit_mode == MRT_JIT_AUTO) { const int rc = jit_poll(m, m->jit_requested ? 0 : 3); if (rc) { c->err = m->err; return rc; } }
        }
        c->pending++;
        if ((uint64_t)c->pending >= (uint64_t)c->spp_per_launch * n_render(c)) { if (int rc = flush(c, true)) return rc; }
        if (int rc = harvest(c, false)) return rc;
    } else {
        const uint32_t n = c->pending + n_passes;
        c->pending = 0;
        if (int rc = run_passes(c, n, true)) return rc;
        if (int rc = harvest(c, true)) return rc;
    }
    if (seconds) { *seconds = c->unreported_s; c->unreported_s = 0.0; }
    return MRT_OK;
}

int mrt_device_seconds(mrt_ctx* c, double* total) {
    if (!c || !total) return MRT_ERR_INVALID;
    if (int rc = harvest(c, false)) return rc;
    *total = c->total_s;
    return MRT_OK;
}

int mrt_film_size(mrt_ctx* c, uint32_t* nw, uint32_t* nh, uint32_t* passes) {
    if (!c) return MRT_ERR_INVALID;
    if (nw) *nw = c->nw;
    if (nh) *nh = c->nh;
    if (passes) *passes = c->passes_total + c->pending;
    return MRT_OK;
}

int mrt_accum(mrt_ctx* c, float* rgb, uint32_t* passes) {
    if (!c || !rgb) return MRT_ERR_INVALID;
    if (!c->have_frame) return fail(c, MRT_ERR_STATE, "accum before set_frame");
    if (c->pending && c->have_scene) { if (int rc = flush(c, true)) return rc; }
    mrt_ctx* f = film_ctx(c);
    const uint32_t npix = f->nw * f->nh;
    const bool gather = is_group(c) && c->p2p;
    if (gather) { if (int rc = group_fence(c, &mrt_ctx::ev_sync)) return rc; }
    else if (is_group(c)) { if (int rc = group_collapse(c)) return rc; }
    CK(cudaSetDevice(f->device));
    CK(f->d_rgb.alloc((size_t)npix * 3));
    if (gather) CK(mrt_launch_unpack_peers(peer_accums(c), f->d_rgb.p, npix, f->stream));
    else CK(mrt_launch_unpack(f->d_accum.p, f->d_rgb.p, npix, f->stream));
    f->launches++;
    CK(cudaMemcpyAsync(rgb, f->d_rgb.p, (size_t)npix * 3 * sizeof(float), cudaMemcpyDeviceToHost, f->stream));
    CK(cudaStreamSynchronize(f->stream));
    if (gather) { if (int rc = group_fence(c, &mrt_ctx::ev_band)) return rc; }  // later launches wait for the gather
    if (passes) *passes = c->passes_total;
    return harvest(c, false);
}

int mrt_accum_device(mrt_ctx* c, void** dptr, size_t* n_floats, void** cuda_stream) {
    if (!c) return MRT_ERR_INVALID;
    if (!c->have_frame) return fail(c, MRT_ERR_STATE, "accum_device before set_frame");
    if (c->pending && c->have_scene) { if (int rc = flush(c, false)) return rc; }
    if (is_group(c)) { if (int rc = group_collapse(c)) return rc; }
    mrt_ctx* f = film_ctx(c);
    if (dptr) *dptr = f->d_accum.p;
    if (n_floats) *n_floats = (size_t)f->nw * f->nh * 4;
    if (cuda_stream) *cuda_stream = (void*)f->stream;
    return MRT_OK;
}

int mrt_set_stream(mrt_ctx* c, void* cuda_stream) {
    if (!c) return MRT_ERR_INVALID;
    if (c->pending && c->have_scene && c->have_frame) { if (int rc = flush(c, true)) return rc; }
    if (is_group(c)) { if (int rc = sync_streams(c)) return rc; FWD(mrt_set_stream(m, cuda_stream)); }  // the film device's stream
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    c->stream = cuda_stream ? (cudaStream_t)cuda_stream : c->own_stream;
    return MRT_OK;
}

int mrt_set_passes(mrt_ctx* c, uint32_t passes) {
    if (!c) return MRT_ERR_INVALID;
    if (c->pending && c->have_scene && c->have_frame) { if (int rc = flush(c, false)) return rc; }
    c->passes_total = passes;
    return MRT_OK;
}

int mrt_ipc_export(mrt_ctx* c, uint8_t* accum_handle, uint8_t* image_handle) {
    if (!c || !accum_handle || !image_handle) return MRT_ERR_INVALID;
    if (is_group(c)) return fail(c, MRT_ERR_INVALID, "mrt_ipc_*: one context per process and device (a group gathers its own devices)");
    if (!c->have_frame) return fail(c, MRT_ERR_STATE, "ipc_export before set_frame");
    static_assert(sizeof(cudaIpcMemHandle_t) == MRT_IPC_HANDLE_BYTES, "CUDA IPC handles are 64 bytes");
    CK(cudaSetDevice(c->device));
    CK(c->d_ss.alloc((size_t)c->nw * c->nh * 3));
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, c->d_accum.p));
    std::memcpy(accum_handle, &h, sizeof h);
    CK(cudaIpcGetMemHandle(&h, c->d_ss.p));
    std::memcpy(image_handle, &h, sizeof h);
    return MRT_OK;
}

int mrt_ipc_attach(mrt_ctx* c, uint32_t rank, uint32_t world, const uint8_t* accum_handles, const uint8_t* film_image_handle) {
    if (!c || !accum_handles || !film_image_handle) return MRT_ERR_INVALID;
    if (is_group(c)) return fail(c, MRT_ERR_INVALID, "mrt_ipc_*: one context per process and device");
    if (!c->have_frame) return fail(c, MRT_ERR_STATE, "ipc_attach before set_frame");
    if (world == 0 || world > MRT_MAX_GROUP || rank >= world) return fail(c, MRT_ERR_INVALID, "bad rank / world (at most 16 ranks)");
    CK(cudaSetDevice(c->device));
    ipc_detach(c);
    CK(c->d_ss.alloc((size_t)c->nw * c->nh * 3));
    c->ipc_accums = PeerAccums{};
    c->ipc_accums.n = world;
    for (uint32_t r = 0; r < world; r++) {
        if (r == rank) { c->ipc_accums.p[r] = c->d_accum.p; continue; }
        cudaIpcMemHandle_t h;
        std::memcpy(&h, accum_handles + (size_t)r * MRT_IPC_HANDLE_BYTES, sizeof h);
        void* p = nullptr;
        const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) { ipc_detach(c); return cuda_fail(c, e, "cudaIpcOpenMemHandle (accumulator)"); }
        c->ipc_mapped.push_back(p);
        c->ipc_accums.p[r] = static_cast<const float4*>(p);
    }
    if (rank == 0) c->ipc_image = c->d_ss.p;
    else {
        cudaIpcMemHandle_t h;
        std::memcpy(&h, film_image_handle, sizeof h);
        void* p = nullptr;
        const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) { ipc_detach(c); return cuda_fail(c, e, "cudaIpcOpenMemHandle (film image)"); }
        c->ipc_mapped.push_back(p);
        c->ipc_image = static_cast<uint8_t*>(p);
    }
    c->ipc_rank = rank; c->ipc_world = world;
    return MRT_OK;
}

int mrt_ipc_tonemap_band(mrt_ctx* c, uint32_t total_passes) {
    if (!c) return MRT_ERR_INVALID;
    if (c->ipc_world == 0) return fail(c, MRT_ERR_STATE, "ipc_tonemap_band before ipc_attach");
    if (total_passes == 0) return fail(c, MRT_ERR_STATE, "img before any pass");
    if (c->pending && c->have_scene) { if (int rc = flush(c, false)) return rc; }
    CK(cudaSetDevice(c->device));
    const uint32_t npix = c->nw * c->nh, G = c->ipc_world;
    const uint32_t band = ((npix + G - 1u) / G + 3u) & ~3u;
    const uint32_t first = std::min(npix, c->ipc_rank * band), count = std::min(npix - first, band);
    c->passes_total = total_passes;
    CK(mrt_launch_tonemap_peers(c->ipc_accums, c->ipc_image, first, count, 1.0f / (float)total_passes, c->frame.gamma, c->frame.exp, c->stream));
    if (count) c->launches++;
    return MRT_OK;
}

int mrt_img_gathered(mrt_ctx* c, uint8_t* rgb) {
    if (!c || !rgb) return MRT_ERR_INVALID;
    if (c->ipc_world == 0 || c->ipc_rank != 0) return fail(c, MRT_ERR_STATE, "img_gathered: not the film rank of an attached gather");
    return resize_and_copy(c, rgb);
}

int mrt_img_ss(mrt_ctx* c, uint8_t* rgb) {
    if (!c || !rgb) return MRT_ERR_INVALID;
    if (int rc = tonemap_ss(c)) return rc;
    mrt_ctx* f = film_ctx(c);
    CK(cudaSetDevice(f->device));
    CK(cudaMemcpyAsync(rgb, f->d_ss.p, (size_t)f->nw * f->nh * 3, cudaMemcpyDeviceToHost, f->stream));
    CK(cudaStreamSynchronize(f->stream));
    return harvest(c, false);
}

int mrt_img(mrt_ctx* c, uint8_t* rgb) {
    if (!c || !rgb) return MRT_ERR_INVALID;
    if (int rc = tonemap_ss(c)) return rc;
    mrt_ctx* f = film_ctx(c);
    const int rc = resize_and_copy(f, rgb);
    if (rc) { if (f != c) c->err = f->err; return rc; }
    return harvest(c, false);
}

int mrt_trace_primary(mrt_ctx* c, mrt_hit* out) {
    if (!c || !out) return MRT_ERR_INVALID;
    if (int rc = ready_to_render(c, "trace_primary")) return rc;
    if (is_group(c)) FWD(mrt_trace_primary(m, out));
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
    if (spp) {
        c->spp_per_launch = spp;
        for (mrt_ctx* m : c->members) m->spp_per_launch = spp;
    }
    if (current) *current = c->spp_per_launch;
    return MRT_OK;
}

int mrt_jit_status(mrt_ctx* c, uint32_t* eligible, uint32_t* compiled, uint64_t* launches, double* compile_seconds) {
    if (!c) return MRT_ERR_INVALID;
    if (is_group(c)) {
        mrt_ctx* f = film_ctx(c);
        const int rc = mrt_jit_status(f, eligible, compiled, nullptr, compile_seconds);
        if (!f->jit_err.empty()) c->err = f->err;
        if (launches) { *launches = 0; for (mrt_ctx* m : c->members) *launches += m->jit_launches; }
        return rc;
    }
    if (compile_seconds && c->jit_from_disk) *compile_seconds = -c->jit_seconds;  // negative: loaded from the disk cache
    if (eligible) *eligible = c->jit_header.empty() ? 0u : 1u;
    if (compiled) *compiled = c->jit_kernel ? 1u : 0u;
    if (launches) *launches = c->jit_launches;
    if (compile_seconds && !c->jit_from_disk) *compile_seconds = c->jit_seconds;
    if (!c->jit_err.empty()) c->err = c->jit_err;  // readable through mrt_last_error
    return MRT_OK;
}

int mrt_scene_info(mrt_ctx* c, uint32_t* scene_bvh, uint32_t* specialised, uint32_t* features) {
    if (!c) return MRT_ERR_INVALID;
    if (!c->have_scene) return fail(c, MRT_ERR_STATE, "scene_info before set_scene");
    const mrt_ctx* f = film_ctx(c);
    if (scene_bvh) *scene_bvh = f->gscene.bvh ? 1u : 0u;
    if (specialised) *specialised = f->jit_kernel ? 1u : 0u;
    if (features) *features = f->features;
    return MRT_OK;
}

int mrt_launch_count(mrt_ctx* c, uint64_t* n) {
    if (!c || !n) return MRT_ERR_INVALID;
    *n = c->launches;
    for (mrt_ctx* m : c->members) *n += m->launches;
    return MRT_OK;
}

int mrt_fp32_peak(mrt_ctx* c, double* tflops, double* seconds) {
    if (!c || !tflops) return MRT_ERR_INVALID;
    if (is_group(c)) FWD(mrt_fp32_peak(m, tflops, seconds));
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
