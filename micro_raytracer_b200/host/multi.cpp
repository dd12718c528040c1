// multi.cpp — multi-GPU rendering from one process: one context per device, sample split
// (mrt_set_partition), asynchronous launches on every device, then a single ncclReduce(sum) of the
// float4 accumulators to the first device inside one NCCL group, ordered on each context's own
// stream (mrt_accum_device) so no host synchronisation sits between the last launch and the
// reduce.  NCCL is dlopen'ed (libnccl.so.2): the binary starts without it and a one-GPU render
// never needs it.  SURVEY.md §8(e); the torchrun / one-process-per-GPU twin is
// micro_raytracer_b200/distributed.py.
#include "multi.hpp"

#include <dlfcn.h>

#include <chrono>
#include <string>

namespace mrt_host {

namespace {
// the handful of NCCL entry points, declared by hand so that no NCCL header is needed to build
typedef void* ncclComm_t;
typedef int ncclResult_t;
enum { kNcclSuccess = 0, kNcclFloat = 7, kNcclSum = 0 };
struct Nccl {
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Reduce)(const void*, void*, size_t, int, int, int, ncclComm_t, void*) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string why;
    bool ok = false;
};
Nccl& nccl() {
    static Nccl n;
    static bool tried = false;
    if (tried) return n;
    tried = true;
    void* h = nullptr;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
        h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) { n.why = std::string("cannot load libnccl.so.2: ") + dlerror(); return n; }
    auto sym = [&](const char* s) { void* p = dlsym(h, s); if (!p) n.why = std::string("missing NCCL symbol ") + s; return p; };
    n.CommInitAll = reinterpret_cast<decltype(n.CommInitAll)>(sym("ncclCommInitAll"));
    n.CommDestroy = reinterpret_cast<decltype(n.CommDestroy)>(sym("ncclCommDestroy"));
    n.GroupStart = reinterpret_cast<decltype(n.GroupStart)>(sym("ncclGroupStart"));
    n.GroupEnd = reinterpret_cast<decltype(n.GroupEnd)>(sym("ncclGroupEnd"));
    n.Reduce = reinterpret_cast<decltype(n.Reduce)>(sym("ncclReduce"));
    n.GetErrorString = reinterpret_cast<decltype(n.GetErrorString)>(sym("ncclGetErrorString"));
    n.ok = n.why.empty();
    return n;
}
void nccl_check(ncclResult_t r, const char* what) {
    if (r != kNcclSuccess) throw Error(std::string(what) + ": " + (nccl().GetErrorString ? nccl().GetErrorString(r) : "NCCL error"));
}
void check(mrt_ctx* c, int rc, const char* what) {
    if (rc) {
        const char* e = mrt_last_error(c);
        throw Error(e && *e ? std::string(e) : std::string(what) + " failed");
    }
}
}  // namespace

int cuda_device_count() {
    int n = 0;
    mrt_device_count(&n);
    return n;
}

MultiSampler::MultiSampler(const std::vector<int>& devices, uint32_t workers, uint32_t n_dim, uint64_t seed) {
    if (devices.empty()) throw Error("no CUDA device given");
    for (int d : devices) samplers_.push_back(std::make_unique<Sampler>(workers, n_dim, d, seed));
    if (devices.size() > 1) {
        Nccl& n = nccl();
        if (!n.ok) throw Error("multi-GPU rendering needs NCCL: " + n.why);
        comms_.resize(devices.size(), nullptr);
        nccl_check(n.CommInitAll(comms_.data(), (int)devices.size(), devices.data()), "ncclCommInitAll");
    }
}
MultiSampler::~MultiSampler() {
    for (void* c : comms_) if (c) nccl().CommDestroy(c);
}

double MultiSampler::execute(const Scene& scene, const Frame& frame, const RayTracer& rt, uint32_t n_passes) {
    const auto t0 = std::chrono::steady_clock::now();
    const uint32_t G = (uint32_t)samplers_.size();
    if (G == 1) {
        check(samplers_[0]->ctx(), mrt_reset(samplers_[0]->ctx()), "mrt_reset");
        samplers_[0]->execute(scene, frame, rt, n_passes);
        return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }
    for (uint32_t g = 0; g < G; g++) {  // queue every device's share without waiting
        Sampler& s = *samplers_[g];
        s.bind(scene, frame, rt);
        check(s.ctx(), mrt_reset(s.ctx()), "mrt_reset");
        check(s.ctx(), mrt_set_partition(s.ctx(), g, G), "mrt_set_partition");
        const uint32_t mine = n_passes / G + (g < n_passes % G ? 1u : 0u);  // |{g, g+G, ...} below n_passes|
        if (mine) check(s.ctx(), mrt_execute_async(s.ctx(), mine), "mrt_execute_async");
    }
    Nccl& n = nccl();
    nccl_check(n.GroupStart(), "ncclGroupStart");
    for (uint32_t g = 0; g < G; g++) {
        void* ptr = nullptr; size_t count = 0; void* stream = nullptr;
        check(samplers_[g]->ctx(), mrt_accum_device(samplers_[g]->ctx(), &ptr, &count, &stream), "mrt_accum_device");
        nccl_check(n.Reduce(ptr, ptr, count, kNcclFloat, kNcclSum, 0, comms_[g], stream), "ncclReduce");
    }
    nccl_check(n.GroupEnd(), "ncclGroupEnd");
    for (uint32_t g = 0; g < G; g++) check(samplers_[g]->ctx(), mrt_sync(samplers_[g]->ctx()), "mrt_sync");
    check(samplers_[0]->ctx(), mrt_set_passes(samplers_[0]->ctx(), n_passes), "mrt_set_passes");
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

Image MultiSampler::img(const Frame& frame) { return samplers_[0]->img(frame); }

}  // namespace mrt_host
