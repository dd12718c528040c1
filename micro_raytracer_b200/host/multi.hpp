// multi.hpp — one render over several GPUs of a box from a single process:
// sample split + one NCCL reduce of the accumulators.  See multi.cpp.
#pragma once
#include <memory>
#include <vector>

#include "render.hpp"

namespace mrt_host {

// ≙ one `Sampler` whose passes are rendered by `devices.size()` GPUs.  GPU g of G renders the global
// sample indices g, g+G, ... of every supersampled pixel (the counter-based RNG makes the image
// independent of G up to the f32 summation order); the per-GPU accumulators are then summed onto
// the first device by ONE ncclReduce over NVLink — the exchange step that replaces the reference's
// Mutex<HashMap> merge (sampler.rs:60-70) — and the film is read out there.
class MultiSampler {
public:
    MultiSampler(const std::vector<int>& devices, uint32_t workers = 24, uint32_t n_dim = 64, uint64_t seed = 0x5EED);
    ~MultiSampler();
    MultiSampler(const MultiSampler&) = delete;
    MultiSampler& operator=(const MultiSampler&) = delete;

    // Renders `n_passes` passes in total (a fresh film each call) and reduces; returns wall seconds.
    double execute(const Scene& scene, const Frame& frame, const RayTracer& rt, uint32_t n_passes);
    Image img(const Frame& frame);
    size_t world() const { return samplers_.size(); }

private:
    std::vector<std::unique_ptr<Sampler>> samplers_;
    std::vector<void*> comms_;  // ncclComm_t per device
};

int cuda_device_count();  // via libmrt.so; 0 without a driver

}  // namespace mrt_host
