// mrt_jit.h — run-time (NVRTC) scene-specialised path kernel, see mrt_jit.cu.
#pragma once
#include <string>

#include <cuda_runtime.h>

#include "mrt_device.cuh"

// The two entry points of a specialised cubin: `pinhole` for frames with aperture 0 (every sample of a pixel starts with the
// same ray: no lens code at all), `thin` for thin-lens frames.
struct MrtJitKernels {
    cudaKernel_t thin = nullptr, pinhole = nullptr;
    cudaKernel_t pick(const FilmParams& fp, bool allow_pinhole) const { return allow_pinhole && fp.aprt == 0.0f ? pinhole : thin; }
};
struct MrtJitInfo { bool pending = false; bool from_disk = false; double seconds = 0.0; std::string err; };

// The kernel specialised for `scene_header` (the text mrt_api.cu generates: feature mask + instance
// tables as literal X-macro lists).  The first request starts an NVRTC compile on a background thread
// (or loads the cubin from the on-disk cache); wait_ms < 0 blocks until it is over, wait_ms > 0 waits at most that
// long (a cubin found in the on-disk cache is ready within ~3 ms), 0 only polls.  Returns nullptr
// while the compile is pending, when NVRTC is unavailable, or when the compile failed (info->err).
const MrtJitKernels* mrt_jit_kernel(const std::string& scene_header, int wait_ms, MrtJitInfo* info);
// Blocks until a compile started for `scene_header` (if any) is over.
void mrt_jit_wait(const std::string& scene_header);
cudaError_t mrt_jit_launch(const MrtJitKernels* k, const SceneCommon& scene, const FilmParams& fp, cudaStream_t st, bool allow_pinhole = true);
// the kernel of a BVH scene (header with MRT_JIT_BVH) takes the whole GlobalScene
cudaError_t mrt_jit_launch_bvh(const MrtJitKernels* k, const GlobalScene& scene, const FilmParams& fp, cudaStream_t st, bool allow_pinhole = true);
