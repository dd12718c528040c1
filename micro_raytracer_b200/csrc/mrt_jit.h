// mrt_jit.h — run-time (NVRTC) scene-specialised path kernel, see mrt_jit.cu.
#pragma once
#include <string>

#include <cuda_runtime.h>

#include "mrt_device.cuh"

// Compiles (or fetches from the process-wide cache) the kernel specialised for `scene_header`
// (the text mrt_api.cu generates: feature mask + instance tables as literal X-macro lists).
// Returns nullptr when NVRTC is unavailable or the compile failed (*err says why).
cudaKernel_t mrt_jit_kernel(const std::string& scene_header, double* compile_seconds, std::string* err);
cudaError_t mrt_jit_launch(cudaKernel_t k, const SceneCommon& scene, const FilmParams& fp, cudaStream_t st);
