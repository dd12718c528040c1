// mrt_kernels.h — host-callable launchers of the kernels in mrt_kernels.cu.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/mrt.h"

#ifndef MRT_PATH_BLOCK
#define MRT_PATH_BLOCK 128
#endif
#ifndef MRT_PATH_MINBLOCKS
#define MRT_PATH_MINBLOCKS 0  // > 0: second __launch_bounds__ argument (forces the register budget)
#endif
#if MRT_PATH_MINBLOCKS > 0
#define MRT_PATH_BOUNDS __launch_bounds__(MRT_PATH_BLOCK, MRT_PATH_MINBLOCKS)
#else
#define MRT_PATH_BOUNDS __launch_bounds__(MRT_PATH_BLOCK)
#endif

// the accumulators of a device group's members (the first is the film device's own), passed by value
#define MRT_MAX_GROUP 16
struct PeerAccums { const float4* p[MRT_MAX_GROUP]; uint32_t n; };

struct ParamScene;
struct GlobalScene;
struct FilmParams;

cudaError_t mrt_launch_path(uint32_t features, bool in_param, const ParamScene* ps, const GlobalScene* gs,
                            const FilmParams& fp, cudaStream_t st);
cudaError_t mrt_launch_primary(const GlobalScene& gs, const FilmParams& fp, mrt_hit* out, const uint32_t* obj_inst, cudaStream_t st);
cudaError_t mrt_launch_tonemap(const float4* accum, uint8_t* out, uint32_t npix, float inv_n, float gamma, float exp, cudaStream_t st);
cudaError_t mrt_launch_unpack(const float4* accum, float* out, uint32_t npix, cudaStream_t st);
// device groups: tone-map pixels [first, first + count) of the SUM of the members' accumulators into `out` (the film
// device's u8 supersampled image, possibly peer memory); the sum as packed RGB; dst += src
cudaError_t mrt_launch_tonemap_peers(const PeerAccums& acc, uint8_t* out, uint32_t first, uint32_t count, float inv_n, float gamma, float exp, cudaStream_t st);
cudaError_t mrt_launch_unpack_peers(const PeerAccums& acc, float* out, uint32_t npix, cudaStream_t st);
cudaError_t mrt_launch_accum_add(float4* dst, const float4* src, uint32_t npix, cudaStream_t st);
cudaError_t mrt_launch_lanczos_weights(uint32_t in_n, uint32_t out_n, uint32_t max_taps, int32_t* left, int32_t* cnt, float* w, cudaStream_t st);
cudaError_t mrt_launch_lanczos_vertical(const uint8_t* src, float* tmp, uint32_t w, uint32_t nh, uint32_t max_taps,
                                        const int32_t* left, const int32_t* cnt, const float* wt, cudaStream_t st);
cudaError_t mrt_launch_lanczos_horizontal(const float* tmp, uint8_t* dst, uint32_t w, uint32_t nw, uint32_t nh, uint32_t max_taps,
                                          const int32_t* left, const int32_t* cnt, const float* wt, cudaStream_t st);
cudaError_t mrt_launch_fp32_peak(float* out, int blocks, int iters, cudaStream_t st);
