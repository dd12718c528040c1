// microbench.cu — issue-rate / pipe-throughput probes that back the kernel design in DESIGN.md.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/microbench tools/microbench.cu
// Prints, per variant, warp-instructions per cycle per SM sub-partition and the FP32 TFLOP/s.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t pk(float x, float y) { uint64_t r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(x), "f"(y)); return r; }
__device__ __forceinline__ void up(uint64_t v, float& x, float& y) { asm("mov.b64 {%0,%1}, %2;" : "=f"(x), "=f"(y) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d; asm("fma.rn.ftz.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d;
}
struct Consts { float a, b, c, d; };

// V: 0 FFMA reg | 1 FFMA2 reg | 2 FFMA2 uniform operands | 3 FFMA + FMNMX 1:1 | 4 FFMA2 + FMNMX 1:1
// 5 FFMA2 + 2 FMNMX | 6 FMNMX only | 7 FFMA uniform operands | 8 FFMA:FMNMX 2:1 | 9 FSETP+FSEL pairs
template <int V>
__global__ void __launch_bounds__(256) k(float* out, int iters, float ra, float rb, const __grid_constant__ Consts cs) {
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = (float)(threadIdx.x + i) * 1e-3f;
    uint64_t w[8];
#pragma unroll
    for (int i = 0; i < 8; i++) w[i] = pk(v[2 * i], v[2 * i + 1]);
    const uint64_t a2 = pk(ra, ra), b2 = pk(rb, rb);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
            if constexpr (V == 0) {
#pragma unroll
                for (int i = 0; i < 16; i++) v[i] = fmaf(v[i], ra, rb);
            } else if constexpr (V == 7) {
#pragma unroll
                for (int i = 0; i < 16; i++) v[i] = fmaf(v[i], cs.a, cs.b);
            } else if constexpr (V == 1) {
#pragma unroll
                for (int i = 0; i < 8; i++) w[i] = fma2(w[i], a2, b2);
            } else if constexpr (V == 2) {
#pragma unroll
                for (int i = 0; i < 8; i++) w[i] = fma2(w[i], pk(cs.a, cs.b), pk(cs.c, cs.d));
            } else if constexpr (V == 3) {
#pragma unroll
                for (int i = 0; i < 8; i++) { v[i] = fmaf(v[i], ra, rb); v[8 + i] = (r & 1) ? fmaxf(v[8 + i], v[i]) : fminf(v[8 + i], v[i]); }
            } else if constexpr (V == 8) {
#pragma unroll
                for (int i = 0; i < 5; i++) { v[i] = fmaf(v[i], ra, rb); v[5 + i] = fmaf(v[5 + i], ra, rb); v[10 + i] = (r & 1) ? fmaxf(v[10 + i], v[i]) : fminf(v[10 + i], v[i]); }
            } else if constexpr (V == 4) {
#pragma unroll
                for (int i = 0; i < 4; i++) { w[i] = fma2(w[i], a2, b2); float x, y; up(w[i], x, y); v[8 + i] = (r & 1) ? fmaxf(v[8 + i], x) : fminf(v[8 + i], x); }
            } else if constexpr (V == 5) {
#pragma unroll
                for (int i = 0; i < 4; i++) { w[i] = fma2(w[i], a2, b2); float x, y; up(w[i], x, y); v[8 + i] = (r & 1) ? fmaxf(v[8 + i], x) : fminf(v[8 + i], x); v[12 + i] = (r & 1) ? fminf(v[12 + i], y) : fmaxf(v[12 + i], y); }
            } else if constexpr (V == 6) {
#pragma unroll
                for (int i = 0; i < 16; i++) v[i] = (r & 1) ? fmaxf(v[i], v[(i + 1) & 15]) : fminf(v[i], v[(i + 1) & 15]);
            } else if constexpr (V == 9) {
#pragma unroll
                for (int i = 0; i < 8; i++) v[i] = (v[i] < v[8 + i]) ? v[i] : ra;
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; i++) s += v[i];
#pragma unroll
    for (int i = 0; i < 8; i++) { float x, y; up(w[i], x, y); s += x + y; }
    if (s == 12345.678f) out[0] = s;
}

template <int V>
void run(const char* name, double inst_per_iter, double fma_per_iter, float* d, int sms, double mhz) {
    const int blocks = sms * 8, iters = 8192;
    Consts cs{0.999f, 0.001f, 0.998f, 0.002f};
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<V><<<blocks, 256>>>(d, 64, 0.999f, 0.001f, cs);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0);
        k<V><<<blocks, 256>>>(d, iters, 0.999f, 0.001f, cs);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double warps = (double)blocks * 8, sec = best * 1e-3;
    const double winst = warps * iters * inst_per_iter;            // warp-instructions
    const double cyc = sec * mhz * 1e6;
    printf("%-28s %8.3f ms  %6.3f warp-inst/clk/SMSP (at %.0f MHz)  %7.2f TFLOP/s\n", name, best,
           winst / (cyc * sms * 4), mhz, warps * 32 * iters * fma_per_iter * 2 / sec / 1e12);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double mhz = khz / 1000.0;
    printf("%s, %d SMs, clock attr %.0f MHz\n", p.name, p.multiProcessorCount, mhz);
    float* d; cudaMalloc(&d, 64);
    const int sms = p.multiProcessorCount;
    run<0>("FFMA reg", 64, 64, d, sms, mhz);
    run<7>("FFMA const operands", 64, 64, d, sms, mhz);
    run<1>("FFMA2 reg", 32, 64, d, sms, mhz);
    run<2>("FFMA2 const operands", 32, 64, d, sms, mhz);
    run<6>("FMNMX only", 64, 0, d, sms, mhz);
    run<3>("FFMA:FMNMX 1:1", 64, 32, d, sms, mhz);
    run<8>("FFMA:FMNMX 2:1", 60, 40, d, sms, mhz);
    run<4>("FFMA2:FMNMX 1:1", 32, 32, d, sms, mhz);
    run<5>("FFMA2:FMNMX 1:2", 48, 32, d, sms, mhz);
    run<9>("FSETP+FSEL", 64, 0, d, sms, mhz);
    printf("cuda status: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
