// Micro-benchmark: is the packed FP32 instruction (FFMA2 / fma.rn.f32x2, sm_100a) issued at the scalar FFMA rate?
// Prints warp-instructions per clock per SM for scalar FFMA, FFMA2, and a 1:1 mix of FFMA2 with LDS.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }

template <int MODE> __global__ void __launch_bounds__(256) k(float* out, int iters, float seed) {
    __shared__ float sm[1024];
    sm[threadIdx.x] = seed; sm[threadIdx.x + 256] = seed; __syncthreads();
    float a[8]; u64 p[8];
    for (int i = 0; i < 8; ++i) { a[i] = seed + i + threadIdx.x; float2 t = make_float2(a[i], a[i] + 1.f); p[i] = *reinterpret_cast<u64*>(&t); }
    float2 m2 = make_float2(seed, seed); const u64 mm = *reinterpret_cast<u64*>(&m2);
    float acc = 0.f;
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) a[i] = fma1(a[i], seed, seed);
        } else if (MODE == 1) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) p[i] = fma2(p[i], mm, mm);
        } else if (MODE == 2) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) { p[i] = fma2(p[i], mm, mm); acc += sm[(threadIdx.x + i * 32 + r) & 1023]; }
        } else {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) { a[i] = fma1(a[i], seed, seed); acc += sm[(threadIdx.x + i * 32 + r) & 1023]; }
        }
    }
    float s = acc;
    for (int i = 0; i < 8; ++i) { float2 t = *reinterpret_cast<float2*>(&p[i]); s += a[i] + t.x + t.y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char* name, float* out, int sms, double ghz, int fp_per_iter, int all_per_iter) {
    const int iters = 20000, blocks = sms * 8;
    k<MODE><<<blocks, 256>>>(out, 100, 0.5f);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); k<MODE><<<blocks, 256>>>(out, iters, 0.5f); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double warps = (double)blocks * 8, clk = ms * 1e-3 * ghz * 1e9;
    printf("%-28s %8.3f ms  fp-instr/clk/SM %.2f  all-instr/clk/SM %.2f (at %.3f GHz nominal)\n", name, ms, warps * iters * fp_per_iter / clk / sms, warps * iters * all_per_iter / clk / sms, ghz);
}
int main() {
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    float* out; cudaMalloc(&out, pr.multiProcessorCount * 8 * 256 * 4);
    const double ghz = pr.clockRate * 1e-6;
    run<0>("scalar FFMA", out, pr.multiProcessorCount, ghz, 32, 32);
    run<1>("packed FFMA2", out, pr.multiProcessorCount, ghz, 32, 32);
    run<3>("FFMA + LDS + FADD", out, pr.multiProcessorCount, ghz, 32, 96);
    run<2>("FFMA2 + LDS + FADD", out, pr.multiProcessorCount, ghz, 32, 96);
    return 0;
}
