// Throughput probe: 3-register scalar FFMA vs packed FFMA2 (fma.rn.f32x2, sm_100a) with register operands, with and without a
// shared-memory (broadcast LDS.128) operand stream. tools/probes/ffma2_probe.cu measured the constant-operand forms only.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probes/_ffma_forms_probe tools/probes/ffma_forms_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>   // 0 scalar FFMA, 1 FFMA2, 2 scalar + 2 LDS.128 per 32 FMA, 3 FFMA2 + 2 LDS.128 per 32 FMA, 4 FFMA2 + 2 LDS.128 per 64 FMA
__global__ void __launch_bounds__(256) probe(float* out, const float* in, int iters, long long* clk) {
    __shared__ float4 sh[256];
    sh[threadIdx.x] = make_float4(in[threadIdx.x], in[threadIdx.x + 1], in[threadIdx.x + 2], in[threadIdx.x + 3]);
    __syncthreads();
    float2 acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = make_float2(in[i], in[i + 1]);
    float2 a = make_float2(in[threadIdx.x & 7], in[(threadIdx.x & 7) + 1]);
    float2 b = make_float2(in[32], in[33]);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE >= 2) {
            const float4 w0 = sh[(it * 2) & 255];                 // warp-uniform address: broadcast
            const float4 w1 = sh[(it * 2 + 1) & 255];
            if (MODE == 2) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    acc[4 * i].x = fmaf(a.x, w0.x, acc[4 * i].x); acc[4 * i].y = fmaf(a.y, w0.y, acc[4 * i].y);
                    acc[4 * i + 1].x = fmaf(a.x, w0.z, acc[4 * i + 1].x); acc[4 * i + 1].y = fmaf(a.y, w0.w, acc[4 * i + 1].y);
                    acc[4 * i + 2].x = fmaf(a.x, w1.x, acc[4 * i + 2].x); acc[4 * i + 2].y = fmaf(a.y, w1.y, acc[4 * i + 2].y);
                    acc[4 * i + 3].x = fmaf(a.x, w1.z, acc[4 * i + 3].x); acc[4 * i + 3].y = fmaf(a.y, w1.w, acc[4 * i + 3].y);
                }
            } else {
                const float2 p0 = make_float2(w0.x, w0.y), p1 = make_float2(w0.z, w0.w), p2 = make_float2(w1.x, w1.y), p3 = make_float2(w1.z, w1.w);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    acc[4 * i] = __ffma2_rn(a, p0, acc[4 * i]); acc[4 * i + 1] = __ffma2_rn(a, p1, acc[4 * i + 1]);
                    acc[4 * i + 2] = __ffma2_rn(a, p2, acc[4 * i + 2]); acc[4 * i + 3] = __ffma2_rn(a, p3, acc[4 * i + 3]);
                }
                if (MODE == 4) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        acc[4 * i] = __ffma2_rn(b, p1, acc[4 * i]); acc[4 * i + 1] = __ffma2_rn(b, p2, acc[4 * i + 1]);
                        acc[4 * i + 2] = __ffma2_rn(b, p3, acc[4 * i + 2]); acc[4 * i + 3] = __ffma2_rn(b, p0, acc[4 * i + 3]);
                    }
                }
            }
        } else if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) { acc[i].x = fmaf(a.x, b.x, acc[i].x); acc[i].y = fmaf(a.y, b.y, acc[i].y); }
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = __ffma2_rn(a, b, acc[i]);
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int ctas_per_sm, int fma_per_iter) {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float *out, *in;
    long long* clk;
    const int grid = sms * ctas_per_sm, iters = 20000;
    cudaMalloc(&out, grid * 256 * 4); cudaMalloc(&in, 1024 * 4); cudaMalloc(&clk, grid * 8);
    cudaMemset(in, 0, 1024 * 4);
    probe<MODE><<<grid, 256>>>(out, in, 100, clk);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<MODE><<<grid, 256>>>(out, in, iters, clk);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    static long long h[8192]; cudaMemcpy(h, clk, grid * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < grid; ++i) avg += h[i]; avg /= grid;
    const double lanes = (double)fma_per_iter * iters * 256 * ctas_per_sm / avg;
    printf("%-40s %d CTAs/SM (%2d warps/SMSP): %6.1f FMA lanes/clk/SM, %.2f ms, %5.1f TFLOP/s, %s\n", name, ctas_per_sm, ctas_per_sm * 2, lanes, ms,
           2.0 * fma_per_iter * iters * 256.0 * grid / ms * 1e-9, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(in); cudaFree(clk);
}

int main() {
    for (int c : {1, 2, 4, 8}) {
        run<0>("scalar FFMA (3 registers)", c, 32);
        run<1>("FFMA2 (3 register pairs)", c, 32);
        run<2>("scalar FFMA + 2 LDS.128 / 32 FMA", c, 32);
        run<3>("FFMA2 + 2 LDS.128 / 32 FMA", c, 32);
        run<4>("FFMA2 + 2 LDS.128 / 64 FMA", c, 64);
    }
    return 0;
}
