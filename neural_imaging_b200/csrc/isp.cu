// Element-wise tails of the classic ISP model (reference models/pipelines.py:415-446, models/layers.py:238-258):
//   gamma   : y = pow(clip(x, lo, hi), e), the clip being straight-through (stop_gradient(clip(x) - x) + x)
//   residual: y = x_bilinear - alpha * f (alpha: trainable scalar on the device), optional straight-through clip to [0, 1]
#include "ni_common.cuh"

namespace {
constexpr int kT = 256;

__global__ void gamma_clip_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, long long n, float lo, float hi, float e) {
    const long long i = (long long)blockIdx.x * kT + threadIdx.x;
    if (i < n) y[i] = powf(fminf(fmaxf(x[i], lo), hi), e);
}
__global__ void gamma_clip_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx, long long n,
                                      float lo, float hi, float e) {
    const long long i = (long long)blockIdx.x * kT + threadIdx.x;
    if (i < n) dx[i] = dy[i] * e * powf(fminf(fmaxf(x[i], lo), hi), e - 1.f);
}
__global__ void residual_alpha_fwd_kernel(const float* __restrict__ xb, const float* __restrict__ f, const float* __restrict__ alpha,
                                          float* __restrict__ y, long long n, int clip) {
    const long long i = (long long)blockIdx.x * kT + threadIdx.x;
    if (i < n) { const float v = xb[i] - (*alpha) * f[i]; y[i] = clip ? ni_clamp01(v) : v; }
}
// df = -alpha * dy ; dalpha -= sum dy * f
__global__ void residual_alpha_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ f, const float* __restrict__ alpha,
                                          float* __restrict__ df, float* __restrict__ dalpha, long long n) {
    const float a = *alpha;
    float s = 0.f;
    for (long long i = (long long)blockIdx.x * kT + threadIdx.x; i < n; i += (long long)gridDim.x * kT) {
        const float g = dy[i];
        df[i] = -a * g;
        s -= g * f[i];
    }
    __shared__ float red[kT / 32];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0 && dalpha) {
        float t = 0.f;
        for (int k = 0; k < kT / 32; ++k) t += red[k];
        atomicAdd(dalpha, t);
    }
}
}  // namespace

extern "C" int ni_gamma_clip_fwd(const float* x, float* y, long long n, float lo, float hi, float exponent, cudaStream_t st) {
    NI_REQUIRE(x && y && n >= 0 && lo <= hi, "ni_gamma_clip_fwd: invalid arguments");
    if (n == 0) return NI_OK;
    gamma_clip_fwd_kernel<<<ni_cdiv(n, kT), kT, 0, st>>>(x, y, n, lo, hi, exponent);
    NI_LAUNCH_CHECK(); NI_COUNT_LAUNCH(1);
    return NI_OK;
}
extern "C" int ni_gamma_clip_bwd(const float* x, const float* dy, float* dx, long long n, float lo, float hi, float exponent, cudaStream_t st) {
    NI_REQUIRE(x && dy && dx && n >= 0 && lo <= hi, "ni_gamma_clip_bwd: invalid arguments");
    if (n == 0) return NI_OK;
    gamma_clip_bwd_kernel<<<ni_cdiv(n, kT), kT, 0, st>>>(x, dy, dx, n, lo, hi, exponent);
    NI_LAUNCH_CHECK(); NI_COUNT_LAUNCH(1);
    return NI_OK;
}
extern "C" int ni_residual_alpha_fwd(const float* xb, const float* f, const float* alpha, float* y, long long n, int clip, cudaStream_t st) {
    NI_REQUIRE(xb && f && alpha && y && n >= 0, "ni_residual_alpha_fwd: invalid arguments");
    if (n == 0) return NI_OK;
    residual_alpha_fwd_kernel<<<ni_cdiv(n, kT), kT, 0, st>>>(xb, f, alpha, y, n, clip);
    NI_LAUNCH_CHECK(); NI_COUNT_LAUNCH(1);
    return NI_OK;
}
// dalpha: one float, ACCUMULATED into (zero it first), may be NULL
extern "C" int ni_residual_alpha_bwd(const float* dy, const float* f, const float* alpha, float* df, float* dalpha, long long n, cudaStream_t st) {
    NI_REQUIRE(dy && f && alpha && df && n >= 0, "ni_residual_alpha_bwd: invalid arguments");
    if (n == 0) return NI_OK;
    long long grid = ni_cdiv(n, kT * 4);
    if (grid > 8LL * ni_num_sms()) grid = 8LL * ni_num_sms();
    residual_alpha_bwd_kernel<<<(int)grid, kT, 0, st>>>(dy, f, alpha, df, dalpha, n);
    NI_LAUNCH_CHECK(); NI_COUNT_LAUNCH(1);
    return NI_OK;
}
