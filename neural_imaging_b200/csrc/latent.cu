// Fused discrete-latent kernels of the learned codec (TwitterDCN): scaling -> soft-codebook quantisation -> soft
// histogram for the differentiable entropy, forward and backward, in float64 like the reference.
//
// Replaces DiscreteLatent.call (models/layers.py:195-203), Quantization('soft-codebook') (models/layers.py:139-170) and
// tf_helpers.entropy (helpers/tf_helpers.py:290-333). The reference materialises the (n_values x 32) float64 weight
// matrix TWICE (2.7 GB each at M = 1280); here the weights of one value live in registers and only the 32-bin
// histogram leaves the kernel.
//   w_k   = (1 + (gamma (v - c_k))^2 / nu)^(-(nu+1)/2)           (t-Student, nu > 0)   or   exp(-gamma (v - c_k)^2)  (nu <= 0)
//   wn_k  = (w_k + 1e-72) / sum_j (w_j + 1e-72)
//   soft  = sum_k wn_k c_k ;  hard = c[argmax_k wn_k] ;  out = float(hard - soft) + float(soft)      (straight-through)
//   hist_k = mean_i wn_ik ;  h = clip(hist, 1e-9) / sum clip ;  H = -sum h ln h / 0.6931
#include "ni_common.cuh"

namespace {

constexpr int kMaxCodes = 256;
constexpr int kT = 128;

struct LatentParams {
    long long n;
    int ncodes;
    double nu, gamma;
    int rounding;      // 0 'soft-codebook', 1 'sin', 2 'soft' (round forward, sine gradient), 3 'identity' (models/layers.py:118-170)
    int e2;            // nu + 1 when that is a small integer (toolbox default nu = 50 -> 51), else 0: see kernel_weight
    double c_dlog;     // -(nu + 1) * gamma / nu
};

LatentParams make_params(long long n, int ncodes, double nu, double gamma, int rounding) {
    const double e = nu + 1.0;
    const int e2 = (nu > 0 && e == (double)(int)e && e >= 2.0 && e <= 128.0) ? (int)e : 0;
    return LatentParams{n, ncodes, nu, gamma, rounding, e2, nu > 0 ? -(nu + 1.0) * gamma / nu : 0.0};
}

// scalar rounding modes of the Quantization layer on the (float32) scaled latent; period-1 reduction keeps the SFU argument small
__device__ __forceinline__ float scalar_round(float v, int mode) {
    if (mode == 3) return v;
    if (mode == 2) return rintf(v);
    const float f = v - rintf(v);
    return v - sinf(6.2831855f * f) * (1.f / 6.2831855f);
}
__device__ __forceinline__ double scalar_round_grad(float v, int mode) {
    if (mode == 3) return 1.0;
    const float f = v - rintf(v);
    return 1.0 - (double)cosf(6.2831855f * f);         // d/dv [v - sin(2 pi v) / 2 pi], also the straight-through gradient of 'soft'
}

// The weight of one (value, code) pair. float64 pow() is ~250 instructions and float64 divisions ~35 each; the kernels used to evaluate
// 96 (forward) / 128 (backward) weights per latent value with pow + 3 - 5 divisions each (27 ms per step of config 5). With nu + 1 an
// integer, base^(-(nu+1)/2) = rsqrt(base^(nu+1)): repeated squaring (~9 multiplications for 51) + one rsqrt, <= ~10 ulp from pow's
// result -- far inside the float64 path's 1e-9 tolerance. WANT_DLOG also returns d(ln w)/dv (one division).
template <bool WANT_DLOG>
__device__ __forceinline__ double kernel_weight(double diff, const LatentParams& p, double& dlog) {
    if (p.nu > 0) {
        const double g = p.gamma * diff;
        const double base = 1.0 + g * g / p.nu;
        if (WANT_DLOG) dlog = p.c_dlog * g / base;               // -(nu+1)/2 * (2 gamma g / nu) / base
        if (p.e2 > 0) {
            double r = 1.0, b = base;
            for (int e = p.e2; e; e >>= 1) { if (e & 1) r *= b; b *= b; }
            return rsqrt(r);          // base >= 1; r overflows to inf only where the true weight is < 1e-300: rsqrt(inf) = 0
        }
        return pow(base, -(p.nu + 1.0) / 2.0);
    }
    if (WANT_DLOG) dlog = -2.0 * p.gamma * diff;
    return exp(-p.gamma * diff * diff);
}

// hist_acc[k] += sum_i wn_ik (double atomics, one per block and bin)
__global__ void __launch_bounds__(kT)
latent_fwd_kernel(const float* __restrict__ x, const float* __restrict__ scale, const float* __restrict__ codebook, float* __restrict__ out,
                  double* __restrict__ hist_acc, LatentParams p) {
    __shared__ double sh[kMaxCodes];
    __shared__ float cb[kMaxCodes];
    for (int k = threadIdx.x; k < p.ncodes; k += kT) { sh[k] = 0.0; cb[k] = codebook[k]; }
    __syncthreads();
    const float sc = scale ? *scale : 1.f;
    // every lane of a warp runs the same number of iterations (the histogram step below is a warp-wide reduction): the tail is masked
    const long long stride = (long long)gridDim.x * kT;
    for (long long i0 = (long long)blockIdx.x * kT; i0 < p.n; i0 += stride) {
        const long long i = i0 + threadIdx.x;
        const bool live = i < p.n;
        const float vf = (live ? x[i] : 0.f) * sc;   // float32 multiply, as the reference (latent * scaling_factor)
        const double v = (double)vf;
        float q;
        if (p.rounding == 0) {
            double S = 0.0, soft = 0.0, best = -1.0;
            int arg = 0;
            for (int k = 0; k < p.ncodes; ++k) {
                double dl;
                const double w = kernel_weight<false>(v - (double)cb[k], p, dl) + 1e-72;
                S += w;
                soft += w * (double)cb[k];
                if (w > best) { best = w; arg = k; }
            }
            soft /= S;
            const float softf = (float)soft, hard = cb[arg];
            q = (hard - softf) + softf;
        } else {
            q = scalar_round(vf, p.rounding);
        }
        if (live) out[i] = q;
        if (hist_acc) {
            // the reference estimates the entropy of the QUANTISED latent (models/layers.py:200-201): weights at q
            const double vq = (double)q;
            double Sq = 0.0;
            for (int k = 0; k < p.ncodes; ++k) {
                double dl;
                Sq += kernel_weight<false>(vq - (double)cb[k], p, dl) + 1e-72;
            }
            const double inv_sq = 1.0 / Sq;
            for (int k = 0; k < p.ncodes; ++k) {
                double dl;
                double w = live ? (kernel_weight<false>(vq - (double)cb[k], p, dl) + 1e-72) * inv_sq : 0.0;
                // one shared-memory atomic per warp and bin instead of one per lane (32 lanes on the same address serialise)
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
                if ((threadIdx.x & 31) == 0) atomicAdd(&sh[k], w);
            }
        }
    }
    __syncthreads();
    if (hist_acc)
        for (int k = threadIdx.x; k < p.ncodes; k += kT) atomicAdd(hist_acc + k, sh[k]);
}

// dx_i = scale * dv_i, dv_i = (g_out_i + sum_k gh_k * dwn_k(q_i)/dq) * dsoft/dv ; dscale += sum_i dv_i * x_i
// (the quantised value q = stop_gradient(hard - soft) + soft has dq/dsoft = 1; the entropy is a function of q)
// gh_k = (entropy upstream) * dH/dhist_k / n   (computed by ni_entropy_from_hist from the global histogram)
__global__ void __launch_bounds__(kT)
latent_bwd_kernel(const float* __restrict__ x, const float* __restrict__ scale, const float* __restrict__ codebook,
                  const float* __restrict__ q, const float* __restrict__ g_out, const double* __restrict__ gh, float* __restrict__ dx, double* __restrict__ dscale_acc,
                  LatentParams p) {
    __shared__ float cb[kMaxCodes];
    __shared__ double sgh[kMaxCodes];
    __shared__ double red[kT / 32];
    for (int k = threadIdx.x; k < p.ncodes; k += kT) { cb[k] = codebook[k]; sgh[k] = gh ? gh[k] : 0.0; }
    __syncthreads();
    const float sc = scale ? *scale : 1.f;
    double ds = 0.0;
    for (long long i = (long long)blockIdx.x * kT + threadIdx.x; i < p.n; i += (long long)gridDim.x * kT) {
        const float xf = x[i];
        const double v = (double)(xf * sc);
        double dsoft = 0.0, dent = 0.0;
        if (p.rounding == 0) {
            // d soft / dv = sum_k c_k d(wn_k)/dv, wn_k = (w_k + eps) / S: with A = sum w a (a = dlnw/dv), B = sum c w a, C = sum c (w + eps)
            // this is B / S - C A / S^2 -- one pass over the code book
            double S = 0.0, A = 0.0, B = 0.0, C = 0.0;
            for (int k = 0; k < p.ncodes; ++k) {
                double dl;
                const double c = (double)cb[k];
                const double w = kernel_weight<true>(v - c, p, dl);
                S += w + 1e-72;
                A += w * dl;
                B += c * (w * dl);
                C += c * (w + 1e-72);
            }
            const double inv_s = 1.0 / S;
            dsoft = (B - C * A * inv_s) * inv_s;
        } else {
            dsoft = scalar_round_grad(xf * sc, p.rounding);      // dq/dv of the scalar rounding modes
        }
        if (gh) {
            const double vq = (double)q[i];
            double Sq = 0.0, Aq = 0.0, Bq = 0.0, Cq = 0.0;      // same identity with the histogram gradients gh_k in place of c_k
            for (int k = 0; k < p.ncodes; ++k) {
                double dl;
                const double w = kernel_weight<true>(vq - (double)cb[k], p, dl);
                Sq += w + 1e-72;
                Aq += w * dl;
                Bq += sgh[k] * (w * dl);
                Cq += sgh[k] * (w + 1e-72);
            }
            const double inv_sq = 1.0 / Sq;
            dent = (Bq - Cq * Aq * inv_sq) * inv_sq;
        }
        const double dv = ((g_out ? (double)g_out[i] : 0.0) + dent) * dsoft;
        dx[i] = (float)(dv * (double)sc);
        ds += dv * (double)xf;
    }
    if (dscale_acc) {
        for (int o = 16; o > 0; o >>= 1) ds += __shfl_xor_sync(0xffffffffu, ds, o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ds;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int k = 0; k < kT / 32; ++k) t += red[k];
            atomicAdd(dscale_acc, t);
        }
    }
}

// y = act(x) elementwise (leaky-relu on a residual-branch input, models/compression.py:224) and its backward
__global__ void lrelu_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, long long n, float alpha) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i < n) { const float v = x[i]; y[i] = v > 0.f ? v : alpha * v; }
}
// dx (+)= dy * act'(x)
__global__ void lrelu_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx, long long n, float alpha,
                                 int accumulate) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i < n) { const float g = dy[i] * (x[i] > 0.f ? 1.f : alpha); dx[i] = accumulate ? dx[i] + g : g; }
}

}  // namespace

// out: quantised latent (n floats). hist_acc: 32 (ncodes) doubles, zeroed by the caller, receives sum_i wn_ik (may be NULL).
// scale: device pointer to the trainable scaling factor (NULL = 1). codebook: ncodes floats on the device.
// rounding: 0 'soft-codebook', 1 'sin', 2 'soft', 3 'identity' — the modes models/compression.py:66 accepts for the latent quantiser; the
// entropy estimate (soft histogram of the QUANTISED latent over the code book) is the same for all of them.
extern "C" int ni_latent_quantise_fwd(const float* x, const float* scale, const float* codebook, float* out, double* hist_acc,
                                      long long n, int ncodes, double nu, double gamma, int rounding, cudaStream_t st) {
    NI_REQUIRE(x && codebook && out && n >= 0 && ncodes > 1 && ncodes <= kMaxCodes && rounding >= 0 && rounding <= 3,
               "ni_latent_quantise_fwd: invalid arguments");
    if (n == 0) return NI_OK;
    const LatentParams p = make_params(n, ncodes, nu, gamma, rounding);
    int grid = ni_cdiv(n, kT * 4);
    if (grid > 16 * ni_num_sms()) grid = 16 * ni_num_sms();
    latent_fwd_kernel<<<grid, kT, 0, st>>>(x, scale, codebook, out, hist_acc, p);
    NI_LAUNCH_CHECK(); NI_COUNT_LAUNCH(1);
    return NI_OK;
}

// q: the quantised latent written by the forward (needed when gh != NULL). g_out: gradient w.r.t. the quantised latent (NULL = 0). gh: ncodes doubles = d(loss)/d(hist_k) / n (NULL = no entropy term).
// dscale_acc: one double, zeroed by the caller, receives d(loss)/d(scale) (may be NULL).
extern "C" int ni_latent_softcodebook_fwd(const float* x, const float* scale, const float* codebook, float* out, double* hist_acc,
                                          long long n, int ncodes, double nu, double gamma, cudaStream_t st) {
    return ni_latent_quantise_fwd(x, scale, codebook, out, hist_acc, n, ncodes, nu, gamma, 0, st);
}
extern "C" int ni_latent_quantise_bwd(const float* x, const float* scale, const float* codebook, const float* q, const float* g_out,
                                      const double* gh, float* dx, double* dscale_acc, long long n, int ncodes, double nu, double gamma,
                                      int rounding, cudaStream_t st) {
    NI_REQUIRE(x && codebook && dx && (q || !gh) && n >= 0 && ncodes > 1 && ncodes <= kMaxCodes && rounding >= 0 && rounding <= 3,
               "ni_latent_quantise_bwd: invalid arguments");
    if (n == 0) return NI_OK;
    const LatentParams p = make_params(n, ncodes, nu, gamma, rounding);
    int grid = ni_cdiv(n, kT * 4);
    if (grid > 16 * ni_num_sms()) grid = 16 * ni_num_sms();
    latent_bwd_kernel<<<grid, kT, 0, st>>>(x, scale, codebook, q, g_out, gh, dx, dscale_acc, p);
    NI_LAUNCH_CHECK(); NI_COUNT_LAUNCH(1);
    return NI_OK;
}

extern "C" int ni_latent_softcodebook_bwd(const float* x, const float* scale, const float* codebook, const float* q, const float* g_out,
                                          const double* gh, float* dx, double* dscale_acc, long long n, int ncodes, double nu, double gamma, cudaStream_t st) {
    return ni_latent_quantise_bwd(x, scale, codebook, q, g_out, gh, dx, dscale_acc, n, ncodes, nu, gamma, 0, st);
}
extern "C" int ni_leaky_relu_fwd(const float* x, float* y, long long n, float alpha, cudaStream_t st) {
    NI_REQUIRE(x && y && n >= 0, "ni_leaky_relu_fwd: invalid arguments");
    if (n == 0) return NI_OK;
    lrelu_fwd_kernel<<<ni_cdiv(n, 256), 256, 0, st>>>(x, y, n, alpha);
    NI_LAUNCH_CHECK(); NI_COUNT_LAUNCH(1);
    return NI_OK;
}
extern "C" int ni_leaky_relu_bwd(const float* x, const float* dy, float* dx, long long n, float alpha, int accumulate, cudaStream_t st) {
    NI_REQUIRE(x && dy && dx && n >= 0, "ni_leaky_relu_bwd: invalid arguments");
    if (n == 0) return NI_OK;
    lrelu_bwd_kernel<<<ni_cdiv(n, 256), 256, 0, st>>>(x, dy, dx, n, alpha, accumulate);
    NI_LAUNCH_CHECK(); NI_COUNT_LAUNCH(1);
    return NI_OK;
}

namespace {
// Single-block epilogue of the entropy estimate: H (float) and gh_k = upstream * dH/dhist_k / n from the accumulated histogram.
__global__ void entropy_from_hist_kernel(const double* __restrict__ hist_acc, double n, int ncodes, double upstream, float* __restrict__ H_out,
                                         double* __restrict__ gh_out) {
    __shared__ double hc[256], dHdp[256];
    __shared__ double T, dot;
    const int k = threadIdx.x;
    if (k < ncodes) hc[k] = fmax(hist_acc[k] / n, 1e-9);
    __syncthreads();
    if (k == 0) { double t = 0; for (int j = 0; j < ncodes; ++j) t += hc[j]; T = t; }
    __syncthreads();
    if (k < ncodes) { const double p = hc[k] / T; dHdp[k] = -(log(p) + 1.0) / 0.6931; }
    __syncthreads();
    if (k == 0) {
        double h = 0, d = 0;
        for (int j = 0; j < ncodes; ++j) { const double p = hc[j] / T; h -= p * log(p); d += p * dHdp[j]; }
        dot = d;
        if (H_out) *H_out = (float)(h / 0.6931);
    }
    __syncthreads();
    if (k < ncodes && gh_out) {
        const double pass = (hist_acc[k] / n) >= 1e-9 ? 1.0 : 0.0;     // clip_by_value passes the gradient inside the range
        gh_out[k] = upstream * pass * (dHdp[k] - dot) / T / n;
    }
}
}  // namespace

// hist_acc: ncodes doubles (sum over the n values of the normalised weights). H_out: float entropy estimate (may be NULL).
// gh_out: ncodes doubles = upstream * dH/dhist_k / n for ni_latent_softcodebook_bwd (may be NULL).
extern "C" int ni_entropy_from_hist(const double* hist_acc, long long n, int ncodes, double upstream, float* h_out, double* gh_out,
                                    cudaStream_t st) {
    NI_REQUIRE(hist_acc && n > 0 && ncodes > 1 && ncodes <= 256, "ni_entropy_from_hist: invalid arguments");
    entropy_from_hist_kernel<<<1, 256, 0, st>>>(hist_acc, (double)n, ncodes, upstream, h_out, gh_out);
    NI_LAUNCH_CHECK(); NI_COUNT_LAUNCH(1);
    return NI_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// Quantization layer, scalar modes (models/layers.py:118-136): 0 'round', 1 'sin', 2 'soft' (forward value = round, the sine only
// shapes the gradient), 3 'harmonic' (taylor_terms terms), 4 'identity'. The differentiable JPEG fuses these into its own kernel;
// this entry point serves the stand-alone layer (models.layers.Quantization).
namespace {
__global__ void quantize_scalar_kernel(const float* __restrict__ x, float* __restrict__ y, long long n, int mode, int taylor_terms) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float v = x[i];
        float r;
        // period-1 argument reduction, exact in FP32: sin(2 pi k v) == sin(2 pi k (v - rint(v))) for integer k
        const float f = v - rintf(v);
        const float two_pi = 6.283185307179586f, pi = 3.141592653589793f;
        if (mode == 0 || mode == 2) r = rintf(v);
        else if (mode == 1) r = v - sinf(two_pi * f) / two_pi;
        else if (mode == 3) {
            r = v - sinf(two_pi * f) / pi;
            for (int k = 2; k < taylor_terms; ++k) r += ((k & 1) ? -1.f : 1.f) * sinf(two_pi * (float)k * f) / ((float)k * pi);
        } else r = v;
        y[i] = r;
    }
}
}  // namespace

extern "C" int ni_quantize_scalar(const float* x, float* y, long long n, int mode, int taylor_terms, cudaStream_t st) {
    NI_REQUIRE(x && y && n >= 0 && mode >= 0 && mode <= 4 && taylor_terms >= 1, "ni_quantize_scalar: invalid arguments");
    if (n == 0) return NI_OK;
    long long blocks = (n + 255) / 256;
    const long long cap = 16LL * ni_num_sms();
    quantize_scalar_kernel<<<(unsigned)(blocks > cap ? cap : blocks), 256, 0, st>>>(x, y, n, mode, taylor_terms);
    NI_LAUNCH_CHECK();
    NI_COUNT_LAUNCH(1);
    return NI_OK;
}
