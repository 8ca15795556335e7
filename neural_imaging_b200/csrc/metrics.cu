// Image-quality metrics on the device: structural similarity in ONE fused kernel.
//
// Replaces (a) tf.image.ssim(a, b, 1.0) as called by DCN.training_step / ssim_loss (models/compression.py:89,
// helpers/tf_helpers.py:39-40): 11 x 11 Gaussian window (sigma 1.5), VALID, k1 = 0.01, k2 = 0.03, and (b) the validation metric
// helpers/metrics.py:9-26 = skimage structural_similarity(multichannel=True, data_range=1): uniform 7 x 7 window, sample covariance
// (N / (N - 1)), border of 3 pixels cropped. Both are "separable window -> five local moments -> SSIM map -> mean over the VALID
// region and the channels"; they differ in the window and in the covariance normalisation only. The TensorFlow graph materialises
// five filtered copies of both images; here a tile's moments live in shared memory and only one float per image leaves the SM.
#include "ni_common.cuh"

namespace {

constexpr int kMaxWin = 15;
constexpr int TW = 32, TH = 16;

struct SsimParams {
    float win[kMaxWin];
    int k;
    float cov_norm, c1, c2, inv_count;
};

__global__ void __launch_bounds__(256) ssim_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, int h, int w,
                                                   int c, const SsimParams P) {
    extern __shared__ float sm[];
    const int k = P.k, HW = TW + k - 1, HH = TH + k - 1;
    float* sa = sm;                       // [HH][HW]
    float* sb = sa + HH * HW;             // [HH][HW]
    float* hm = sb + HH * HW;             // 5 x [HH][TW] horizontally filtered moments
    __shared__ float red[8];
    const int img = blockIdx.z / c, ch = blockIdx.z - img * c;
    const int ox0 = blockIdx.x * TW, oy0 = blockIdx.y * TH;
    const int ow = w - k + 1, oh = h - k + 1;
    const long long base = (long long)img * h * w * c + ch;
    for (int i = threadIdx.x; i < HH * HW; i += 256) {
        const int hy = i / HW, hx = i - hy * HW;
        const int y = oy0 + hy, x = ox0 + hx;
        float va = 0.f, vb = 0.f;
        if (y < h && x < w) {
            const long long o = base + ((long long)y * w + x) * c;
            va = __ldg(a + o); vb = __ldg(b + o);
        }
        sa[i] = va; sb[i] = vb;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < HH * TW; i += 256) {
        const int hy = i / TW, ox = i - hy * TW;
        float m0 = 0.f, m1 = 0.f, e00 = 0.f, e11 = 0.f, e01 = 0.f;
        for (int j = 0; j < k; ++j) {
            const float wj = P.win[j], x = sa[hy * HW + ox + j], y = sb[hy * HW + ox + j];
            m0 = fmaf(wj, x, m0); m1 = fmaf(wj, y, m1);
            e00 = fmaf(wj, x * x, e00); e11 = fmaf(wj, y * y, e11); e01 = fmaf(wj, x * y, e01);
        }
        hm[i] = m0; hm[HH * TW + i] = m1; hm[2 * HH * TW + i] = e00; hm[3 * HH * TW + i] = e11; hm[4 * HH * TW + i] = e01;
    }
    __syncthreads();
    float acc = 0.f;
    for (int i = threadIdx.x; i < TH * TW; i += 256) {
        const int oy = i / TW, ox = i - oy * TW;
        if (oy0 + oy >= oh || ox0 + ox >= ow) continue;
        float m0 = 0.f, m1 = 0.f, e00 = 0.f, e11 = 0.f, e01 = 0.f;
        for (int j = 0; j < k; ++j) {
            const float wj = P.win[j];
            const int r = (oy + j) * TW + ox;
            m0 = fmaf(wj, hm[r], m0); m1 = fmaf(wj, hm[HH * TW + r], m1);
            e00 = fmaf(wj, hm[2 * HH * TW + r], e00); e11 = fmaf(wj, hm[3 * HH * TW + r], e11); e01 = fmaf(wj, hm[4 * HH * TW + r], e01);
        }
        const float v0 = P.cov_norm * (e00 - m0 * m0), v1 = P.cov_norm * (e11 - m1 * m1), v01 = P.cov_norm * (e01 - m0 * m1);
        const float lum = (2.f * m0 * m1 + P.c1) / (m0 * m0 + m1 * m1 + P.c1);
        const float cs = (2.f * v01 + P.c2) / (v0 + v1 + P.c2);
        acc += lum * cs;
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int i = 0; i < 8; ++i) s += red[i];
        atomicAdd(out + img, s * P.inv_count);
    }
}

}  // namespace

extern "C" int ni_ssim(const float* a, const float* b, float* out_n, int n, int h, int w, int c, const float* win_host, int k, float cov_norm, float c1,
                       float c2, cudaStream_t st) {
    NI_REQUIRE(a && b && out_n && win_host, "ni_ssim: null pointer");
    NI_REQUIRE(n > 0 && c > 0 && k >= 1 && k <= kMaxWin && h >= k && w >= k, "ni_ssim: invalid sizes (window %d on %d x %d)", k, h, w);
    NI_REQUIRE((long long)n * c <= 65535, "ni_ssim: n * c must not exceed 65535");
    SsimParams P;
    for (int i = 0; i < kMaxWin; ++i) P.win[i] = i < k ? win_host[i] : 0.f;
    P.k = k; P.cov_norm = cov_norm; P.c1 = c1; P.c2 = c2;
    const int ow = w - k + 1, oh = h - k + 1;
    P.inv_count = 1.0f / ((float)ow * (float)oh * (float)c);
    NI_CUDA(cudaMemsetAsync(out_n, 0, sizeof(float) * n, st));
    const int HW = TW + k - 1, HH = TH + k - 1;
    const size_t smem = sizeof(float) * (size_t)(2 * HH * HW + 5 * HH * TW);
    dim3 grid((unsigned)ni_cdiv(ow, TW), (unsigned)ni_cdiv(oh, TH), (unsigned)(n * c));
    ssim_kernel<<<grid, 256, smem, st>>>(a, b, out_n, h, w, c, P);
    NI_LAUNCH_CHECK();
    NI_COUNT_LAUNCH(1);
    return NI_OK;
}
