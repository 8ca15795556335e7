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

__global__ void __launch_bounds__(256) ssim_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out,
                                                   float* __restrict__ stats, int h, int w, int c, const SsimParams P) {
    extern __shared__ float sm[];
    const int k = P.k, HW = TW + k - 1, HH = TH + k - 1;
    float* sa = sm;                       // [HH][HW]
    float* sb = sa + HH * HW;             // [HH][HW]
    float* hm = sb + HH * HW;             // 5 x [HH][TW] horizontally filtered moments
    __shared__ float red[16];
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
    float acc = 0.f, acc_cs = 0.f;
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
        acc_cs += cs;
    }
    for (int o = 16; o > 0; o >>= 1) {
        acc += __shfl_xor_sync(0xffffffffu, acc, o);
        acc_cs += __shfl_xor_sync(0xffffffffu, acc_cs, o);
    }
    if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = acc; red[8 + (threadIdx.x >> 5)] = acc_cs; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f, s_cs = 0.f;
        for (int i = 0; i < 8; ++i) { s += red[i]; s_cs += red[8 + i]; }
        if (out) atomicAdd(out + img, s * P.inv_count);
        if (stats) {     // per (image, channel): mean SSIM and mean contrast-structure over the VALID region
            atomicAdd(stats + 2 * blockIdx.z, s * P.inv_count * (float)c);
            atomicAdd(stats + 2 * blockIdx.z + 1, s_cs * P.inv_count * (float)c);
        }
    }
}

// ---- backward of the mean SSIM / mean cs of every (image, channel) with respect to `a` ------------------------------------------
// With the local moments m0 = G*a, m1 = G*b, e00 = G*a^2, e11 = G*b^2, e01 = G*ab at every VALID window position p:
//   lum = (2 m0 m1 + c1) / Ld,  cs = (2 cn (e01 - m0 m1) + c2) / Cd,  Ld = m0^2 + m1^2 + c1,  Cd = cn (e00 + e11 - m0^2 - m1^2) + c2
// and an upstream gradient gS on lum*cs and gC on cs (k = gS lum + gC):
//   A = d/dm0 = 2 gS cs (m1 - lum m0) / Ld + 2 k cn (cs m0 - m1) / Cd,   B = d/de01 = 2 k cn / Cd,   D = d/de00 = -k cn cs / Cd
//   da(q) = sum_p G(q - p) [A(p) + b(q) B(p) + 2 a(q) D(p)]
// One CTA produces a BTW x BTH tile of da: inputs with a halo of 2(k-1), moments and the three maps on the (k-1)-haloed tile, then the
// transposed separable filter — everything stays in shared memory, HBM sees a, b once (plus halo) and da once.
constexpr int BTW = 32, BTH = 16;

__global__ void __launch_bounds__(256) ssim_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ coef,
                                                       float* __restrict__ da, int h, int w, int c, const SsimParams P, int accumulate) {
    extern __shared__ float sm[];
    const int k = P.k, R = k - 1;
    const int IW = BTW + 2 * R, IH = BTH + 2 * R, MW = BTW + R, MH = BTH + R;
    float* sa = sm;                        // [IH][IW]
    float* sb = sa + IH * IW;              // [IH][IW]
    float* hm = sb + IH * IW;              // 5 x [IH][MW] horizontally filtered moments; later 3 x [MH][BTW] horizontally filtered maps
    float* mp = hm + 5 * IH * MW;          // 3 x [MH][MW] maps A, B, D
    const int img = blockIdx.z / c, ch = blockIdx.z - img * c;
    const int qx0 = blockIdx.x * BTW, qy0 = blockIdx.y * BTH;
    const int ow = w - k + 1, oh = h - k + 1;
    const long long base = (long long)img * h * w * c + ch;
    const float inv_pos = 1.0f / ((float)ow * (float)oh);
    const float gS = coef[2 * blockIdx.z] * inv_pos, gC = coef[2 * blockIdx.z + 1] * inv_pos;
    for (int i = threadIdx.x; i < IH * IW; i += 256) {
        const int iy = i / IW, ix = i - iy * IW;
        const int y = qy0 - R + iy, x = qx0 - R + ix;
        float va = 0.f, vb = 0.f;
        if (y >= 0 && y < h && x >= 0 && x < w) {
            const long long o = base + ((long long)y * w + x) * c;
            va = __ldg(a + o); vb = __ldg(b + o);
        }
        sa[i] = va; sb[i] = vb;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < IH * MW; i += 256) {
        const int iy = i / MW, mx = i - iy * MW;
        float m0 = 0.f, m1 = 0.f, e00 = 0.f, e11 = 0.f, e01 = 0.f;
        for (int j = 0; j < k; ++j) {
            const float wj = P.win[j], x = sa[iy * IW + mx + j], y = sb[iy * IW + mx + j];
            m0 = fmaf(wj, x, m0); m1 = fmaf(wj, y, m1);
            e00 = fmaf(wj, x * x, e00); e11 = fmaf(wj, y * y, e11); e01 = fmaf(wj, x * y, e01);
        }
        hm[i] = m0; hm[IH * MW + i] = m1; hm[2 * IH * MW + i] = e00; hm[3 * IH * MW + i] = e11; hm[4 * IH * MW + i] = e01;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < MH * MW; i += 256) {
        const int my = i / MW, mx = i - my * MW;
        const int py = qy0 - R + my, px = qx0 - R + mx;
        float A = 0.f, B = 0.f, D = 0.f;
        if (py >= 0 && py < oh && px >= 0 && px < ow) {
            float m0 = 0.f, m1 = 0.f, e00 = 0.f, e11 = 0.f, e01 = 0.f;
            for (int j = 0; j < k; ++j) {
                const float wj = P.win[j];
                const int r = (my + j) * MW + mx;
                m0 = fmaf(wj, hm[r], m0); m1 = fmaf(wj, hm[IH * MW + r], m1);
                e00 = fmaf(wj, hm[2 * IH * MW + r], e00); e11 = fmaf(wj, hm[3 * IH * MW + r], e11); e01 = fmaf(wj, hm[4 * IH * MW + r], e01);
            }
            const float cn = P.cov_norm;
            const float Ld = m0 * m0 + m1 * m1 + P.c1;
            const float Cd = cn * (e00 + e11 - m0 * m0 - m1 * m1) + P.c2;
            const float lum = (2.f * m0 * m1 + P.c1) / Ld;
            const float cs = (2.f * cn * (e01 - m0 * m1) + P.c2) / Cd;
            const float kk = gS * lum + gC;
            A = 2.f * gS * cs * (m1 - lum * m0) / Ld + 2.f * kk * cn * (cs * m0 - m1) / Cd;
            B = 2.f * kk * cn / Cd;
            D = -kk * cn * cs / Cd;
        }
        mp[i] = A; mp[MH * MW + i] = B; mp[2 * MH * MW + i] = D;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < MH * BTW; i += 256) {       // transposed horizontal pass: p_x = q_x - j
        const int my = i / BTW, x = i - my * BTW;
        float tA = 0.f, tB = 0.f, tD = 0.f;
        for (int j = 0; j < k; ++j) {
            const float wj = P.win[j];
            const int r = my * MW + x + R - j;
            tA = fmaf(wj, mp[r], tA); tB = fmaf(wj, mp[MH * MW + r], tB); tD = fmaf(wj, mp[2 * MH * MW + r], tD);
        }
        hm[i] = tA; hm[MH * BTW + i] = tB; hm[2 * MH * BTW + i] = tD;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < BTH * BTW; i += 256) {      // transposed vertical pass and the chain rule through a^2, ab
        const int y = i / BTW, x = i - y * BTW;
        const int qy = qy0 + y, qx = qx0 + x;
        if (qy >= h || qx >= w) continue;
        float tA = 0.f, tB = 0.f, tD = 0.f;
        for (int j = 0; j < k; ++j) {
            const float wj = P.win[j];
            const int r = (y + R - j) * BTW + x;
            tA = fmaf(wj, hm[r], tA); tB = fmaf(wj, hm[MH * BTW + r], tB); tD = fmaf(wj, hm[2 * MH * BTW + r], tD);
        }
        const float va = sa[(y + R) * IW + x + R], vb = sb[(y + R) * IW + x + R];
        const float g = tA + vb * tB + 2.f * va * tD;
        const long long o = base + ((long long)qy * w + qx) * c;
        da[o] = accumulate ? da[o] + g : g;
    }
}

// tf_helpers.ssim_loss / msssim_loss on the per-(image, channel) statistics of `levels` scales: loss and the coefficients the backward
// kernels need. stats / coef: [levels][n*c][2] (mean SSIM, mean cs). levels == 1: plain SSIM (no relu, as tf.image.ssim).
__global__ void msssim_combine_kernel(const float* __restrict__ stats, float* __restrict__ coef, float* __restrict__ loss_acc, int nc, int levels,
                                      float w0, float w1, float w2, float w3, float w4, float loss_scale, float grad_scale) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float contrib = 0.f;
    if (i < nc) {
        const float wts[5] = {w0, w1, w2, w3, w4};
        const float gl = -255.0f / (float)nc * grad_scale;        // d loss / d (ms-)ssim of this (image, channel)
        if (levels == 1) {
            contrib = 255.0f * (1.0f - stats[2 * i]) / (float)nc;
            coef[2 * i] = gl; coef[2 * i + 1] = 0.f;
        } else {
            float v[5], ms = 1.f;
            for (int l = 0; l < levels; ++l) {
                v[l] = fmaxf(stats[((long long)l * nc + i) * 2 + (l == levels - 1 ? 0 : 1)], 0.f);
                ms *= powf(v[l], wts[l]);
            }
            contrib = 255.0f * (1.0f - ms) / (float)nc;
            for (int l = 0; l < levels; ++l) {
                const float g = v[l] > 0.f ? gl * wts[l] * ms / v[l] : 0.f;
                const long long o = ((long long)l * nc + i) * 2;
                coef[o] = l == levels - 1 ? g : 0.f;
                coef[o + 1] = l == levels - 1 ? 0.f : g;
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
    if ((threadIdx.x & 31) == 0 && contrib != 0.f) atomicAdd(loss_acc, contrib * loss_scale);
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
    ssim_kernel<<<grid, 256, smem, st>>>(a, b, out_n, nullptr, h, w, c, P);
    NI_LAUNCH_CHECK();
    NI_COUNT_LAUNCH(1);
    return NI_OK;
}

static int fill_params(SsimParams& P, const float* win_host, int k, float cov_norm, float c1, float c2, int h, int w, int c) {
    for (int i = 0; i < kMaxWin; ++i) P.win[i] = i < k ? win_host[i] : 0.f;
    P.k = k; P.cov_norm = cov_norm; P.c1 = c1; P.c2 = c2;
    P.inv_count = 1.0f / ((float)(w - k + 1) * (float)(h - k + 1) * (float)c);
    return 0;
}

extern "C" int ni_ssim_stats(const float* a, const float* b, float* stats_nc2, int n, int h, int w, int c, const float* win_host, int k,
                             float cov_norm, float c1, float c2, cudaStream_t st) {
    NI_REQUIRE(a && b && stats_nc2 && win_host, "ni_ssim_stats: null pointer");
    NI_REQUIRE(n > 0 && c > 0 && k >= 1 && k <= kMaxWin && h >= k && w >= k, "ni_ssim_stats: invalid sizes (window %d on %d x %d)", k, h, w);
    NI_REQUIRE((long long)n * c <= 65535, "ni_ssim_stats: n * c must not exceed 65535");
    SsimParams P;
    fill_params(P, win_host, k, cov_norm, c1, c2, h, w, c);
    NI_CUDA(cudaMemsetAsync(stats_nc2, 0, sizeof(float) * 2 * n * c, st));
    const int HW = TW + k - 1, HH = TH + k - 1;
    const size_t smem = sizeof(float) * (size_t)(2 * HH * HW + 5 * HH * TW);
    dim3 grid((unsigned)ni_cdiv(w - k + 1, TW), (unsigned)ni_cdiv(h - k + 1, TH), (unsigned)(n * c));
    ssim_kernel<<<grid, 256, smem, st>>>(a, b, nullptr, stats_nc2, h, w, c, P);
    NI_LAUNCH_CHECK();
    NI_COUNT_LAUNCH(1);
    return NI_OK;
}

extern "C" int ni_ssim_bwd(const float* a, const float* b, const float* coef_nc2, float* da, int accumulate, int n, int h, int w, int c,
                           const float* win_host, int k, float cov_norm, float c1, float c2, cudaStream_t st) {
    NI_REQUIRE(a && b && coef_nc2 && da && win_host, "ni_ssim_bwd: null pointer");
    NI_REQUIRE(n > 0 && c > 0 && k >= 1 && k <= kMaxWin && h >= k && w >= k, "ni_ssim_bwd: invalid sizes (window %d on %d x %d)", k, h, w);
    NI_REQUIRE((long long)n * c <= 65535, "ni_ssim_bwd: n * c must not exceed 65535");
    SsimParams P;
    fill_params(P, win_host, k, cov_norm, c1, c2, h, w, c);
    const int R = k - 1, IW = BTW + 2 * R, IH = BTH + 2 * R, MW = BTW + R, MH = BTH + R;
    const size_t smem = sizeof(float) * (size_t)(2 * IH * IW + 5 * IH * MW + 3 * MH * MW);
    static bool attr_set = false;
    if (!attr_set) {
        NI_CUDA(cudaFuncSetAttribute(ssim_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(float) * (2 * 44 * 60 + 5 * 44 * 46 + 3 * 30 * 46))));
        attr_set = true;
    }
    dim3 grid((unsigned)ni_cdiv(w, BTW), (unsigned)ni_cdiv(h, BTH), (unsigned)(n * c));
    ssim_bwd_kernel<<<grid, 256, smem, st>>>(a, b, coef_nc2, da, h, w, c, P, accumulate);
    NI_LAUNCH_CHECK();
    NI_COUNT_LAUNCH(1);
    return NI_OK;
}

extern "C" int ni_msssim_combine(const float* stats, float* coef, float* loss_acc, int n, int c, int levels, const float* weights_host,
                                 float loss_scale, float grad_scale, cudaStream_t st) {
    NI_REQUIRE(stats && coef && loss_acc, "ni_msssim_combine: null pointer");
    NI_REQUIRE(n > 0 && c > 0 && levels >= 1 && levels <= 5 && (levels == 1 || weights_host), "ni_msssim_combine: invalid arguments");
    float wts[5] = {1.f, 0.f, 0.f, 0.f, 0.f};
    for (int i = 0; i < levels && weights_host; ++i) wts[i] = weights_host[i];
    const int nc = n * c;
    msssim_combine_kernel<<<ni_cdiv(nc, 128), 128, 0, st>>>(stats, coef, loss_acc, nc, levels, wts[0], wts[1], wts[2], wts[3], wts[4], loss_scale,
                                                            grad_scale);
    NI_LAUNCH_CHECK();
    NI_COUNT_LAUNCH(1);
    return NI_OK;
}
