// Photo-manipulation kernels (forward + backward) and the channel down-sampling used by the workflow.
// Reference: helpers/tf_helpers.py:68-184,271-287 (manipulation_*, soft_quantization) and
// workflows/manipulation_classification.py:231-245 (run_downsampling). All tensors are NHWC float32 with C = 3.
// These are HBM-bound gather stencils: one thread per pixel (all three channels), neighbours come from L1/L2.
#include "ni_common.cuh"

namespace {

constexpr int kT = 256;

// ---------------------------------------------------------------- padding index maps (tf.pad semantics)
__device__ __forceinline__ int reflect_idx(int u, int n) {  // REFLECT: edge pixel not repeated
    if (n == 1) return 0;
    if (u < 0) u = -u;
    if (u >= n) u = 2 * (n - 1) - u;
    return u;
}
__device__ __forceinline__ int symm_idx(int u, int n) {  // SYMMETRIC: edge pixel repeated
    if (u < 0) u = -u - 1;
    if (u >= n) u = 2 * n - 1 - u;
    return u;
}
// All padded coordinates u in [-pad, n-1+pad] that map onto p. Returns count (<= 3).
__device__ __forceinline__ int reflect_aliases(int p, int n, int pad, int (&u)[3]) {
    int c = 0;
    u[c++] = p;
    if (p >= 1 && p <= pad) u[c++] = -p;
    const int hi = 2 * (n - 1) - p;
    if (p <= n - 2 && hi <= n - 1 + pad) u[c++] = hi;
    return c;
}
__device__ __forceinline__ int symm_aliases(int p, int n, int pad, int (&u)[3]) {
    int c = 0;
    u[c++] = p;
    if (p <= pad - 1) u[c++] = -p - 1;
    if (p >= n - pad) u[c++] = 2 * n - 1 - p;
    return c;
}

// ---------------------------------------------------------------- HSV (tensorflow/core/kernels/colorspace_op.h)
__device__ __forceinline__ void rgb_to_hsv(float r, float g, float b, float& h, float& s, float& v) {
    v = fmaxf(r, fmaxf(g, b));
    const float range = v - fminf(r, fminf(g, b));
    s = v > 0.f ? range / v : 0.f;
    const float norm = (1.f / range) * (1.f / 6.f);
    float hh;
    if (r == v) hh = norm * (g - b);
    else if (g == v) hh = norm * (b - r) + 2.f / 6.f;
    else hh = norm * (r - g) + 4.f / 6.f;
    if (!(range > 0.f)) hh = 0.f;
    if (hh < 0.f) hh += 1.f;
    h = hh;
}
__device__ __forceinline__ void hsv_to_rgb(float h, float s, float v, float& r, float& g, float& b) {
    const float dh = h * 6.f;
    const float dr = fminf(fmaxf(fabsf(dh - 3.f) - 1.f, 0.f), 1.f);
    const float dg = fminf(fmaxf(-fabsf(dh - 2.f) + 2.f, 0.f), 1.f);
    const float db = fminf(fmaxf(-fabsf(dh - 4.f) + 2.f, 0.f), 1.f);
    const float one_s = -s + 1.f;
    r = (one_s + s * dr) * v;
    g = (one_s + s * dg) * v;
    b = (one_s + s * db) * v;
}

__device__ __forceinline__ float3 ld3(const float* p) { return make_float3(p[0], p[1], p[2]); }
__device__ __forceinline__ void st3(float* p, float a, float b, float c) { p[0] = a; p[1] = b; p[2] = c; }

// ---------------------------------------------------------------- sharpen (tf_helpers.py:156-184)
// f[9]: 3x3 filter for the H and V channels; the S channel takes the single tap [2,2] (reference quirk).
struct Filt9 { float f[9]; };

__global__ void sharpen_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int N, int H, int W, Filt9 flt) {
    const long long i = (long long)blockIdx.x * kT + threadIdx.x;
    const long long total = (long long)N * H * W;
    if (i >= total) return;
    const int px = (int)(i % W), py = (int)((i / W) % H);
    const float* img = x + (i / ((long long)H * W)) * (long long)H * W * 3;
    float ah = 0.f, av = 0.f, as = 0.f;
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            const int sy = symm_idx(py + a - 1, H), sx = symm_idx(px + b - 1, W);
            const float3 p = ld3(img + ((long long)sy * W + sx) * 3);
            float h, s, v;
            rgb_to_hsv(p.x, p.y, p.z, h, s, v);
            ah = fmaf(flt.f[a * 3 + b], h, ah);
            av = fmaf(flt.f[a * 3 + b], v, av);
            if (a == 2 && b == 2) as = s;
        }
    float r, g, b;
    hsv_to_rgb(ah, as, av, r, g, b);
    st3(y + i * 3, ni_clamp01(r), ni_clamp01(g), ni_clamp01(b));
}

// ---------------------------------------------------------------- gaussian blur (tf_helpers.py:113-125)
struct Filt1D { float w[16]; int k; };  // 2-D filter = outer(w, w) is NOT what the reference uses; see Filt2D

struct Filt2D { float w[11 * 11]; int k; };

__global__ void gaussian_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, unsigned char* __restrict__ mask,
                                    int N, int H, int W, Filt2D flt, int clip) {
    const long long i = (long long)blockIdx.x * kT + threadIdx.x;
    const long long total = (long long)N * H * W;
    if (i >= total) return;
    const int px = (int)(i % W), py = (int)((i / W) % H);
    const float* img = x + (i / ((long long)H * W)) * (long long)H * W * 3;
    const int k = flt.k, pad = k / 2;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int a = 0; a < k; ++a) {
        const int sy = reflect_idx(py + a - pad, H);
        for (int b = 0; b < k; ++b) {
            const int sx = reflect_idx(px + b - pad, W);
            const float3 p = ld3(img + ((long long)sy * W + sx) * 3);
            const float w = flt.w[a * k + b];
            a0 = fmaf(w, p.x, a0); a1 = fmaf(w, p.y, a1); a2 = fmaf(w, p.z, a2);
        }
    }
    if (mask) {
        const unsigned char m = (unsigned char)(((a0 >= 0.f && a0 <= 1.f) ? 1 : 0) | ((a1 >= 0.f && a1 <= 1.f) ? 2 : 0) |
                                                ((a2 >= 0.f && a2 <= 1.f) ? 4 : 0));
        mask[i] = clip ? m : (unsigned char)7;
    }
    if (clip) { a0 = ni_clamp01(a0); a1 = ni_clamp01(a1); a2 = ni_clamp01(a2); }
    st3(y + i * 3, a0, a1, a2);
}

// dx[p] = sum over padded aliases u of p, taps d: w[d] * (mask*dy)[u - d + pad ...]  (transpose of pad+conv)
__global__ void gaussian_bwd_kernel(const float* __restrict__ dy, const unsigned char* __restrict__ mask,
                                    float* __restrict__ dx, int N, int H, int W, Filt2D flt, float scale, int accumulate) {
    const long long i = (long long)blockIdx.x * kT + threadIdx.x;
    const long long total = (long long)N * H * W;
    if (i >= total) return;
    const int px = (int)(i % W), py = (int)((i / W) % H);
    const long long ioff = (i / ((long long)H * W)) * (long long)H * W;
    const int k = flt.k, pad = k / 2;
    int uy[3], ux[3];
    const int ny = reflect_aliases(py, H, pad, uy), nx = reflect_aliases(px, W, pad, ux);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int iy = 0; iy < ny; ++iy)
        for (int a = 0; a < k; ++a) {
            const int qy = uy[iy] - (a - pad);  // output row whose tap a reads padded row uy
            if (qy < 0 || qy >= H) continue;
            for (int ix = 0; ix < nx; ++ix)
                for (int b = 0; b < k; ++b) {
                    const int qx = ux[ix] - (b - pad);
                    if (qx < 0 || qx >= W) continue;
                    const long long q = ioff + (long long)qy * W + qx;
                    const unsigned char m = mask ? mask[q] : (unsigned char)7;
                    const float3 g = ld3(dy + q * 3);
                    const float w = flt.w[a * k + b];
                    if (m & 1) a0 = fmaf(w, g.x, a0);
                    if (m & 2) a1 = fmaf(w, g.y, a1);
                    if (m & 4) a2 = fmaf(w, g.z, a2);
                }
        }
    float* o = dx + i * 3;
    if (accumulate) st3(o, o[0] + scale * a0, o[1] + scale * a1, o[2] + scale * a2);
    else st3(o, scale * a0, scale * a1, scale * a2);
}

// ---------------------------------------------------------------- bilinear resize (tf.image.resize v2: half-pixel centres)
__device__ __forceinline__ void interp_w(int o, float scale, int in_size, int& lo, int& hi, float& lerp) {
    const float in = ((float)o + 0.5f) * scale - 0.5f;
    const float in_f = floorf(in);
    lo = max((int)in_f, 0);
    hi = min((int)ceilf(in), in_size - 1);
    lerp = in - in_f;
}

__global__ void resize_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int N, int IH, int IW, int OH, int OW,
                                  int clip) {
    const long long i = (long long)blockIdx.x * kT + threadIdx.x;
    const long long total = (long long)N * OH * OW;
    if (i >= total) return;
    const int ox = (int)(i % OW), oy = (int)((i / OW) % OH);
    const long long n = i / ((long long)OH * OW);
    const float sy = (float)IH / (float)OH, sx = (float)IW / (float)OW;
    int y0, y1, x0, x1; float ly, lx;
    interp_w(oy, sy, IH, y0, y1, ly);
    interp_w(ox, sx, IW, x0, x1, lx);
    const float* img = x + n * (long long)IH * IW * 3;
    const float3 tl = ld3(img + ((long long)y0 * IW + x0) * 3), tr = ld3(img + ((long long)y0 * IW + x1) * 3);
    const float3 bl = ld3(img + ((long long)y1 * IW + x0) * 3), br = ld3(img + ((long long)y1 * IW + x1) * 3);
    float o[3];
    const float tlv[3] = {tl.x, tl.y, tl.z}, trv[3] = {tr.x, tr.y, tr.z}, blv[3] = {bl.x, bl.y, bl.z}, brv[3] = {br.x, br.y, br.z};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float top = tlv[c] + (trv[c] - tlv[c]) * lx;
        const float bot = blv[c] + (brv[c] - blv[c]) * lx;
        o[c] = top + (bot - top) * ly;
        if (clip) o[c] = ni_clamp01(o[c]);
    }
    st3(y + i * 3, o[0], o[1], o[2]);
}

// Scatter form of ResizeBilinearGrad; dx must be zero-initialised.
__global__ void resize_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, int N, int IH, int IW, int OH, int OW,
                                  float gscale) {
    const long long i = (long long)blockIdx.x * kT + threadIdx.x;
    const long long total = (long long)N * OH * OW;
    if (i >= total) return;
    const int ox = (int)(i % OW), oy = (int)((i / OW) % OH);
    const long long n = i / ((long long)OH * OW);
    const float sy = (float)IH / (float)OH, sx = (float)IW / (float)OW;
    int y0, y1, x0, x1; float ly, lx;
    interp_w(oy, sy, IH, y0, y1, ly);
    interp_w(ox, sx, IW, x0, x1, lx);
    float* img = dx + n * (long long)IH * IW * 3;
    const float3 g = ld3(dy + i * 3);
    const float gv[3] = {g.x * gscale, g.y * gscale, g.z * gscale};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        atomicAdd(img + ((long long)y0 * IW + x0) * 3 + c, gv[c] * (1.f - ly) * (1.f - lx));
        atomicAdd(img + ((long long)y0 * IW + x1) * 3 + c, gv[c] * (1.f - ly) * lx);
        atomicAdd(img + ((long long)y1 * IW + x0) * 3 + c, gv[c] * ly * (1.f - lx));
        atomicAdd(img + ((long long)y1 * IW + x1) * 3 + c, gv[c] * ly * lx);
    }
}

// ---------------------------------------------------------------- soft 8-bit quantisation helpers (tf_helpers.py:271-277)
constexpr float kTwoPi = 6.2831855f;
__device__ __forceinline__ float soft_quant(float t) { return rintf(255.f * t) / 255.f; }            // forward value
__device__ __forceinline__ float soft_quant_grad(float t) { return 1.f - cosf(kTwoPi * (255.f * t)); }  // d/dt

// Philox4x32-10 -> 4 uniforms -> Box-Muller normals (own generator; cannot match tf.random.normal bit-wise).
__device__ __forceinline__ void philox4(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                        uint32_t (&out)[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__device__ __forceinline__ void normal3(unsigned long long seed, unsigned long long idx, float (&z)[3]) {
    uint32_t r[4];
    philox4((uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)idx, (uint32_t)(idx >> 32), 0x6E695F62u, 0u, r);
    const float u0 = ((float)r[0] + 0.5f) * 2.3283064e-10f, u1 = ((float)r[1] + 0.5f) * 2.3283064e-10f;
    const float u2 = ((float)r[2] + 0.5f) * 2.3283064e-10f, u3 = ((float)r[3] + 0.5f) * 2.3283064e-10f;
    const float ra = sqrtf(-2.f * logf(u0)), rb = sqrtf(-2.f * logf(u2));
    z[0] = ra * cospif(2.f * u1);
    z[1] = ra * sinpif(2.f * u1);
    z[2] = rb * cospif(2.f * u3);
}

// awgn: y = clip(soft_quant(x + s*noise)); gamma: y = clip(soft_quant(x^s), 1/255, 1)^(1/s)
template <bool BWD>
__global__ void awgn_kernel(const float* __restrict__ x, const float* __restrict__ noise, const float* __restrict__ dy,
                            float* __restrict__ out, long long npix, float strength, unsigned long long seed, int accumulate) {
    const long long i = (long long)blockIdx.x * kT + threadIdx.x;
    if (i >= npix) return;
    float z[3];
    if (noise) { z[0] = noise[i * 3]; z[1] = noise[i * 3 + 1]; z[2] = noise[i * 3 + 2]; }
    else normal3(seed, (unsigned long long)i, z);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float t = x[i * 3 + c] + strength * z[c];
        const float q = soft_quant(t);
        if (!BWD) out[i * 3 + c] = ni_clamp01(q);
        else {
            const float g = (q >= 0.f && q <= 1.f) ? dy[i * 3 + c] * soft_quant_grad(t) : 0.f;
            out[i * 3 + c] = accumulate ? out[i * 3 + c] + g : g;
        }
    }
}

template <bool BWD>
__global__ void gamma_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ out,
                             long long nval, float s, int accumulate) {
    const long long i = (long long)blockIdx.x * kT + threadIdx.x;
    if (i >= nval) return;
    const float xv = x[i];
    const float t = powf(xv, s);
    const float q = soft_quant(t);
    const float lo = 1.f / 255.f;
    const float qc = fminf(fmaxf(q, lo), 1.f);
    if (!BWD) out[i] = powf(qc, 1.f / s);
    else {
        float g = 0.f;
        if (q >= lo && q <= 1.f) g = dy[i] * (1.f / s) * powf(qc, 1.f / s - 1.f) * soft_quant_grad(t) * s * powf(xv, s - 1.f);
        out[i] = accumulate ? out[i] + g : g;
    }
}

// ---------------------------------------------------------------- median (tf_helpers.py:91-110)
// Selects, per channel, the element ranked (k*k+1)/2 - 1 in descending order (ties: lower patch index first),
// i.e. the true median for odd k*k. BWD routes the gradient to the selected source pixel (top_k gather grad).
template <bool BWD>
__global__ void median_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ out, int N,
                              int H, int W, int k) {
    const long long i = (long long)blockIdx.x * kT + threadIdx.x;
    const long long total = (long long)N * H * W;
    if (i >= total) return;
    const int px = (int)(i % W), py = (int)((i / W) % H);
    const long long ioff = (i / ((long long)H * W)) * (long long)H * W;
    const int pad = k / 2, area = k * k, want = (area + 1) / 2 - 1;
    for (int c = 0; c < 3; ++c) {
        float mv = 0.f; long long msrc = 0;
        for (int e = 0; e < area; ++e) {
            const int sy = reflect_idx(py + e / k - pad, H), sx = reflect_idx(px + e % k - pad, W);
            const float ve = x[(ioff + (long long)sy * W + sx) * 3 + c];
            int rank = 0;
            for (int j = 0; j < area; ++j) {
                const int ty = reflect_idx(py + j / k - pad, H), tx = reflect_idx(px + j % k - pad, W);
                const float vj = x[(ioff + (long long)ty * W + tx) * 3 + c];
                rank += (vj > ve) || (vj == ve && j < e);
            }
            if (rank == want) { mv = ve; msrc = (ioff + (long long)sy * W + sx) * 3 + c; }
        }
        if (!BWD) out[i * 3 + c] = mv;
        else atomicAdd(out + msrc, dy[i * 3 + c]);
    }
}

// ---------------------------------------------------------------- average pooling k x k, stride k, SAME
__global__ void avgpool_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int N, int H, int W, int k, int OH,
                                   int OW) {
    const long long i = (long long)blockIdx.x * kT + threadIdx.x;
    const long long total = (long long)N * OH * OW;
    if (i >= total) return;
    const int ox = (int)(i % OW), oy = (int)((i / OW) % OH);
    const float* img = x + (i / ((long long)OH * OW)) * (long long)H * W * 3;
    // TF SAME: pad_total = max((O-1)*k + k - I, 0), before = total/2; padded cells are excluded from the mean
    const int pty = max((OH - 1) * k + k - H, 0) / 2, ptx = max((OW - 1) * k + k - W, 0) / 2;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f; int cnt = 0;
    for (int a = 0; a < k; ++a) {
        const int sy = oy * k + a - pty;
        if (sy < 0 || sy >= H) continue;
        for (int b = 0; b < k; ++b) {
            const int sx = ox * k + b - ptx;
            if (sx < 0 || sx >= W) continue;
            const float3 p = ld3(img + ((long long)sy * W + sx) * 3);
            a0 += p.x; a1 += p.y; a2 += p.z; ++cnt;
        }
    }
    const float inv = 1.f / (float)cnt;
    st3(y + i * 3, a0 * inv, a1 * inv, a2 * inv);
}
__global__ void avgpool_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, int N, int H, int W, int k, int OH,
                                   int OW) {
    const long long i = (long long)blockIdx.x * kT + threadIdx.x;
    const long long total = (long long)N * H * W;
    if (i >= total) return;
    const int px = (int)(i % W), py = (int)((i / W) % H);
    const long long n = i / ((long long)H * W);
    const int pty = max((OH - 1) * k + k - H, 0) / 2, ptx = max((OW - 1) * k + k - W, 0) / 2;
    const int oy = (py + pty) / k, ox = (px + ptx) / k;
    const int y0 = max(oy * k - pty, 0), y1 = min(oy * k - pty + k, H), x0 = max(ox * k - ptx, 0), x1 = min(ox * k - ptx + k, W);
    const float inv = 1.f / (float)((y1 - y0) * (x1 - x0));
    const float3 g = ld3(dy + ((n * OH + oy) * OW + ox) * 3);
    st3(dx + i * 3, g.x * inv, g.y * inv, g.z * inv);
}

// ---------------------------------------------------------------- misc elementwise
__global__ void axpy_kernel(float* __restrict__ y, const float* __restrict__ x, float a, long long n) {
    const long long i = (long long)blockIdx.x * kT + threadIdx.x;
    if (i < n) y[i] = fmaf(a, x[i], y[i]);
}

inline int grid_for(long long n) { return ni_cdiv(n, kT); }

}  // namespace

#define NI_MANIP_COMMON(name)                                                                                \
    NI_REQUIRE(n >= 0 && h > 0 && w > 0, name ": invalid shape %d x %d x %d", n, h, w);                      \
    if (n == 0) return NI_OK;                                                                                \
    const long long npix = (long long)n * h * w;                                                             \
    (void)npix;

extern "C" int ni_manip_sharpen_fwd(const float* x, float* y, int n, int h, int w, const float* filt9, cudaStream_t st) {
    NI_REQUIRE(x && y && filt9, "ni_manip_sharpen_fwd: null pointer");
    NI_MANIP_COMMON("ni_manip_sharpen_fwd");
    Filt9 f;
    for (int i = 0; i < 9; ++i) f.f[i] = filt9[i];
    sharpen_fwd_kernel<<<grid_for(npix), kT, 0, st>>>(x, y, n, h, w, f);
    NI_LAUNCH_CHECK(); NI_COUNT_LAUNCH(1);
    return NI_OK;
}

static int make_filt2d(Filt2D& f, const float* filt, int k, const char* who) {
    NI_REQUIRE(filt && k >= 1 && k <= 11 && (k % 2) == 1, "%s: kernel size must be odd and in [1, 11], got %d", who, k);
    f.k = k;
    for (int i = 0; i < k * k; ++i) f.w[i] = filt[i];
    return NI_OK;
}

extern "C" int ni_manip_gaussian_fwd(const float* x, float* y, unsigned char* mask, int n, int h, int w, const float* filt,
                                     int k, int clip, cudaStream_t st) {
    NI_REQUIRE(x && y, "ni_manip_gaussian_fwd: null pointer");
    NI_MANIP_COMMON("ni_manip_gaussian_fwd");
    NI_REQUIRE(h > k / 2 && w > k / 2, "ni_manip_gaussian_fwd: REFLECT padding needs H, W > %d", k / 2);
    Filt2D f;
    int rc = make_filt2d(f, filt, k, "ni_manip_gaussian_fwd");
    if (rc) return rc;
    gaussian_fwd_kernel<<<grid_for(npix), kT, 0, st>>>(x, y, mask, n, h, w, f, clip);
    NI_LAUNCH_CHECK(); NI_COUNT_LAUNCH(1);
    return NI_OK;
}

extern "C" int ni_manip_gaussian_bwd(const float* dy, const unsigned char* mask, float* dx, int n, int h, int w,
                                     const float* filt, int k, float scale, int accumulate, cudaStream_t st) {
    NI_REQUIRE(dy && dx, "ni_manip_gaussian_bwd: null pointer");
    NI_MANIP_COMMON("ni_manip_gaussian_bwd");
    Filt2D f;
    int rc = make_filt2d(f, filt, k, "ni_manip_gaussian_bwd");
    if (rc) return rc;
    gaussian_bwd_kernel<<<grid_for(npix), kT, 0, st>>>(dy, mask, dx, n, h, w, f, scale, accumulate);
    NI_LAUNCH_CHECK(); NI_COUNT_LAUNCH(1);
    return NI_OK;
}

extern "C" int ni_resize_bilinear_fwd(const float* x, float* y, int n, int ih, int iw, int oh, int ow, int clip,
                                      cudaStream_t st) {
    NI_REQUIRE(x && y, "ni_resize_bilinear_fwd: null pointer");
    NI_REQUIRE(n >= 0 && ih > 0 && iw > 0 && oh > 0 && ow > 0, "ni_resize_bilinear_fwd: invalid shape");
    if (n == 0) return NI_OK;
    resize_fwd_kernel<<<grid_for((long long)n * oh * ow), kT, 0, st>>>(x, y, n, ih, iw, oh, ow, clip);
    NI_LAUNCH_CHECK(); NI_COUNT_LAUNCH(1);
    return NI_OK;
}

// dx (n,ih,iw,3) must be zero-filled by the caller (scatter-add).
extern "C" int ni_resize_bilinear_bwd(const float* dy, float* dx, int n, int ih, int iw, int oh, int ow, float scale,
                                      cudaStream_t st) {
    NI_REQUIRE(dy && dx, "ni_resize_bilinear_bwd: null pointer");
    NI_REQUIRE(n >= 0 && ih > 0 && iw > 0 && oh > 0 && ow > 0, "ni_resize_bilinear_bwd: invalid shape");
    if (n == 0) return NI_OK;
    resize_bwd_kernel<<<grid_for((long long)n * oh * ow), kT, 0, st>>>(dy, dx, n, ih, iw, oh, ow, scale);
    NI_LAUNCH_CHECK(); NI_COUNT_LAUNCH(1);
    return NI_OK;
}

extern "C" int ni_manip_awgn_fwd(const float* x, const float* noise, float* y, int n, int h, int w, float strength,
                                 unsigned long long seed, cudaStream_t st) {
    NI_REQUIRE(x && y, "ni_manip_awgn_fwd: null pointer");
    NI_MANIP_COMMON("ni_manip_awgn_fwd");
    awgn_kernel<false><<<grid_for(npix), kT, 0, st>>>(x, noise, nullptr, y, npix, strength, seed, 0);
    NI_LAUNCH_CHECK(); NI_COUNT_LAUNCH(1);
    return NI_OK;
}
extern "C" int ni_manip_awgn_bwd(const float* x, const float* noise, const float* dy, float* dx, int n, int h, int w,
                                 float strength, unsigned long long seed, int accumulate, cudaStream_t st) {
    NI_REQUIRE(x && dy && dx, "ni_manip_awgn_bwd: null pointer");
    NI_MANIP_COMMON("ni_manip_awgn_bwd");
    awgn_kernel<true><<<grid_for(npix), kT, 0, st>>>(x, noise, dy, dx, npix, strength, seed, accumulate);
    NI_LAUNCH_CHECK(); NI_COUNT_LAUNCH(1);
    return NI_OK;
}

extern "C" int ni_manip_gamma_fwd(const float* x, float* y, int n, int h, int w, float strength, cudaStream_t st) {
    NI_REQUIRE(x && y, "ni_manip_gamma_fwd: null pointer");
    NI_MANIP_COMMON("ni_manip_gamma_fwd");
    NI_REQUIRE(strength != 0.f, "ni_manip_gamma_fwd: strength must be non-zero");
    gamma_kernel<false><<<grid_for(npix * 3), kT, 0, st>>>(x, nullptr, y, npix * 3, strength, 0);
    NI_LAUNCH_CHECK(); NI_COUNT_LAUNCH(1);
    return NI_OK;
}
extern "C" int ni_manip_gamma_bwd(const float* x, const float* dy, float* dx, int n, int h, int w, float strength,
                                  int accumulate, cudaStream_t st) {
    NI_REQUIRE(x && dy && dx, "ni_manip_gamma_bwd: null pointer");
    NI_MANIP_COMMON("ni_manip_gamma_bwd");
    NI_REQUIRE(strength != 0.f, "ni_manip_gamma_bwd: strength must be non-zero");
    gamma_kernel<true><<<grid_for(npix * 3), kT, 0, st>>>(x, dy, dx, npix * 3, strength, accumulate);
    NI_LAUNCH_CHECK(); NI_COUNT_LAUNCH(1);
    return NI_OK;
}

extern "C" int ni_manip_median_fwd(const float* x, float* y, int n, int h, int w, int k, cudaStream_t st) {
    NI_REQUIRE(x && y, "ni_manip_median_fwd: null pointer");
    NI_MANIP_COMMON("ni_manip_median_fwd");
    NI_REQUIRE(k >= 1 && (k % 2) == 1 && h > k / 2 && w > k / 2, "ni_manip_median_fwd: kernel must be odd and < 2*min(H,W)");
    median_kernel<false><<<grid_for(npix), kT, 0, st>>>(x, nullptr, y, n, h, w, k);
    NI_LAUNCH_CHECK(); NI_COUNT_LAUNCH(1);
    return NI_OK;
}
// dx must be zero-filled (or hold the value to accumulate into): the gradient is scattered with atomics.
extern "C" int ni_manip_median_bwd(const float* x, const float* dy, float* dx, int n, int h, int w, int k, cudaStream_t st) {
    NI_REQUIRE(x && dy && dx, "ni_manip_median_bwd: null pointer");
    NI_MANIP_COMMON("ni_manip_median_bwd");
    NI_REQUIRE(k >= 1 && (k % 2) == 1 && h > k / 2 && w > k / 2, "ni_manip_median_bwd: kernel must be odd and < 2*min(H,W)");
    median_kernel<true><<<grid_for(npix), kT, 0, st>>>(x, dy, dx, n, h, w, k);
    NI_LAUNCH_CHECK(); NI_COUNT_LAUNCH(1);
    return NI_OK;
}

extern "C" int ni_avgpool_fwd(const float* x, float* y, int n, int h, int w, int k, cudaStream_t st) {
    NI_REQUIRE(x && y, "ni_avgpool_fwd: null pointer");
    NI_MANIP_COMMON("ni_avgpool_fwd");
    NI_REQUIRE(k >= 1, "ni_avgpool_fwd: invalid factor %d", k);
    const int oh = (h + k - 1) / k, ow = (w + k - 1) / k;
    avgpool_fwd_kernel<<<grid_for((long long)n * oh * ow), kT, 0, st>>>(x, y, n, h, w, k, oh, ow);
    NI_LAUNCH_CHECK(); NI_COUNT_LAUNCH(1);
    return NI_OK;
}
extern "C" int ni_avgpool_bwd(const float* dy, float* dx, int n, int h, int w, int k, cudaStream_t st) {
    NI_REQUIRE(dy && dx, "ni_avgpool_bwd: null pointer");
    NI_MANIP_COMMON("ni_avgpool_bwd");
    NI_REQUIRE(k >= 1, "ni_avgpool_bwd: invalid factor %d", k);
    const int oh = (h + k - 1) / k, ow = (w + k - 1) / k;
    avgpool_bwd_kernel<<<grid_for(npix), kT, 0, st>>>(dy, dx, n, h, w, k, oh, ow);
    NI_LAUNCH_CHECK(); NI_COUNT_LAUNCH(1);
    return NI_OK;
}

// y += a * x   (gradient accumulation between branches; n = number of floats)
extern "C" int ni_axpy(float* y, const float* x, float a, long long n, cudaStream_t st) {
    NI_REQUIRE(y && x && n >= 0, "ni_axpy: invalid arguments");
    if (n == 0) return NI_OK;
    axpy_kernel<<<grid_for(n), kT, 0, st>>>(y, x, a, n);
    NI_LAUNCH_CHECK(); NI_COUNT_LAUNCH(1);
    return NI_OK;
}
