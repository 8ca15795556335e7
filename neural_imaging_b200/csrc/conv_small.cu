// Direct FP32 convolutions for the layers that are NOT dense contractions (3/4 input channels or 3/12 output
// channels): the FAN front end (constrained 5x5 3->3 filter + 5x5 3->32 conv, reference models/forensics.py:62-68),
// the U-Net input (3x3 4->32) and output (3x3 32->12 + depth_to_space) layers (models/pipelines.py:190,215-218).
// An implicit-GEMM tile would be > 90% padding for these shapes (measured: 64.6 ms for the 3->32 dgrad of one step
// with the generic 64x64 tile), so they get thread-per-pixel stencils with all weights resident in shared memory.
//
//   fprop : one thread = one output pixel x all COUT channels (COUT accumulators), 16x16-pixel tile + halo in smem.
//   dgrad : the same kernel with the roles of the channel axes swapped and the taps flipped (prep kernel).
//   wgrad : one thread = one (filter row a, ci, co) triple with K accumulators (the K taps of the row), sweeping
//           the pixels of 8x32 tiles four at a time; per-block register accumulation, one atomicAdd per output at the end.
#include "conv_desc.h"
#include "ni_common.cuh"
#include "views.cuh"

namespace {

constexpr int TS = 16;   // fprop tile: 16 x 16 output pixels, 256 threads

struct SmallParams {
    TensorView src, dst;
    int n, k, pad_t, pad_l, pad_mode;
    int act, bias_mod; float alpha; int accumulate;
};

__device__ __forceinline__ int mirror(int u, int n, int mode) {
    if (mode == NI_PAD_SYMMETRIC) { if (u < 0) u = -u - 1; if (u >= n) u = 2 * n - 1 - u; }
    else { if (u < 0) u = -u; if (u >= n) u = 2 * (n - 1) - u; }
    return u;
}

__device__ __forceinline__ float act_small(float v, int act, float alpha) {
    switch (act) {
        case NI_ACT_LEAKY_RELU: return v > 0.f ? v : alpha * v;
        case NI_ACT_RELU: return fmaxf(v, 0.f);
        case NI_ACT_TANH: return tanhf(v);
        case NI_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
        case NI_ACT_CLIP01: return ni_clamp01(v);
        default: return v;
    }
}

// pixel stride of the shared input tile: keeps 128-bit channel loads of neighbouring pixels on different banks
template <int CIN> struct TileStride { static constexpr int v = (CIN % 32 == 0) ? CIN + 4 : CIN; };

template <int CIN, int COUT, int K>
__global__ void __launch_bounds__(256)
conv_small_fprop_kernel(SmallParams p, const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                        float* __restrict__ y) {
    constexpr int COUTP = (COUT + 3) / 4 * 4;
    constexpr int IT = TS + K - 1;
    constexpr int PS = TileStride<CIN>::v;
    extern __shared__ __align__(16) float smem[];
    float* sw = smem;                              // [K*K][CIN][COUTP]
    float* sx = smem + K * K * CIN * COUTP;        // [IT][IT][PS]
    const int tid = threadIdx.x;
    const int tiles_x = (p.dst.W + TS - 1) / TS, tiles_y = (p.dst.H + TS - 1) / TS;
    const int tx0 = (blockIdx.x % tiles_x) * TS, ty0 = ((blockIdx.x / tiles_x) % tiles_y) * TS, n = blockIdx.x / (tiles_x * tiles_y);

    for (int i = tid; i < K * K * CIN * COUTP; i += 256) {
        const int co = i % COUTP, r = i / COUTP;
        sw[i] = co < COUT ? w[r * COUT + co] : 0.f;
    }
    for (int i = tid; i < IT * IT * CIN; i += 256) {
        const int c = i % CIN, px = (i / CIN) % IT, py = i / (CIN * IT);
        int sy = ty0 + py - p.pad_t, sx_ = tx0 + px - p.pad_l;
        float v = 0.f;
        if (p.pad_mode != NI_PAD_ZERO) { sy = mirror(sy, p.src.H, p.pad_mode); sx_ = mirror(sx_, p.src.W, p.pad_mode); }
        if (sy >= 0 && sy < p.src.H && sx_ >= 0 && sx_ < p.src.W) v = __ldg(x + view_addr(p.src, n, sy, sx_, c));
        sx[(py * IT + px) * PS + c] = v;
    }
    __syncthreads();

    const int lx = tid % TS, ly = tid / TS;
    float acc[COUTP];
#pragma unroll
    for (int j = 0; j < COUTP; ++j) acc[j] = 0.f;
#pragma unroll 1
    for (int a = 0; a < K; ++a) {
#pragma unroll
        for (int b = 0; b < K; ++b) {
            const float* px = sx + ((ly + a) * IT + lx + b) * PS;
            const float4* wt = reinterpret_cast<const float4*>(sw + (a * K + b) * CIN * COUTP);
            if (CIN % 4 == 0) {
#pragma unroll
                for (int c4 = 0; c4 < CIN / 4; ++c4) {
                    const float4 xv = *reinterpret_cast<const float4*>(px + c4 * 4);
                    const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e)
#pragma unroll
                        for (int j = 0; j < COUTP / 4; ++j) {
                            const float4 wv = wt[(c4 * 4 + e) * (COUTP / 4) + j];
                            acc[4 * j] = fmaf(xs[e], wv.x, acc[4 * j]); acc[4 * j + 1] = fmaf(xs[e], wv.y, acc[4 * j + 1]);
                            acc[4 * j + 2] = fmaf(xs[e], wv.z, acc[4 * j + 2]); acc[4 * j + 3] = fmaf(xs[e], wv.w, acc[4 * j + 3]);
                        }
                }
            } else {
#pragma unroll
                for (int c = 0; c < CIN; ++c) {
                    const float xs = px[c];
#pragma unroll
                    for (int j = 0; j < COUTP / 4; ++j) {
                        const float4 wv = wt[c * (COUTP / 4) + j];
                        acc[4 * j] = fmaf(xs, wv.x, acc[4 * j]); acc[4 * j + 1] = fmaf(xs, wv.y, acc[4 * j + 1]);
                        acc[4 * j + 2] = fmaf(xs, wv.z, acc[4 * j + 2]); acc[4 * j + 3] = fmaf(xs, wv.w, acc[4 * j + 3]);
                    }
                }
            }
        }
    }
    const int ox = tx0 + lx, oy = ty0 + ly;
    if (ox >= p.dst.W || oy >= p.dst.H) return;
    if (p.dst.mode == NI_MODE_PLAIN && (COUT % 4 == 0) && (p.dst.pitch % 4 == 0) && (p.dst.coff % 4 == 0)) {
        float4* o = reinterpret_cast<float4*>(y + view_addr(p.dst, n, oy, ox, 0));
#pragma unroll
        for (int j = 0; j < COUT / 4; ++j) {
            float v[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int co = 4 * j + e;
                float t = acc[co];
                if (bias) t += __ldg(bias + (p.bias_mod > 0 ? co % p.bias_mod : co));
                v[e] = act_small(t, p.act, p.alpha);
            }
            float4 r = make_float4(v[0], v[1], v[2], v[3]);
            if (p.accumulate) { const float4 old = o[j]; r.x += old.x; r.y += old.y; r.z += old.z; r.w += old.w; }
            o[j] = r;
        }
    } else {
#pragma unroll
        for (int co = 0; co < COUT; ++co) {
            float t = acc[co];
            if (bias) t += __ldg(bias + (p.bias_mod > 0 ? co % p.bias_mod : co));
            t = act_small(t, p.act, p.alpha);
            float* o = y + view_addr(p.dst, n, oy, ox, co);
            *o = p.accumulate ? *o + t : t;
        }
    }
}

// (K,K,cin,cout) -> (K,K,cout,cin) with both spatial axes flipped: dgrad as a forward conv of dy
__global__ void flip_transpose_kernel(const float* __restrict__ w, float* __restrict__ wf, int k, int cin, int cout) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    const int total = k * k * cin * cout;
    if (i >= total) return;
    const int ci = i % cin, co = (i / cin) % cout, tap = i / (cin * cout);
    const int a = tap / k, b = tap % k;
    wf[i] = w[(((k - 1 - a) * k + (k - 1 - b)) * cin + ci) * cout + co];
}

// ---------------------------------------------------------------------------------------------------- wgrad
constexpr int WH = 8, WW = 32;   // wgrad pixel tile

struct SmallWgradParams {
    TensorView xin, dyv;
    int n, k, pad_t, pad_l, pad_mode;
    int tiles_x, tiles_y, tiles_total;
};

template <int CIN, int COUT, int K>
__global__ void __launch_bounds__(256)
conv_small_wgrad_kernel(SmallWgradParams p, const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dw) {
    constexpr int XH = WH + K - 1, XW = WW + K - 1;
    constexpr int NC = K * CIN * COUT;                 // (filter row, ci, co) combos
    constexpr int PER = (NC + 255) / 256;
    extern __shared__ __align__(16) float smem[];
    float* sx = smem;                                  // [XH][XW][CIN]
    float* sd = smem + XH * XW * CIN;                  // [WH][WW][COUT]
    const int tid = threadIdx.x;
    float acc[PER][K];
#pragma unroll
    for (int j = 0; j < PER; ++j)
#pragma unroll
        for (int b = 0; b < K; ++b) acc[j][b] = 0.f;

    for (int tile = blockIdx.x; tile < p.tiles_total; tile += gridDim.x) {
        const int tx0 = (tile % p.tiles_x) * WW, ty0 = ((tile / p.tiles_x) % p.tiles_y) * WH, n = tile / (p.tiles_x * p.tiles_y);
        __syncthreads();
        for (int i = tid; i < XH * XW * CIN; i += 256) {
            const int c = i % CIN, px = (i / CIN) % XW, py = i / (CIN * XW);
            int sy = ty0 + py - p.pad_t, sxx = tx0 + px - p.pad_l;
            float v = 0.f;
            if (p.pad_mode != NI_PAD_ZERO) { sy = mirror(sy, p.xin.H, p.pad_mode); sxx = mirror(sxx, p.xin.W, p.pad_mode); }
            if (sy >= 0 && sy < p.xin.H && sxx >= 0 && sxx < p.xin.W) v = __ldg(x + view_addr(p.xin, n, sy, sxx, c));
            sx[i] = v;
        }
        for (int i = tid; i < WH * WW * COUT; i += 256) {
            const int c = i % COUT, px = (i / COUT) % WW, py = i / (COUT * WW);
            const int oy = ty0 + py, ox = tx0 + px;
            sd[i] = (oy < p.dyv.H && ox < p.dyv.W) ? __ldg(dy + view_addr(p.dyv, n, oy, ox, c)) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            const int cmb = tid + 256 * j;
            if (cmb >= NC) continue;
            const int co = cmb % COUT, ci = (cmb / COUT) % CIN, a = cmb / (COUT * CIN);
#pragma unroll 1
            for (int yy = 0; yy < WH; ++yy) {
                const float* xr = sx + ((yy + a) * XW) * CIN + ci;
                const float* dr = sd + (yy * WW) * COUT + co;
#pragma unroll 2
                for (int xq = 0; xq < WW; xq += 4) {
                    float xv[K + 3];
#pragma unroll
                    for (int t = 0; t < K + 3; ++t) xv[t] = xr[(xq + t) * CIN];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float g = dr[(xq + e) * COUT];
#pragma unroll
                        for (int b = 0; b < K; ++b) acc[j][b] = fmaf(xv[e + b], g, acc[j][b]);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < PER; ++j) {
        const int cmb = tid + 256 * j;
        if (cmb >= NC) continue;
        const int co = cmb % COUT, ci = (cmb / COUT) % CIN, a = cmb / (COUT * CIN);
#pragma unroll
        for (int b = 0; b < K; ++b) atomicAdd(dw + (((a * K + b) * CIN) + ci) * COUT + co, acc[j][b]);
    }
}

struct Shape { int cin, cout, k; };

template <int CIN, int COUT, int K>
int launch_fprop(const SmallParams& p, const float* x, const float* w, const float* bias, float* y, cudaStream_t st) {
    constexpr int COUTP = (COUT + 3) / 4 * 4;
    constexpr int IT = TS + K - 1;
    const size_t smem = sizeof(float) * (K * K * CIN * COUTP + IT * IT * TileStride<CIN>::v);
    NI_CUDA(cudaFuncSetAttribute(conv_small_fprop_kernel<CIN, COUT, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int tiles = ((p.dst.W + TS - 1) / TS) * ((p.dst.H + TS - 1) / TS) * p.n;
    conv_small_fprop_kernel<CIN, COUT, K><<<tiles, 256, smem, st>>>(p, x, w, bias, y);
    NI_LAUNCH_CHECK(); NI_COUNT_LAUNCH(1);
    return NI_OK;
}

template <int CIN, int COUT, int K>
int launch_wgrad(SmallWgradParams p, const float* x, const float* dy, float* dw, cudaStream_t st) {
    const size_t smem = sizeof(float) * ((WH + K - 1) * (WW + K - 1) * CIN + WH * WW * COUT);
    NI_CUDA(cudaFuncSetAttribute(conv_small_wgrad_kernel<CIN, COUT, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    p.tiles_x = (p.dyv.W + WW - 1) / WW; p.tiles_y = (p.dyv.H + WH - 1) / WH; p.tiles_total = p.tiles_x * p.tiles_y * p.n;
    int grid = p.tiles_total < 6 * ni_num_sms() ? p.tiles_total : 6 * ni_num_sms();
    conv_small_wgrad_kernel<CIN, COUT, K><<<grid, 256, smem, st>>>(p, x, dy, dw);
    NI_LAUNCH_CHECK(); NI_COUNT_LAUNCH(1);
    return NI_OK;
}

// FAN / U-Net ends; INet + ClassicISP 1x1 chains; DNet ends; TwitterDCN output layer
#define NI_SMALL_SHAPES(X)                                                                              \
    X(3, 3, 5) X(3, 32, 5) X(4, 32, 3) X(32, 12, 3) X(32, 3, 5) X(12, 32, 3)                            \
    X(4, 12, 1) X(3, 3, 1) X(3, 12, 1) X(12, 3, 1) X(12, 12, 1) X(3, 3, 3)                              \
    X(4, 64, 3) X(6, 64, 3) X(64, 6, 3) X(64, 3, 1) X(3, 64, 1) X(64, 12, 3) X(12, 64, 3)

bool has_shape(int cin, int cout, int k) {
#define X(a, b, c) if (cin == a && cout == b && k == c) return true;
    NI_SMALL_SHAPES(X)
#undef X
    return false;
}

}  // namespace

int ni_get_scratch2(size_t bytes, float** out);

// op: 0 fprop, 1 dgrad, 2 wgrad
extern "C" int ni_conv2d_small_supported(const ni_conv_desc* d, int op) {
    if (!d || d->n <= 0 || d->stride != 1 || d->kh != d->kw) return 0;
    if (op == 0) return d->in_mode == NI_MODE_PLAIN && has_shape(d->cin, d->cout, d->kh);
    if (op == 1) return d->pad_mode == NI_PAD_ZERO && d->in_mode == NI_MODE_PLAIN && has_shape(d->cout, d->cin, d->kh);
    return d->in_mode == NI_MODE_PLAIN && has_shape(d->cin, d->cout, d->kh);
}

extern "C" int ni_conv2d_fprop_small(const ni_conv_desc* d, const float* x, const float* w, const float* bias, float* y, cudaStream_t st) {
    NI_REQUIRE(ni_conv2d_small_supported(d, 0) && x && w && y, "ni_conv2d_fprop_small: unsupported problem or null pointer");
    SmallParams p;
    p.src = TensorView{d->h, d->w, d->cin, d->in_pitch, d->in_coff, d->in_mode};
    p.dst = TensorView{d->oh, d->ow, d->cout, d->out_pitch, d->out_coff, d->out_mode};
    p.n = d->n; p.k = d->kh; p.pad_t = d->pad_t; p.pad_l = d->pad_l; p.pad_mode = d->pad_mode;
    p.act = d->act; p.bias_mod = d->bias_mod; p.alpha = d->act_alpha; p.accumulate = d->accumulate;
#define X(a, b, c) if (d->cin == a && d->cout == b && d->kh == c) return launch_fprop<a, b, c>(p, x, w, bias, y, st);
    NI_SMALL_SHAPES(X)
#undef X
    return NI_ERR_UNSUPPORTED;
}

// dx = forward conv of dy with the flipped / channel-swapped filter (stride 1, zero padding K-1-pad).
extern "C" int ni_conv2d_dgrad_small(const ni_conv_desc* d, const float* dy, const float* w, float* dx, cudaStream_t st) {
    NI_REQUIRE(ni_conv2d_small_supported(d, 1) && dy && w && dx, "ni_conv2d_dgrad_small: unsupported problem or null pointer");
    float* wf = nullptr;
    const int k = d->kh, total = k * k * d->cin * d->cout;
    int rc = ni_get_scratch2(sizeof(float) * (size_t)total, &wf);
    if (rc) return rc;
    flip_transpose_kernel<<<ni_cdiv(total, 256), 256, 0, st>>>(w, wf, k, d->cin, d->cout);
    NI_LAUNCH_CHECK(); NI_COUNT_LAUNCH(1);
    SmallParams p;
    p.src = TensorView{d->oh, d->ow, d->cout, d->out_pitch, d->out_coff, d->out_mode};
    p.dst = TensorView{d->h, d->w, d->cin, d->in_pitch, d->in_coff, d->in_mode};
    p.n = d->n; p.k = k; p.pad_t = k - 1 - d->pad_t; p.pad_l = k - 1 - d->pad_l; p.pad_mode = NI_PAD_ZERO;
    p.act = NI_ACT_NONE; p.bias_mod = 0; p.alpha = 0.f; p.accumulate = d->accumulate;
#define X(a, b, c) if (d->cout == a && d->cin == b && d->kh == c) return launch_fprop<a, b, c>(p, dy, wf, nullptr, dx, st);
    NI_SMALL_SHAPES(X)
#undef X
    return NI_ERR_UNSUPPORTED;
}

extern "C" int ni_conv2d_wgrad_small(const ni_conv_desc* d, const float* x, const float* dy, float* dw, cudaStream_t st) {
    NI_REQUIRE(ni_conv2d_small_supported(d, 2) && x && dy && dw, "ni_conv2d_wgrad_small: unsupported problem or null pointer");
    if (!d->accumulate) NI_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)d->kh * d->kw * d->cin * d->cout, st));
    SmallWgradParams p;
    p.xin = TensorView{d->h, d->w, d->cin, d->in_pitch, d->in_coff, d->in_mode};
    p.dyv = TensorView{d->oh, d->ow, d->cout, d->out_pitch, d->out_coff, d->out_mode};
    p.n = d->n; p.k = d->kh; p.pad_t = d->pad_t; p.pad_l = d->pad_l; p.pad_mode = d->pad_mode;
#define X(a, b, c) if (d->cin == a && d->cout == b && d->kh == c) return launch_wgrad<a, b, c>(p, x, dy, dw, st);
    NI_SMALL_SHAPES(X)
#undef X
    return NI_ERR_UNSUPPORTED;
}
