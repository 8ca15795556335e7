// FP32 SIMT implicit-GEMM convolution (fprop / dgrad / wgrad), NHWC activations, HWIO weights.
//
// This is the general path: any channel count, stride 1/2, TF SAME/VALID zero padding, channel pitch/offset
// addressing (concat-free U-Net skip connections), space_to_depth / depth_to_space addressing (transposed 2x2
// convs and the pixel-shuffle outputs), fused bias + activation epilogue. It serves the small-channel layers
// that are not dense contractions (Cin = 3/4, Cout = 12) and validates the tcgen05 path on the device.
// Replaces the cuDNN/Eigen calls behind Keras Conv2D / Conv2DTranspose in the reference
// (models/pipelines.py:190-218, models/forensics.py:65-90, models/compression.py:219-266).
#include "conv_desc.h"
#include "ni_common.cuh"
#include "views.cuh"

namespace {

constexpr int BM = 64, BN = 64, BK = 16, NT = 256;

struct GemmParams {
    TensorView src, dst;
    int kh, kw, stride, pad_t, pad_l, pad_mode;
    int Kc, Nc;          // channels contracted per tap, output channels
    long long M;         // number of target pixels
    int act; float alpha; int accumulate; int bias_mod;
};

__device__ __forceinline__ float apply_act(float v, int act, float alpha) {
    switch (act) {
        case NI_ACT_LEAKY_RELU: return v > 0.f ? v : alpha * v;
        case NI_ACT_RELU: return fmaxf(v, 0.f);
        case NI_ACT_TANH: return tanhf(v);
        case NI_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
        case NI_ACT_CLIP01: return ni_clamp01(v);
        default: return v;
    }
}

// Source coordinate of target pixel coordinate t for tap a. Returns false if it falls in the zero padding.
__device__ __forceinline__ int mirror_idx(int u, int n, int pad_mode) {
    if (pad_mode == NI_PAD_SYMMETRIC) { if (u < 0) u = -u - 1; if (u >= n) u = 2 * n - 1 - u; }
    else { if (u < 0) u = -u; if (u >= n) u = 2 * (n - 1) - u; }
    return u;
}

template <bool DGRAD>
__device__ __forceinline__ bool src_coord(int t, int a, int stride, int pad, int lim, int pad_mode, int& s) {
    if (!DGRAD) {
        s = t * stride + a - pad;
        if (pad_mode != NI_PAD_ZERO) { s = mirror_idx(s, lim, pad_mode); return true; }
    } else {
        const int u = t + pad - a;
        if (u < 0) return false;
        if (stride == 2) { if (u & 1) return false; s = u >> 1; } else s = u;
    }
    return s >= 0 && s < lim;
}

template <bool DGRAD, bool FASTK>
__global__ void __launch_bounds__(NT)
conv_gemm_kernel(GemmParams p, const float* __restrict__ S, const float* __restrict__ Bm, const float* __restrict__ bias,
                 float* __restrict__ D) {
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN];
    const int tid = threadIdx.x;
    const long long m0 = (long long)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int ty = tid / 16, tx = tid % 16;

    // A-load role: pixel am, 4 consecutive k
    const int am = tid / 4, akq = (tid % 4) * 4;
    const long long gm = m0 + am;
    const bool mvalid = gm < p.M;
    int pn = 0, py = 0, px = 0;
    if (mvalid) {
        px = (int)(gm % p.dst.W);
        const long long t = gm / p.dst.W;
        py = (int)(t % p.dst.H);
        pn = (int)(t / p.dst.H);
    }
    // B-load role
    const int bk = tid / 16, bn4 = (tid % 16) * 4;
    const bool bvec = (p.Nc % 4) == 0;

    float acc[4][4] = {};
    const int taps = p.kh * p.kw;
    const int Ktot = taps * p.Kc;
    for (int k0 = 0; k0 < Ktot; k0 += BK) {
        // ---- A tile
        float av[4] = {0.f, 0.f, 0.f, 0.f};
        if (FASTK) {
            const int tap = k0 / p.Kc, c = k0 - tap * p.Kc + akq;
            const int a = tap / p.kw, b = tap - a * p.kw;
            int sy, sx;
            if (mvalid && src_coord<DGRAD>(py, a, p.stride, p.pad_t, p.src.H, p.pad_mode, sy) &&
                src_coord<DGRAD>(px, b, p.stride, p.pad_l, p.src.W, p.pad_mode, sx)) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(S + view_addr(p.src, pn, sy, sx, c)));
                av[0] = v.x; av[1] = v.y; av[2] = v.z; av[3] = v.w;
            }
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int kk = k0 + akq + q;
                if (mvalid && kk < Ktot) {
                    const int tap = kk / p.Kc, c = kk - tap * p.Kc;
                    const int a = tap / p.kw, b = tap - a * p.kw;
                    int sy, sx;
                    if (src_coord<DGRAD>(py, a, p.stride, p.pad_t, p.src.H, p.pad_mode, sy) &&
                        src_coord<DGRAD>(px, b, p.stride, p.pad_l, p.src.W, p.pad_mode, sx))
                        av[q] = __ldg(S + view_addr(p.src, pn, sy, sx, c));
                }
            }
        }
        // ---- B tile
        float bv[4] = {0.f, 0.f, 0.f, 0.f};
        {
            const int kk = k0 + bk;
            if (kk < Ktot) {
                const float* row = Bm + (long long)kk * p.Nc + n0 + bn4;
                if (bvec && n0 + bn4 + 3 < p.Nc) {
                    const float4 v = __ldg(reinterpret_cast<const float4*>(row));
                    bv[0] = v.x; bv[1] = v.y; bv[2] = v.z; bv[3] = v.w;
                } else {
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (n0 + bn4 + q < p.Nc) bv[q] = __ldg(row + q);
                }
            }
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 4; ++q) As[akq + q][am] = av[q];
        *reinterpret_cast<float4*>(&Bs[bk][bn4]) = make_float4(bv[0], bv[1], bv[2], bv[3]);
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            const float ar[4] = {a4.x, a4.y, a4.z, a4.w}, br[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
        }
    }
    // ---- epilogue
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long gmi = m0 + ty * 4 + i;
        if (gmi >= p.M) continue;
        const int ox = (int)(gmi % p.dst.W);
        const long long t = gmi / p.dst.W;
        const int oy = (int)(t % p.dst.H), on = (int)(t / p.dst.H);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int co = n0 + tx * 4 + j;
            if (co >= p.Nc) continue;
            float v = acc[i][j];
            if (bias) v += __ldg(bias + (p.bias_mod > 0 ? co % p.bias_mod : co));
            v = apply_act(v, p.act, p.alpha);
            float* o = D + view_addr(p.dst, on, oy, ox, co);
            *o = p.accumulate ? *o + v : v;
        }
    }
}

// dW[tap][ci][co] += sum over pixels of x[pixel shifted by tap][ci] * dy[pixel][co].
// GEMM view: M = (tap, ci) flattened (so that 3/4-channel layers still fill a 64-wide tile), N = co, K = pixels;
// split-K over pixel ranges with atomic accumulation into dW.
struct WgradParams {
    TensorView xin, dyv;
    int kh, kw, stride, pad_t, pad_l, pad_mode;
    int cin, cout;
    long long npix;      // n * oh * ow
    int pix_per_block;
};

__global__ void __launch_bounds__(NT)
conv_wgrad_kernel(WgradParams p, const float* __restrict__ X, const float* __restrict__ DY, float* __restrict__ DW) {
    __shared__ __align__(16) float As[BK][BM];
    __shared__ __align__(16) float Bs[BK][BN];
    const int tid = threadIdx.x, ty = tid / 16, tx = tid % 16;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int Mtot = p.kh * p.kw * p.cin;
    const long long k_begin = (long long)blockIdx.z * p.pix_per_block;
    const long long k_end = min(k_begin + (long long)p.pix_per_block, p.npix);
    const int lk = tid / 16, l4 = (tid % 16) * 4;
    // the 4 consecutive flattened-M entries of this thread lie in one tap and are 16-byte aligned when cin % 4 == 0
    const bool avec = p.xin.mode == NI_MODE_PLAIN && (p.xin.pitch % 4) == 0 && (p.xin.coff % 4) == 0 && (p.cin % 4) == 0;
    const bool bvec = p.dyv.mode == NI_MODE_PLAIN && (p.dyv.pitch % 4) == 0 && (p.dyv.coff % 4) == 0 && (p.cout % 4) == 0;
    int a_tap[4], a_ci[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int mm = m0 + l4 + q;
        a_tap[q] = mm < Mtot ? mm / p.cin : -1;
        a_ci[q] = mm < Mtot ? mm % p.cin : 0;
    }

    float acc[4][4] = {};
    for (long long k0 = k_begin; k0 < k_end; k0 += BK) {
        const long long gp = k0 + lk;
        float av[4] = {0.f, 0.f, 0.f, 0.f}, bv[4] = {0.f, 0.f, 0.f, 0.f};
        if (gp < k_end) {
            const int ox = (int)(gp % p.dyv.W);
            const long long t = gp / p.dyv.W;
            const int oy = (int)(t % p.dyv.H), on = (int)(t / p.dyv.H);
            if (avec) {
                if (a_tap[0] >= 0) {
                    const int a = a_tap[0] / p.kw, b = a_tap[0] - a * p.kw;
                    int sy, sx;
                    if (src_coord<false>(oy, a, p.stride, p.pad_t, p.xin.H, p.pad_mode, sy) &&
                        src_coord<false>(ox, b, p.stride, p.pad_l, p.xin.W, p.pad_mode, sx)) {
                        const float4 v = __ldg(reinterpret_cast<const float4*>(X + view_addr(p.xin, on, sy, sx, a_ci[0])));
                        av[0] = v.x; av[1] = v.y; av[2] = v.z; av[3] = v.w;
                    }
                }
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (a_tap[q] < 0) continue;
                    const int a = a_tap[q] / p.kw, b = a_tap[q] - a * p.kw;
                    int sy, sx;
                    if (src_coord<false>(oy, a, p.stride, p.pad_t, p.xin.H, p.pad_mode, sy) &&
                        src_coord<false>(ox, b, p.stride, p.pad_l, p.xin.W, p.pad_mode, sx))
                        av[q] = __ldg(X + view_addr(p.xin, on, sy, sx, a_ci[q]));
                }
            }
            if (bvec && n0 + l4 + 3 < p.cout) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(DY + view_addr(p.dyv, on, oy, ox, n0 + l4)));
                bv[0] = v.x; bv[1] = v.y; bv[2] = v.z; bv[3] = v.w;
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (n0 + l4 + q < p.cout) bv[q] = __ldg(DY + view_addr(p.dyv, on, oy, ox, n0 + l4 + q));
            }
        }
        __syncthreads();
        *reinterpret_cast<float4*>(&As[lk][l4]) = make_float4(av[0], av[1], av[2], av[3]);
        *reinterpret_cast<float4*>(&Bs[lk][l4]) = make_float4(bv[0], bv[1], bv[2], bv[3]);
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            const float ar[4] = {a4.x, a4.y, a4.z, a4.w}, br[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int mm = m0 + ty * 4 + i;
        if (mm >= Mtot) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int co = n0 + tx * 4 + j;
            if (co >= p.cout) continue;
            atomicAdd(DW + (long long)mm * p.cout + co, acc[i][j]);
        }
    }
}

// (kh,kw,cin,cout) -> (kh,kw,cout,cin)
__global__ void transpose_io_kernel(const float* __restrict__ w, float* __restrict__ wt, int taps, int cin, int cout) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    const long long total = (long long)taps * cin * cout;
    if (i >= total) return;
    const int ci = (int)(i % cin);
    const long long t = i / cin;
    const int co = (int)(t % cout), tap = (int)(t / cout);
    wt[i] = w[((long long)tap * cin + ci) * cout + co];
}

int check_desc(const ni_conv_desc* d, const char* who) {
    NI_REQUIRE(d, "%s: null descriptor", who);
    NI_REQUIRE(d->n >= 0 && d->h > 0 && d->w > 0 && d->cin > 0 && d->cout > 0 && d->kh > 0 && d->kw > 0 && d->oh > 0 && d->ow > 0,
               "%s: invalid shape", who);
    NI_REQUIRE(d->stride == 1 || d->stride == 2, "%s: stride must be 1 or 2, got %d", who, d->stride);
    NI_REQUIRE(d->in_mode == NI_MODE_PLAIN || (d->in_mode == NI_MODE_BLOCK2 && d->cin % 4 == 0),
               "%s: in_mode BLOCK2 needs cin %% 4 == 0", who);
    NI_REQUIRE(d->out_mode == NI_MODE_PLAIN || (d->out_mode == NI_MODE_BLOCK2 && d->cout % 4 == 0),
               "%s: out_mode BLOCK2 needs cout %% 4 == 0", who);
    const int cin_phys = d->in_mode == NI_MODE_BLOCK2 ? d->cin / 4 : d->cin;
    const int cout_phys = d->out_mode == NI_MODE_BLOCK2 ? d->cout / 4 : d->cout;
    NI_REQUIRE(d->in_pitch >= d->in_coff + cin_phys && d->in_coff >= 0, "%s: input pitch/offset too small", who);
    NI_REQUIRE(d->out_pitch >= d->out_coff + cout_phys && d->out_coff >= 0, "%s: output pitch/offset too small", who);
    return NI_OK;
}

TensorView in_view(const ni_conv_desc* d) { return TensorView{d->h, d->w, d->cin, d->in_pitch, d->in_coff, d->in_mode}; }
TensorView out_view(const ni_conv_desc* d) { return TensorView{d->oh, d->ow, d->cout, d->out_pitch, d->out_coff, d->out_mode}; }

bool fastk_ok(const TensorView& v, int Kc) {
    if (Kc % BK) return false;
    if ((v.pitch % 4) || (v.coff % 4)) return false;
    if (v.mode == NI_MODE_BLOCK2 && ((v.C / 4) % BK)) return false;
    return true;
}

}  // namespace

extern "C" int ni_conv2d_fprop_simt(const ni_conv_desc* d, const float* x, const float* w, const float* bias, float* y,
                                    cudaStream_t st) {
    int rc = check_desc(d, "ni_conv2d_fprop");
    if (rc) return rc;
    NI_REQUIRE(x && w && y, "ni_conv2d_fprop: null pointer");
    if (d->n == 0) return NI_OK;
    GemmParams p;
    p.src = in_view(d); p.dst = out_view(d);
    p.kh = d->kh; p.kw = d->kw; p.stride = d->stride; p.pad_t = d->pad_t; p.pad_l = d->pad_l;
    p.pad_mode = d->pad_mode;
    p.Kc = d->cin; p.Nc = d->cout; p.M = (long long)d->n * d->oh * d->ow;
    p.act = d->act; p.alpha = d->act_alpha; p.accumulate = d->accumulate; p.bias_mod = d->bias_mod;
    dim3 grid(ni_cdiv(p.M, BM), ni_cdiv(p.Nc, BN));
    if (fastk_ok(p.src, p.Kc)) conv_gemm_kernel<false, true><<<grid, NT, 0, st>>>(p, x, w, bias, y);
    else conv_gemm_kernel<false, false><<<grid, NT, 0, st>>>(p, x, w, bias, y);
    NI_LAUNCH_CHECK(); NI_COUNT_LAUNCH(1);
    return NI_OK;
}

// dx = conv-transpose of dy. wt is the weight tensor with the channel axes swapped: (kh, kw, cout, cin).
extern "C" int ni_conv2d_dgrad_simt(const ni_conv_desc* d, const float* dy, const float* wt, float* dx, cudaStream_t st) {
    int rc = check_desc(d, "ni_conv2d_dgrad");
    if (rc) return rc;
    NI_REQUIRE(dy && wt && dx, "ni_conv2d_dgrad: null pointer");
    if (d->n == 0) return NI_OK;
    GemmParams p;
    p.src = out_view(d); p.dst = in_view(d);
    p.kh = d->kh; p.kw = d->kw; p.stride = d->stride; p.pad_t = d->pad_t; p.pad_l = d->pad_l;
    NI_REQUIRE(d->pad_mode == NI_PAD_ZERO, "ni_conv2d_dgrad: mirrored padding is folded by ni_pad_fold, not by dgrad");
    p.pad_mode = NI_PAD_ZERO;
    p.Kc = d->cout; p.Nc = d->cin; p.M = (long long)d->n * d->h * d->w;
    p.act = NI_ACT_NONE; p.alpha = 0.f; p.accumulate = d->accumulate; p.bias_mod = 0;
    dim3 grid(ni_cdiv(p.M, BM), ni_cdiv(p.Nc, BN));
    if (fastk_ok(p.src, p.Kc)) conv_gemm_kernel<true, true><<<grid, NT, 0, st>>>(p, dy, wt, nullptr, dx);
    else conv_gemm_kernel<true, false><<<grid, NT, 0, st>>>(p, dy, wt, nullptr, dx);
    NI_LAUNCH_CHECK(); NI_COUNT_LAUNCH(1);
    return NI_OK;
}

extern "C" int ni_conv2d_wgrad_simt(const ni_conv_desc* d, const float* x, const float* dy, float* dw, cudaStream_t st) {
    int rc = check_desc(d, "ni_conv2d_wgrad");
    if (rc) return rc;
    NI_REQUIRE(x && dy && dw, "ni_conv2d_wgrad: null pointer");
    const int taps = d->kh * d->kw;
    if (!d->accumulate) NI_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)taps * d->cin * d->cout, st));
    if (d->n == 0) return NI_OK;
    WgradParams p;
    p.xin = in_view(d); p.dyv = out_view(d);
    p.kh = d->kh; p.kw = d->kw; p.stride = d->stride; p.pad_t = d->pad_t; p.pad_l = d->pad_l; p.pad_mode = d->pad_mode;
    p.cin = d->cin; p.cout = d->cout; p.npix = (long long)d->n * d->oh * d->ow;
    const int tiles = ni_cdiv((long long)taps * d->cin, BM) * ni_cdiv(d->cout, BN);
    long long splits = (4LL * ni_num_sms() + tiles - 1) / tiles;
    const long long max_splits = (p.npix + 255) / 256;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    if (splits > 65535) splits = 65535;
    long long ppb = (p.npix + splits - 1) / splits;
    ppb = (ppb + BK - 1) / BK * BK;
    splits = (p.npix + ppb - 1) / ppb;
    p.pix_per_block = (int)ppb;
    dim3 grid(ni_cdiv((long long)taps * d->cin, BM), ni_cdiv(d->cout, BN), (unsigned)splits);
    conv_wgrad_kernel<<<grid, NT, 0, st>>>(p, x, dy, dw);
    NI_LAUNCH_CHECK(); NI_COUNT_LAUNCH(1);
    return NI_OK;
}

extern "C" int ni_weight_transpose_io(const float* w, float* wt, int taps, int cin, int cout, cudaStream_t st) {
    NI_REQUIRE(w && wt && taps > 0 && cin > 0 && cout > 0, "ni_weight_transpose_io: invalid arguments");
    const long long total = (long long)taps * cin * cout;
    transpose_io_kernel<<<ni_cdiv(total, 256), 256, 0, st>>>(w, wt, taps, cin, cout);
    NI_LAUNCH_CHECK(); NI_COUNT_LAUNCH(1);
    return NI_OK;
}
