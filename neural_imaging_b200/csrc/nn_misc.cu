// Non-convolution layer kernels of the training path: max-pool, activation backward + bias gradient, global
// average pooling, softmax + sparse categorical cross-entropy (Keras semantics), image losses, the constrained
// residual filter normalisation (models/layers.py:45-53), mirrored-pad gradient folding, fused Keras-Adam.
#include "conv_desc.h"
#include "ni_common.cuh"
#include "views.cuh"

namespace {

constexpr int kT = 256;
inline int grid_for(long long n) { return ni_cdiv(n, kT); }

// ------------------------------------------------------------------ max pooling 2x2 stride 2 (SAME or VALID)
// x: (n,h,w,c) with pitch/coff; y: (n,oh,ow,c) with pitch/coff. Channel-vectorised by 4 when possible.
__global__ void maxpool_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int n, int h, int w, int c, int oh,
                                   int ow, int xp, int xo, int yp, int yo) {
    const long long i = (long long)blockIdx.x * kT + threadIdx.x;
    const long long total = (long long)n * oh * ow * c;
    if (i >= total) return;
    const int ch = (int)(i % c);
    long long t = i / c;
    const int ox = (int)(t % ow); t /= ow;
    const int oy = (int)(t % oh);
    const int nn = (int)(t / oh);
    float m = -INFINITY;
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const int sy = 2 * oy + a, sx = 2 * ox + b;
            if (sy < h && sx < w) m = fmaxf(m, x[(((long long)nn * h + sy) * w + sx) * xp + xo + ch]);
        }
    y[(((long long)nn * oh + oy) * ow + ox) * yp + yo + ch] = m;
}

// dx[pixel] = (pixel is the first maximum of its window ? dy[window] : 0) + (add ? add[pixel] : 0)
__global__ void maxpool_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ add,
                                   float* __restrict__ dx, int n, int h, int w, int c, int oh, int ow, int xp, int xo,
                                   int dyp, int dyo, int addp, int addo, int dxp, int dxo) {
    const long long i = (long long)blockIdx.x * kT + threadIdx.x;
    const long long total = (long long)n * h * w * c;
    if (i >= total) return;
    const int ch = (int)(i % c);
    long long t = i / c;
    const int px = (int)(t % w); t /= w;
    const int py = (int)(t % h);
    const int nn = (int)(t / h);
    const int oy = py >> 1, ox = px >> 1;
    float g = 0.f;
    if (oy < oh && ox < ow) {
        float m = -INFINITY; int arg = -1;
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                const int sy = 2 * oy + a, sx = 2 * ox + b;
                if (sy < h && sx < w) {
                    const float v = x[(((long long)nn * h + sy) * w + sx) * xp + xo + ch];
                    if (v > m) { m = v; arg = a * 2 + b; }
                }
            }
        if (arg == (py & 1) * 2 + (px & 1)) g = dy[(((long long)nn * oh + oy) * ow + ox) * dyp + dyo + ch];
    }
    const long long pix = ((long long)nn * h + py) * w + px;
    if (add) g += add[pix * addp + addo + ch];
    dx[pix * dxp + dxo + ch] = g;
}

// ------------------------------------------------------------------ activation backward (+ bias gradient)
__device__ __forceinline__ float act_grad_from_out(float y, int act, float alpha) {
    switch (act) {
        case NI_ACT_LEAKY_RELU: return y > 0.f ? 1.f : alpha;
        case NI_ACT_RELU: return y > 0.f ? 1.f : 0.f;
        case NI_ACT_TANH: return 1.f - y * y;
        case NI_ACT_SIGMOID: return y * (1.f - y);
        default: return 1.f;  // none; clip01 is straight-through in the ISPs / DCN (models/pipelines.py:223)
    }
}

// dy <- dy * act'(y) (in place, skipped for ACT_NONE) and dbias[c] += sum_pixels dy, on LOGICAL (n,H,W,C) views.
// blockDim = (32, 8): x over channels (coalesced), y over pixels; grid.x over channel groups, grid.y over pixel chunks.
__global__ void act_bwd_bias_kernel(const float* __restrict__ y, float* __restrict__ dy, float* __restrict__ dbias,
                                    long long npix, TensorView yv, TensorView dv, int act, float alpha,
                                    int pix_per_block, int bias_mod) {
    __shared__ float red[8][33];
    const int c = dv.C;
    const int ch = blockIdx.x * 32 + threadIdx.x;
    const long long p0 = (long long)blockIdx.y * pix_per_block;
    const long long p1 = min(p0 + (long long)pix_per_block, npix);
    float s = 0.f;
    if (ch < c) {
        for (long long p = p0 + threadIdx.y; p < p1; p += 8) {
            const int px = (int)(p % dv.W);
            const long long t = p / dv.W;
            const int py = (int)(t % dv.H), pn = (int)(t / dv.H);
            const long long ad = view_addr(dv, pn, py, px, ch);
            float g = dy[ad];
            if (act != NI_ACT_NONE && act != NI_ACT_CLIP01) {
                g *= act_grad_from_out(y[view_addr(yv, pn, py, px, ch)], act, alpha);
                dy[ad] = g;
            }
            s += g;
        }
    }
    red[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && ch < c && dbias) {
        float tsum = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) tsum += red[k][threadIdx.x];
        atomicAdd(dbias + (bias_mod > 0 ? ch % bias_mod : ch), tsum);
    }
}

// Vectorised (float4 over channels) variants for plain NHWC buffers with 16-byte aligned channel slices: these layers are
// pure HBM streams (12 B / element), the scalar versions above were ALU-bound on 64-bit index arithmetic.
__global__ void __launch_bounds__(256)
act_bwd_bias_vec_kernel(const float* __restrict__ y, float* __restrict__ dy, float* __restrict__ dbias, long long npix, int c,
                        int yp, int yo, int dp, int dof, int act, float alpha, int pix_per_block, int bias_mod, int nx, int ny) {
    extern __shared__ float4 red4[];                   // [ny][nx]
    const int tx = threadIdx.x % nx, ty = threadIdx.x / nx;
    const long long p0 = (long long)blockIdx.y * pix_per_block;
    const long long p1 = min(p0 + (long long)pix_per_block, npix);
    const bool need_act = act != NI_ACT_NONE && act != NI_ACT_CLIP01;
    for (int c4 = blockIdx.x * nx + tx; c4 * 4 < c; c4 += gridDim.x * nx) {
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        // four pixels per trip with all eight loads issued before the first store: the in-place update otherwise serialises one
        // memory round trip per pixel (the stores may alias the next loads as far as the compiler knows)
        long long p = p0 + ty;
        for (; p + 3LL * ny < p1; p += 4LL * ny) {
            float4 g[4], yv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) g[u] = *reinterpret_cast<const float4*>(dy + (p + (long long)u * ny) * dp + dof + c4 * 4);
            if (need_act) {
#pragma unroll
                for (int u = 0; u < 4; ++u) yv[u] = __ldg(reinterpret_cast<const float4*>(y + (p + (long long)u * ny) * yp + yo + c4 * 4));
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    g[u].x *= act_grad_from_out(yv[u].x, act, alpha); g[u].y *= act_grad_from_out(yv[u].y, act, alpha);
                    g[u].z *= act_grad_from_out(yv[u].z, act, alpha); g[u].w *= act_grad_from_out(yv[u].w, act, alpha);
                    *reinterpret_cast<float4*>(dy + (p + (long long)u * ny) * dp + dof + c4 * 4) = g[u];
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) { s.x += g[u].x; s.y += g[u].y; s.z += g[u].z; s.w += g[u].w; }
        }
        for (; p < p1; p += ny) {
            float4* gp = reinterpret_cast<float4*>(dy + p * dp + dof + c4 * 4);
            float4 g = *gp;
            if (need_act) {
                const float4 yv = __ldg(reinterpret_cast<const float4*>(y + p * yp + yo + c4 * 4));
                g.x *= act_grad_from_out(yv.x, act, alpha); g.y *= act_grad_from_out(yv.y, act, alpha);
                g.z *= act_grad_from_out(yv.z, act, alpha); g.w *= act_grad_from_out(yv.w, act, alpha);
                *gp = g;
            }
            s.x += g.x; s.y += g.y; s.z += g.z; s.w += g.w;
        }
        if (dbias) {
            red4[ty * nx + tx] = s;
            __syncthreads();
            if (ty == 0) {
                float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int k = 0; k < ny; ++k) { const float4 v = red4[k * nx + tx]; t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w; }
                const float tv[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) { const int ch = c4 * 4 + e; atomicAdd(dbias + (bias_mod > 0 ? ch % bias_mod : ch), tv[e]); }
            }
            __syncthreads();
        }
    }
}

// even h, w; c % 4 == 0; one thread = one 2x2 window x 4 channels
__global__ void __launch_bounds__(256)
maxpool_fwd_vec_kernel(const float* __restrict__ x, float* __restrict__ y, int n, int oh, int ow, int c4n, int xp, int xo, int yp, int yo) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    const long long total = (long long)n * oh * ow * c4n;
    if (i >= total) return;
    const int c4 = (int)(i % c4n);
    long long t = i / c4n;
    const int ox = (int)(t % ow); t /= ow;
    const int oy = (int)(t % oh);
    const long long nn = t / oh;
    const long long r0 = ((nn * 2 * oh + 2 * oy) * (2 * ow) + 2 * ox);
    const float* b = x + xo + c4 * 4;
    const float4 v00 = __ldg(reinterpret_cast<const float4*>(b + r0 * xp)), v01 = __ldg(reinterpret_cast<const float4*>(b + (r0 + 1) * xp));
    const float4 v10 = __ldg(reinterpret_cast<const float4*>(b + (r0 + 2 * ow) * xp)), v11 = __ldg(reinterpret_cast<const float4*>(b + (r0 + 2 * ow + 1) * xp));
    float4 m;
    m.x = fmaxf(fmaxf(v00.x, v01.x), fmaxf(v10.x, v11.x)); m.y = fmaxf(fmaxf(v00.y, v01.y), fmaxf(v10.y, v11.y));
    m.z = fmaxf(fmaxf(v00.z, v01.z), fmaxf(v10.z, v11.z)); m.w = fmaxf(fmaxf(v00.w, v01.w), fmaxf(v10.w, v11.w));
    *reinterpret_cast<float4*>(y + (((nn * oh + oy) * ow + ox)) * yp + yo + c4 * 4) = m;
}

__device__ __forceinline__ void pool_route(float a, float b, float c, float d, float g, float& ga, float& gb, float& gc, float& gd) {
    // first maximum in scan order (a, b, c, d) takes the gradient
    ga = gb = gc = gd = 0.f;
    float m = a; int arg = 0;
    if (b > m) { m = b; arg = 1; }
    if (c > m) { m = c; arg = 2; }
    if (d > m) { m = d; arg = 3; }
    if (arg == 0) ga = g; else if (arg == 1) gb = g; else if (arg == 2) gc = g; else gd = g;
}

__global__ void __launch_bounds__(256)
maxpool_bwd_vec_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ add, float* __restrict__ dx,
                       int n, int oh, int ow, int c4n, int xp, int xo, int dyp, int dyo, int addp, int addo, int dxp, int dxo) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    const long long total = (long long)n * oh * ow * c4n;
    if (i >= total) return;
    const int c4 = (int)(i % c4n);
    long long t = i / c4n;
    const int ox = (int)(t % ow); t /= ow;
    const int oy = (int)(t % oh);
    const long long nn = t / oh;
    const long long r[4] = {(nn * 2 * oh + 2 * oy) * (2 * ow) + 2 * ox, (nn * 2 * oh + 2 * oy) * (2 * ow) + 2 * ox + 1,
                            (nn * 2 * oh + 2 * oy + 1) * (2 * ow) + 2 * ox, (nn * 2 * oh + 2 * oy + 1) * (2 * ow) + 2 * ox + 1};
    float4 v[4], g[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = __ldg(reinterpret_cast<const float4*>(x + r[k] * xp + xo + c4 * 4));
    const float4 gy = __ldg(reinterpret_cast<const float4*>(dy + ((nn * oh + oy) * ow + ox) * dyp + dyo + c4 * 4));
    pool_route(v[0].x, v[1].x, v[2].x, v[3].x, gy.x, g[0].x, g[1].x, g[2].x, g[3].x);
    pool_route(v[0].y, v[1].y, v[2].y, v[3].y, gy.y, g[0].y, g[1].y, g[2].y, g[3].y);
    pool_route(v[0].z, v[1].z, v[2].z, v[3].z, gy.z, g[0].z, g[1].z, g[2].z, g[3].z);
    pool_route(v[0].w, v[1].w, v[2].w, v[3].w, gy.w, g[0].w, g[1].w, g[2].w, g[3].w);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (add) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(add + r[k] * addp + addo + c4 * 4));
            g[k].x += a.x; g[k].y += a.y; g[k].z += a.z; g[k].w += a.w;
        }
        *reinterpret_cast<float4*>(dx + r[k] * dxp + dxo + c4 * 4) = g[k];
    }
}

// Max-pool backward fused with the backward of the activation that produced x and with the bias gradient of that layer:
// dx = (route(dy) + add) * act'(x), dbias[c] += sum_pixels dx. Saves one read-modify-write pass over the largest gradient tensors
// of the step (12 B per element). Grid-stride with a fixed channel quad per thread (gridDim * 256 is a multiple of c4n), per-block
// shared-memory reduction, c atomics per block.
// IDX = int when every pixel index fits 31 bits (always the case in the training step): the 64-bit divisions of the index
// decomposition were a third of the kernel's instructions.
template <typename IDX>
__global__ void __launch_bounds__(256)
maxpool_act_bwd_bias_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ add, float* __restrict__ dx,
                            float* __restrict__ dbias, int n, int oh, int ow, int c4n, int xp, int xo, int dyp, int dyo, int addp, int addo,
                            int dxp, int dxo, int act, float alpha) {
    __shared__ float4 red[256];
    const IDX total = (IDX)n * oh * ow * c4n;
    const int c4 = threadIdx.x % c4n;                 // 256 % c4n == 0 (checked by the launcher)
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    // pixel index advanced incrementally: one decomposition per thread, then additions (stride = gridDim * 256 / c4n pixels)
    const IDX pstride = (IDX)gridDim.x * (256 / c4n);
    IDX pix = (IDX)blockIdx.x * (256 / c4n) + threadIdx.x / c4n;
    const IDX npix = total / c4n;
    const int sx = (int)(pstride % ow), sy = (int)((pstride / ow) % oh);
    const IDX sn = pstride / ((IDX)ow * oh);
    int ox = (int)(pix % ow), oy = (int)((pix / ow) % oh);
    IDX nn = pix / ((IDX)ow * oh);
    for (; pix < npix; pix += pstride) {
        const long long r0 = ((long long)nn * 2 * oh + 2 * oy) * (2 * ow) + 2 * ox;
        const long long r[4] = {r0, r0 + 1, r0 + 2 * ow, r0 + 2 * ow + 1};
        float4 v[4], g[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = __ldg(reinterpret_cast<const float4*>(x + r[k] * xp + xo + c4 * 4));
        const float4 gy = __ldg(reinterpret_cast<const float4*>(dy + (((long long)nn * oh + oy) * ow + ox) * dyp + dyo + c4 * 4));
        pool_route(v[0].x, v[1].x, v[2].x, v[3].x, gy.x, g[0].x, g[1].x, g[2].x, g[3].x);
        pool_route(v[0].y, v[1].y, v[2].y, v[3].y, gy.y, g[0].y, g[1].y, g[2].y, g[3].y);
        pool_route(v[0].z, v[1].z, v[2].z, v[3].z, gy.z, g[0].z, g[1].z, g[2].z, g[3].z);
        pool_route(v[0].w, v[1].w, v[2].w, v[3].w, gy.w, g[0].w, g[1].w, g[2].w, g[3].w);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (add) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(add + r[k] * addp + addo + c4 * 4));
                g[k].x += a.x; g[k].y += a.y; g[k].z += a.z; g[k].w += a.w;
            }
            g[k].x *= act_grad_from_out(v[k].x, act, alpha); g[k].y *= act_grad_from_out(v[k].y, act, alpha);
            g[k].z *= act_grad_from_out(v[k].z, act, alpha); g[k].w *= act_grad_from_out(v[k].w, act, alpha);
            *reinterpret_cast<float4*>(dx + r[k] * dxp + dxo + c4 * 4) = g[k];
            s.x += g[k].x; s.y += g[k].y; s.z += g[k].z; s.w += g[k].w;
        }
        ox += sx; oy += sy; nn += sn;
        if (ox >= ow) { ox -= ow; ++oy; }
        if (oy >= oh) { oy -= oh; ++nn; }
    }
    if (!dbias) return;
    red[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x < c4n) {
        float4 t = red[threadIdx.x];
        for (int j = threadIdx.x + c4n; j < 256; j += c4n) { const float4 u = red[j]; t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w; }
        atomicAdd(dbias + c4 * 4, t.x); atomicAdd(dbias + c4 * 4 + 1, t.y); atomicAdd(dbias + c4 * 4 + 2, t.z); atomicAdd(dbias + c4 * 4 + 3, t.w);
    }
}

// Backward of (activation + 2x2 max-pool) from the code bytes written by ni_conv2d_pool2_fwd: dx[window position == code & 3] =
// dy * (code & 4 ? 1 : slope0), zero elsewhere; dbias[c] += sum. Reads 5 B and writes 16 B per pooled element (the float path reads 20).
__global__ void __launch_bounds__(256)
maxpool_code_bwd_bias_kernel(const unsigned char* __restrict__ code, const float* __restrict__ dy, float* __restrict__ dx,
                             float* __restrict__ dbias, int n, int oh, int ow, int c4n, float slope0) {
    __shared__ float4 red[256];
    const long long npix = (long long)n * oh * ow;
    const int c4 = threadIdx.x % c4n, C = c4n * 4;
    const long long pstride = (long long)gridDim.x * (256 / c4n);
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (long long pix = (long long)blockIdx.x * (256 / c4n) + threadIdx.x / c4n; pix < npix; pix += pstride) {
        const int ox = (int)(pix % ow), oy = (int)((pix / ow) % oh);
        const long long nn = pix / ((long long)ow * oh);
        const float4 g = __ldg(reinterpret_cast<const float4*>(dy + pix * C + c4 * 4));
        const unsigned int cd = __ldg(reinterpret_cast<const unsigned int*>(code + pix * C + c4 * 4));
        const float gv[4] = {g.x, g.y, g.z, g.w};
        float o[4][4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const unsigned int b = (cd >> (8 * e)) & 0xffu;
            const float v = gv[e] * ((b & 4u) ? 1.f : slope0);
#pragma unroll
            for (int k = 0; k < 4; ++k) o[k][e] = (b & 3u) == (unsigned)k ? v : 0.f;
            if (e == 0) s.x += v; else if (e == 1) s.y += v; else if (e == 2) s.z += v; else s.w += v;
        }
        const long long r0 = ((nn * 2 * oh + 2 * oy) * (2LL * ow) + 2 * ox);
        const long long r[4] = {r0, r0 + 1, r0 + 2 * ow, r0 + 2 * ow + 1};
#pragma unroll
        for (int k = 0; k < 4; ++k) *reinterpret_cast<float4*>(dx + r[k] * C + c4 * 4) = make_float4(o[k][0], o[k][1], o[k][2], o[k][3]);
    }
    if (!dbias) return;
    red[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x < c4n) {
        float4 t = red[threadIdx.x];
        for (int j = threadIdx.x + c4n; j < 256; j += c4n) { const float4 u = red[j]; t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w; }
        atomicAdd(dbias + c4 * 4, t.x); atomicAdd(dbias + c4 * 4 + 1, t.y); atomicAdd(dbias + c4 * 4 + 2, t.z); atomicAdd(dbias + c4 * 4 + 3, t.w);
    }
}

// ------------------------------------------------------------------ global average pooling (n,h,w,c) <-> (n,c)
__global__ void gap_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int hw, int c) {
    const int n = blockIdx.x;
    for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
        float s = 0.f;
        for (int p = 0; p < hw; ++p) s += x[((long long)n * hw + p) * c + ch];
        y[(long long)n * c + ch] = s / (float)hw;
    }
}
__global__ void gap_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, long long total, int hw, int c) {
    const long long i = (long long)blockIdx.x * kT + threadIdx.x;
    if (i >= total) return;
    const int ch = (int)(i % c);
    const long long n = i / ((long long)hw * c);
    dx[i] = dy[n * c + ch] / (float)hw;
}

// ------------------------------------------------------------------ softmax + Keras SparseCategoricalCrossentropy
// Forward: probs = softmax(logits); loss_i = -log(q_l / sum_k q_k), q = clip(p, 1e-7, 1-1e-7)  (models/forensics.py:94,
// Keras backend.sparse_categorical_crossentropy with from_logits=False in eager mode). loss_sum accumulates sum_i loss_i.
// Backward (fused, optional): dlogits = d(mean loss)/dlogits * gscale.
__global__ void softmax_ce_kernel(const float* __restrict__ logits, const int* __restrict__ labels, float* __restrict__ probs,
                                  float* __restrict__ loss_sum, float* __restrict__ dlogits, int m, int c, float gscale) {
    const int i = blockIdx.x * kT + threadIdx.x;
    float li = 0.f;
    if (i < m) {
        const float* z = logits + (long long)i * c;
        float mx = -INFINITY;
        for (int k = 0; k < c; ++k) mx = fmaxf(mx, z[k]);
        float se = 0.f;
        for (int k = 0; k < c; ++k) se += expf(z[k] - mx);
        const float eps = 1e-7f;
        float sq = 0.f;
        for (int k = 0; k < c; ++k) {
            const float pk = expf(z[k] - mx) / se;
            if (probs) probs[(long long)i * c + k] = pk;
            sq += fminf(fmaxf(pk, eps), 1.f - eps);
        }
        if (labels) {
            const int l = labels[i];
            const float pl = expf(z[l] - mx) / se;
            const float ql = fminf(fmaxf(pl, eps), 1.f - eps);
            li = -logf(ql / sq);
            if (dlogits) {
                // g_k = dL/dp_k = (q_k/sq - 1[k==l]) / p_k inside the clip range, else 0
                float dot = 0.f;
                for (int k = 0; k < c; ++k) {
                    const float pk = expf(z[k] - mx) / se;
                    const float qk = fminf(fmaxf(pk, eps), 1.f - eps);
                    const float gk = (pk >= eps && pk <= 1.f - eps) ? (qk / sq - (k == l ? 1.f : 0.f)) / pk : 0.f;
                    dot += gk * pk;
                }
                for (int k = 0; k < c; ++k) {
                    const float pk = expf(z[k] - mx) / se;
                    const float qk = fminf(fmaxf(pk, eps), 1.f - eps);
                    const float gk = (pk >= eps && pk <= 1.f - eps) ? (qk / sq - (k == l ? 1.f : 0.f)) / pk : 0.f;
                    dlogits[(long long)i * c + k] = pk * (gk - dot) * gscale;
                }
            }
        }
    }
    if (loss_sum) {
        // block reduce
        __shared__ float red[kT / 32];
        float v = li;
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
        __syncthreads();
        if (threadIdx.x == 0) {
            float s = 0.f;
            for (int k = 0; k < kT / 32; ++k) s += red[k];
            atomicAdd(loss_sum, s);
        }
    }
}

// ------------------------------------------------------------------ image losses (helpers/tf_helpers.py:31-36)
// kind 0: L2 = mean((255a-255b)^2); kind 1: L1 = mean(|255a-255b|). acc[0] += sum of per-element terms.
__global__ void image_loss_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ acc,
                                  long long n, int kind) {
    float s = 0.f;
    for (long long i = (long long)blockIdx.x * kT + threadIdx.x; i < n; i += (long long)gridDim.x * kT) {
        const float d = 255.f * a[i] - 255.f * b[i];
        s += kind == 0 ? d * d : fabsf(d);
    }
    __shared__ float red[kT / 32];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int k = 0; k < kT / 32; ++k) t += red[k];
        atomicAdd(acc, t);
    }
}
// da (+)= scale * d loss / d a
__global__ void image_loss_grad_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ da,
                                       long long n, int kind, float scale, int accumulate) {
    const long long i = (long long)blockIdx.x * kT + threadIdx.x;
    if (i >= n) return;
    const float d = 255.f * a[i] - 255.f * b[i];
    const float g = (kind == 0 ? 2.f * 255.f * d : (d > 0.f ? 255.f : (d < 0.f ? -255.f : 0.f))) * scale / (float)n;
    da[i] = accumulate ? da[i] + g : g;
}

// ------------------------------------------------------------------ constrained residual filter (models/layers.py:45-53)
// k: trainable (5,5,3,3) HWIO. nf = s*k*(1-ind)/sum_{a,b,ci}(k*(1-ind))[co] - s*ind, ind = centre tap on the channel diagonal.
__global__ void constrained_filter_fwd_kernel(const float* __restrict__ k, float* __restrict__ nf, int ks, int ch, float strength) {
    __shared__ float df[16];
    const int total = ks * ks * ch * ch;
    if (threadIdx.x < ch) {
        float s = 0.f;
        for (int i = 0; i < ks * ks * ch; ++i) {
            const int ci = i % ch, tap = i / ch;
            const bool centre = (tap == (ks / 2) * ks + ks / 2) && ci == (int)threadIdx.x;
            if (!centre) s += k[i * ch + threadIdx.x];
        }
        df[threadIdx.x] = s;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int co = i % ch, ci = (i / ch) % ch, tap = i / (ch * ch);
        const bool centre = (tap == (ks / 2) * ks + ks / 2) && ci == co;
        nf[i] = centre ? -strength : strength * k[i] / df[co];
    }
}
// dk_j = m_j * s * (g_j / D - (sum_i g_i k_i m_i) / D^2), per output channel
__global__ void constrained_filter_bwd_kernel(const float* __restrict__ k, const float* __restrict__ dnf, float* __restrict__ dk,
                                              int ks, int ch, float strength) {
    __shared__ float df[16], dot[16];
    const int total = ks * ks * ch * ch;
    if (threadIdx.x < ch) {
        float s = 0.f, d = 0.f;
        for (int i = 0; i < ks * ks * ch; ++i) {
            const int ci = i % ch, tap = i / ch;
            const bool centre = (tap == (ks / 2) * ks + ks / 2) && ci == (int)threadIdx.x;
            if (!centre) { s += k[i * ch + threadIdx.x]; d += k[i * ch + threadIdx.x] * dnf[i * ch + threadIdx.x]; }
        }
        df[threadIdx.x] = s; dot[threadIdx.x] = d;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int co = i % ch, ci = (i / ch) % ch, tap = i / (ch * ch);
        const bool centre = (tap == (ks / 2) * ks + ks / 2) && ci == co;
        dk[i] = centre ? 0.f : strength * (dnf[i] / df[co] - dot[co] / (df[co] * df[co]));
    }
}

// ------------------------------------------------------------------ mirrored-pad gradient folding
// dpad: (n, h+2p, w+2p, c) gradient w.r.t. the padded tensor; dx[pixel] = sum of dpad over all padded aliases.
__global__ void pad_fold_kernel(const float* __restrict__ dpad, float* __restrict__ dx, int n, int h, int w, int c, int pad,
                                int mode, int accumulate) {
    const long long i = (long long)blockIdx.x * kT + threadIdx.x;
    const long long total = (long long)n * h * w * c;
    if (i >= total) return;
    const int ch = (int)(i % c);
    long long t = i / c;
    const int px = (int)(t % w); t /= w;
    const int py = (int)(t % h);
    const int nn = (int)(t / h);
    int uy[3], ux[3], ny = 0, nx = 0;
    uy[ny++] = py; ux[nx++] = px;
    if (mode == NI_PAD_SYMMETRIC) {
        if (py <= pad - 1) uy[ny++] = -py - 1;
        if (py >= h - pad) uy[ny++] = 2 * h - 1 - py;
        if (px <= pad - 1) ux[nx++] = -px - 1;
        if (px >= w - pad) ux[nx++] = 2 * w - 1 - px;
    } else {
        if (py >= 1 && py <= pad) uy[ny++] = -py;
        if (py <= h - 2 && 2 * (h - 1) - py <= h - 1 + pad) uy[ny++] = 2 * (h - 1) - py;
        if (px >= 1 && px <= pad) ux[nx++] = -px;
        if (px <= w - 2 && 2 * (w - 1) - px <= w - 1 + pad) ux[nx++] = 2 * (w - 1) - px;
    }
    const int ph = h + 2 * pad, pw = w + 2 * pad;
    float s = 0.f;
    for (int a = 0; a < ny; ++a)
        for (int b = 0; b < nx; ++b) s += dpad[(((long long)nn * ph + uy[a] + pad) * pw + ux[b] + pad) * c + ch];
    dx[i] = accumulate ? dx[i] + s : s;
}

// ------------------------------------------------------------------ fused Keras Adam on a flat parameter buffer
// m += (g-m)(1-b1); v += (g^2-v)(1-b2); p -= lr_t * m / (sqrt(v)+eps), lr_t = lr*sqrt(1-b2^t)/(1-b1^t) (host-computed).
// gscale multiplies the gradient first (1/world_size after the sum all-reduce). Non-finite gradients raise *flag and
// leave the parameter untouched (the reference raises on NaN gradients, workflows/manipulation_classification.py:281).
// lr_dev != nullptr: the bias-corrected step size is read from device memory (CUDA-graph replays: the host refreshes it through
// a captured pinned-memory copy instead of a baked-in kernel argument).
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            long long n, float lr_t, const float* __restrict__ lr_dev, float b1, float b2, float eps, float gscale,
                            int* __restrict__ flag) {
    const long long i = (long long)blockIdx.x * kT + threadIdx.x;
    if (i >= n) return;
    if (lr_dev) lr_t = __ldg(lr_dev);
    const float gi = g[i] * gscale;
    if (!isfinite(gi)) { if (flag) atomicOr(flag, 1); return; }
    const float mi = m[i] + (gi - m[i]) * (1.f - b1);
    const float vi = v[i] + (gi * gi - v[i]) * (1.f - b2);
    m[i] = mi; v[i] = vi;
    p[i] = p[i] - lr_t * mi / (sqrtf(vi) + eps);
}

__global__ void fill_kernel(float* __restrict__ p, float v, long long n) {
    const long long i = (long long)blockIdx.x * kT + threadIdx.x;
    if (i < n) p[i] = v;
}
// y = a*x + b (elementwise), optional clip to [0,1]
__global__ void affine_kernel(const float* __restrict__ x, float* __restrict__ y, float a, float b, int clip, long long n) {
    const long long i = (long long)blockIdx.x * kT + threadIdx.x;
    if (i < n) { const float v = fmaf(a, x[i], b); y[i] = clip ? ni_clamp01(v) : v; }
}

}  // namespace

extern "C" int ni_maxpool2_fwd(const float* x, float* y, int n, int h, int w, int c, int same, int x_pitch, int x_coff,
                               int y_pitch, int y_coff, cudaStream_t st) {
    NI_REQUIRE(x && y && n >= 0 && h > 0 && w > 0 && c > 0, "ni_maxpool2_fwd: invalid arguments");
    const int oh = same ? (h + 1) / 2 : h / 2, ow = same ? (w + 1) / 2 : w / 2;
    NI_REQUIRE(oh > 0 && ow > 0, "ni_maxpool2_fwd: input smaller than the pooling window");
    if (n == 0) return NI_OK;
    if (!(h & 1) && !(w & 1) && !(c & 3) && !(x_pitch & 3) && !(x_coff & 3) && !(y_pitch & 3) && !(y_coff & 3)) {
        maxpool_fwd_vec_kernel<<<grid_for((long long)n * oh * ow * (c / 4)), kT, 0, st>>>(x, y, n, oh, ow, c / 4, x_pitch, x_coff, y_pitch, y_coff);
        NI_LAUNCH_CHECK(); NI_COUNT_LAUNCH(1);
        return NI_OK;
    }
    maxpool_fwd_kernel<<<grid_for((long long)n * oh * ow * c), kT, 0, st>>>(x, y, n, h, w, c, oh, ow, x_pitch, x_coff, y_pitch, y_coff);
    NI_LAUNCH_CHECK(); NI_COUNT_LAUNCH(1);
    return NI_OK;
}

extern "C" int ni_maxpool2_bwd(const float* x, const float* dy, const float* add, float* dx, int n, int h, int w, int c,
                               int same, int x_pitch, int x_coff, int dy_pitch, int dy_coff, int add_pitch, int add_coff,
                               int dx_pitch, int dx_coff, cudaStream_t st) {
    NI_REQUIRE(x && dy && dx && n >= 0 && h > 0 && w > 0 && c > 0, "ni_maxpool2_bwd: invalid arguments");
    const int oh = same ? (h + 1) / 2 : h / 2, ow = same ? (w + 1) / 2 : w / 2;
    if (n == 0) return NI_OK;
    if (!(h & 1) && !(w & 1) && !(c & 3) && !(x_pitch & 3) && !(x_coff & 3) && !(dy_pitch & 3) && !(dy_coff & 3) && !(dx_pitch & 3) &&
        !(dx_coff & 3) && (!add || (!(add_pitch & 3) && !(add_coff & 3)))) {
        maxpool_bwd_vec_kernel<<<grid_for((long long)n * oh * ow * (c / 4)), kT, 0, st>>>(x, dy, add, dx, n, oh, ow, c / 4, x_pitch, x_coff, dy_pitch,
                                                                                      dy_coff, add_pitch, add_coff, dx_pitch, dx_coff);
        NI_LAUNCH_CHECK(); NI_COUNT_LAUNCH(1);
        return NI_OK;
    }
    maxpool_bwd_kernel<<<grid_for((long long)n * h * w * c), kT, 0, st>>>(x, dy, add, dx, n, h, w, c, oh, ow, x_pitch, x_coff,
                                                                       dy_pitch, dy_coff, add_pitch, add_coff, dx_pitch, dx_coff);
    NI_LAUNCH_CHECK(); NI_COUNT_LAUNCH(1);
    return NI_OK;
}

extern "C" int ni_act_bwd_bias(const float* y, float* dy, float* dbias, int n, int h, int w, int c, int y_pitch, int y_coff,
                               int y_mode, int dy_pitch, int dy_coff, int dy_mode, int act, float alpha, int bias_mod, cudaStream_t st);

// ni_maxpool2_bwd followed by ni_act_bwd_bias on the pooled layer's own output x (conv -> activation -> pool chains of FAN and the
// U-Net encoder), as ONE pass where the layout allows it; dbias (c floats, may be null) is overwritten.
// code / dy: (n, oh, ow, c) pooled layout; dx: (n, 2 oh, 2 ow, c) plain; act: leaky-relu (slope alpha) or relu. dbias is overwritten.
extern "C" int ni_maxpool2_code_bwd_bias(const unsigned char* code, const float* dy, float* dx, float* dbias, int n, int oh, int ow, int c,
                                         int act, float alpha, cudaStream_t st) {
    NI_REQUIRE(code && dy && dx && n >= 0 && oh > 0 && ow > 0 && c > 0 && (c % 4) == 0 && 256 % (c / 4) == 0,
               "ni_maxpool2_code_bwd_bias: invalid arguments (c must be a multiple of 4 with 256 %% (c/4) == 0)");
    NI_REQUIRE(act == NI_ACT_LEAKY_RELU || act == NI_ACT_RELU, "ni_maxpool2_code_bwd_bias: activation must be leaky-relu or relu");
    if (dbias) NI_CUDA(cudaMemsetAsync(dbias, 0, sizeof(float) * c, st));
    const long long npix = (long long)n * oh * ow;
    if (npix == 0) return NI_OK;
    const int c4n = c / 4, ppb = 256 / c4n;
    long long blocks = (npix + ppb - 1) / ppb;
    const long long cap = 16LL * ni_num_sms();
    if (blocks > cap) blocks = cap;
    maxpool_code_bwd_bias_kernel<<<(unsigned)blocks, 256, 0, st>>>(code, dy, dx, dbias, n, oh, ow, c4n, act == NI_ACT_RELU ? 0.f : alpha);
    NI_LAUNCH_CHECK(); NI_COUNT_LAUNCH(1);
    return NI_OK;
}

extern "C" int ni_maxpool2_act_bwd_bias(const float* x, const float* dy, const float* add, float* dx, float* dbias, int n, int h, int w, int c,
                                        int same, int x_pitch, int x_coff, int dy_pitch, int dy_coff, int add_pitch, int add_coff, int dx_pitch,
                                        int dx_coff, int act, float alpha, cudaStream_t st) {
    NI_REQUIRE(x && dy && dx && n >= 0 && h > 0 && w > 0 && c > 0, "ni_maxpool2_act_bwd_bias: invalid arguments");
    if (n == 0) return NI_OK;
    const int oh = same ? (h + 1) / 2 : h / 2, ow = same ? (w + 1) / 2 : w / 2;
    const int c4n = c / 4;
    const bool vec = !(h & 1) && !(w & 1) && !(c & 3) && c4n > 0 && (256 % c4n) == 0 && !(x_pitch & 3) && !(x_coff & 3) && !(dy_pitch & 3) &&
                     !(dy_coff & 3) && !(dx_pitch & 3) && !(dx_coff & 3) && (!add || (!(add_pitch & 3) && !(add_coff & 3)));
    if (!vec) {
        int rc = ni_maxpool2_bwd(x, dy, add, dx, n, h, w, c, same, x_pitch, x_coff, dy_pitch, dy_coff, add_pitch, add_coff, dx_pitch, dx_coff, st);
        if (rc) return rc;
        return ni_act_bwd_bias(x, dx, dbias, n, h, w, c, x_pitch, x_coff, NI_MODE_PLAIN, dx_pitch, dx_coff, NI_MODE_PLAIN, act, alpha, 0, st);
    }
    if (dbias) NI_CUDA(cudaMemsetAsync(dbias, 0, sizeof(float) * c, st));
    const long long total = (long long)n * oh * ow * c4n;
    long long blocks = (total + 255) / 256;
    const long long cap = 16LL * ni_num_sms();
    if (blocks > cap) blocks = cap;
    if (total + 16LL * 148 * 256 < 2147483647LL)
        maxpool_act_bwd_bias_kernel<int><<<(unsigned)blocks, 256, 0, st>>>(x, dy, add, dx, dbias, n, oh, ow, c4n, x_pitch, x_coff, dy_pitch, dy_coff,
                                                                          add_pitch, add_coff, dx_pitch, dx_coff, act, alpha);
    else
        maxpool_act_bwd_bias_kernel<long long><<<(unsigned)blocks, 256, 0, st>>>(x, dy, add, dx, dbias, n, oh, ow, c4n, x_pitch, x_coff, dy_pitch,
                                                                                dy_coff, add_pitch, add_coff, dx_pitch, dx_coff, act, alpha);
    NI_LAUNCH_CHECK(); NI_COUNT_LAUNCH(1);
    return NI_OK;
}

// Logical (n,h,w,c) views of the layer output y and its gradient dy (pitch/coff/mode as in ni_conv_desc):
// dy <- dy * act'(y) in place; dbias[c] (or dbias[c % bias_mod] when bias_mod > 0) = sum over pixels. dbias may be null.
extern "C" int ni_act_bwd_bias(const float* y, float* dy, float* dbias, int n, int h, int w, int c, int y_pitch, int y_coff,
                               int y_mode, int dy_pitch, int dy_coff, int dy_mode, int act, float alpha, int bias_mod,
                               cudaStream_t st) {
    NI_REQUIRE(dy && n >= 0 && h > 0 && w > 0 && c > 0, "ni_act_bwd_bias: invalid arguments");
    NI_REQUIRE(y || act == NI_ACT_NONE || act == NI_ACT_CLIP01, "ni_act_bwd_bias: activation backward needs the layer output");
    const int nb = bias_mod > 0 ? bias_mod : c;
    if (dbias) NI_CUDA(cudaMemsetAsync(dbias, 0, sizeof(float) * nb, st));
    const long long npix = (long long)n * h * w;
    if (npix == 0) return NI_OK;
    if (!dbias && (act == NI_ACT_NONE || act == NI_ACT_CLIP01)) return NI_OK;
    // Bias gradient of a transposed convolution (1x1 conv + depth_to_space(2), no activation): the logical (n,h,w,4F) gradient is the
    // physical (n,2h,2w,F) buffer and the bias index is the physical channel, so the sum over logical pixels and sub-pixel blocks is
    // a plain per-channel sum over the physical tensor: take the vectorised streaming path (the scalar depth_to_space path ran at
    // 10 % of the HBM bandwidth: 1.9 ms per step for 1.2 GB).
    if (dy_mode == NI_MODE_BLOCK2 && (act == NI_ACT_NONE || act == NI_ACT_CLIP01) && dbias && !(c & 3) && bias_mod == c / 4 && !((c / 4) & 3) &&
        !(dy_pitch & 3) && !(dy_coff & 3)) {
        return ni_act_bwd_bias(nullptr, dy, dbias, n, 2 * h, 2 * w, c / 4, dy_pitch, dy_coff, NI_MODE_PLAIN, dy_pitch, dy_coff, NI_MODE_PLAIN, NI_ACT_NONE, alpha,
                               0, st);
    }
    if (y_mode == NI_MODE_PLAIN && dy_mode == NI_MODE_PLAIN && !(c & 3) && !(dy_pitch & 3) && !(dy_coff & 3) &&
        (!y || (!(y_pitch & 3) && !(y_coff & 3)))) {
        int nx = c / 4; if (nx > 64) nx = 64;
        while (256 % nx) --nx;                         // nx divides 256
        const int ny = 256 / nx;
        const int cblocks = ni_cdiv(c / 4, nx);
        long long chunks = (8LL * ni_num_sms() + cblocks - 1) / cblocks;
        if (chunks > (npix + 4 * ny - 1) / (4 * ny)) chunks = (npix + 4 * ny - 1) / (4 * ny);
        if (chunks < 1) chunks = 1;
        if (chunks > 65535) chunks = 65535;
        const int ppb = (int)((npix + chunks - 1) / chunks);
        dim3 grid(cblocks, ni_cdiv(npix, ppb));
        act_bwd_bias_vec_kernel<<<grid, 256, sizeof(float4) * 256, st>>>(y, dy, dbias, npix, c, y_pitch, y_coff, dy_pitch, dy_coff, act, alpha, ppb,
                                                                      bias_mod, nx, ny);
        NI_LAUNCH_CHECK(); NI_COUNT_LAUNCH(1);
        return NI_OK;
    }
    const int cgroups = ni_cdiv(c, 32);
    long long chunks = (4LL * ni_num_sms() + cgroups - 1) / cgroups;
    if (chunks > (npix + 63) / 64) chunks = (npix + 63) / 64;
    if (chunks < 1) chunks = 1;
    if (chunks > 65535) chunks = 65535;
    const int ppb = (int)((npix + chunks - 1) / chunks);
    dim3 grid(cgroups, ni_cdiv(npix, ppb));
    TensorView yv{h, w, c, y_pitch, y_coff, y_mode}, dv{h, w, c, dy_pitch, dy_coff, dy_mode};
    act_bwd_bias_kernel<<<grid, dim3(32, 8), 0, st>>>(y, dy, dbias, npix, yv, dv, act, alpha, ppb, bias_mod);
    NI_LAUNCH_CHECK(); NI_COUNT_LAUNCH(1);
    return NI_OK;
}

extern "C" int ni_gap_fwd(const float* x, float* y, int n, int hw, int c, cudaStream_t st) {
    NI_REQUIRE(x && y && n >= 0 && hw > 0 && c > 0, "ni_gap_fwd: invalid arguments");
    if (n == 0) return NI_OK;
    gap_fwd_kernel<<<n, 256, 0, st>>>(x, y, hw, c);
    NI_LAUNCH_CHECK(); NI_COUNT_LAUNCH(1);
    return NI_OK;
}
extern "C" int ni_gap_bwd(const float* dy, float* dx, int n, int hw, int c, cudaStream_t st) {
    NI_REQUIRE(dy && dx && n >= 0 && hw > 0 && c > 0, "ni_gap_bwd: invalid arguments");
    if (n == 0) return NI_OK;
    const long long total = (long long)n * hw * c;
    gap_bwd_kernel<<<grid_for(total), kT, 0, st>>>(dy, dx, total, hw, c);
    NI_LAUNCH_CHECK(); NI_COUNT_LAUNCH(1);
    return NI_OK;
}

// probs / loss_sum / dlogits / labels may each be null. *loss_sum accumulates the SUM of per-sample losses
// (caller zeroes it and divides by m); dlogits = gscale * d(mean loss)/dlogits requires gscale to include 1/m.
extern "C" int ni_softmax_ce(const float* logits, const int* labels, float* probs, float* loss_sum, float* dlogits, int m,
                             int c, float gscale, cudaStream_t st) {
    NI_REQUIRE(logits && m >= 0 && c > 0, "ni_softmax_ce: invalid arguments");
    NI_REQUIRE(labels || (!loss_sum && !dlogits), "ni_softmax_ce: loss / gradient need labels");
    if (m == 0) return NI_OK;
    softmax_ce_kernel<<<ni_cdiv(m, kT), kT, 0, st>>>(logits, labels, probs, loss_sum, dlogits, m, c, gscale);
    NI_LAUNCH_CHECK(); NI_COUNT_LAUNCH(1);
    return NI_OK;
}

// ---------------------------------------------------------------- stride-2 5x5 convolution as a 3x3 convolution over space_to_depth(2)
// A SAME, stride-2, 5x5 convolution on an even-sized input (models/compression.py:221-222,237: the DCN encoder's down-sampling layers)
// reads input rows 2o + a - 1, a = 0..4, for output row o: in the space_to_depth(2) domain (block i = rows 2i, 2i + 1) these are
// block o - 1 (second row), block o (both rows), block o + 1 (both rows) -- a SAME 3x3 stride-1 convolution over 4 x cin channels with
// the 5x5 taps scattered into a zero-padded 6x6 grid. That form runs on the tcgen05 implicit-GEMM path (stride 1, cin' % 32 == 0);
// the price is 36 / 25 of the MACs. These kernels move the activations and the weights between the two forms.
//   xs[n, i, j, (di*2 + dj)*C + c] = x[n, 2i + di, 2j + dj, c]                       (tf.nn.space_to_depth order)
//   w3[bi, bj, (di*2 + dj)*C + c, f] = w5[2bi + di - 1, 2bj + dj - 1, c, f] or 0
__global__ void s2d2_kernel(const float* __restrict__ x, float* __restrict__ y, long long total4, int h2, int w2, int c4, int inverse,
                            int accumulate) {
    // one thread per 4 channels of one (n, i, j, block offset) cell; x is the plain (n, 2h2, 2w2, C) tensor, y the (n, h2, w2, 4C) one
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total4) return;
    const int cq = (int)(t % c4);
    long long r = t / c4;
    const int blk = (int)(r % 4); r /= 4;
    const int j = (int)(r % w2); r /= w2;
    const int i = (int)(r % h2);
    const long long n = r / h2;
    const int C = c4 * 4;
    const long long plain = (((n * 2 * h2 + 2 * i + (blk >> 1)) * (2LL * w2) + 2 * j + (blk & 1)) * C) + cq * 4;
    const long long deep = (((n * h2 + i) * (long long)w2 + j) * 4 + blk) * C + cq * 4;
    if (!inverse) {
        *reinterpret_cast<float4*>(y + deep) = *reinterpret_cast<const float4*>(x + plain);
    } else {        // y (deep) -> x (plain): here `x` is the destination
        float4 v = *reinterpret_cast<const float4*>(y + deep);
        float4* o = reinterpret_cast<float4*>(const_cast<float*>(x) + plain);
        if (accumulate) { const float4 old = *o; v.x += old.x; v.y += old.y; v.z += old.z; v.w += old.w; }
        *o = v;
    }
}
// channel counts that are not multiples of 4 (the 3-channel image end of the DCN encoder): one element per thread, indexed by the
// PLAIN position so that both sides are read / written in runs of c (plain) and c (deep) floats
__global__ void s2d2_scalar_kernel(float* __restrict__ x, float* __restrict__ y, long long total, int h2, int w2, int c, int inverse, int accumulate) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int cc = (int)(t % c);
    long long r = t / c;
    const int X = (int)(r % (2 * w2)); r /= 2 * w2;
    const int Y = (int)(r % (2 * h2));
    const long long n = r / (2 * h2);
    const int blk = (Y & 1) * 2 + (X & 1);
    const long long deep = (((n * h2 + (Y >> 1)) * (long long)w2 + (X >> 1)) * 4 + blk) * c + cc;
    if (!inverse) y[deep] = x[t];
    else x[t] = accumulate ? x[t] + y[deep] : y[deep];
}
// plain (n, 2*h2, 2*w2, c) <-> deep (n, h2, w2, 4c); inverse = 0: deep <- plain, 1: plain (+)= deep.
extern "C" int ni_space_to_depth2(float* plain, float* deep, int n, int h2, int w2, int c, int inverse, int accumulate, cudaStream_t st) {
    NI_REQUIRE(plain && deep && n >= 0 && h2 > 0 && w2 > 0 && c > 0, "ni_space_to_depth2: invalid arguments");
    const long long total4 = (long long)n * h2 * w2 * c;      // = n*h2*w2*4 blocks * c/4 quads
    if (total4 == 0) return NI_OK;
    if (c % 4) {
        s2d2_scalar_kernel<<<ni_cdiv(total4 * 4, kT), kT, 0, st>>>(plain, deep, total4 * 4, h2, w2, c, inverse, accumulate);
        NI_LAUNCH_CHECK(); NI_COUNT_LAUNCH(1);
        return NI_OK;
    }
    s2d2_kernel<<<ni_cdiv(total4, kT), kT, 0, st>>>(plain, deep, total4, h2, w2, c / 4, inverse, accumulate);
    NI_LAUNCH_CHECK(); NI_COUNT_LAUNCH(1);
    return NI_OK;
}
__global__ void s2conv_weights_kernel(const float* __restrict__ w5, float* __restrict__ w3, int cin, int cout, int inverse) {
    if (!inverse) {        // w3 <- scatter(w5), one thread per w3 element
        const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
        const long long total = 9LL * 4 * cin * cout;
        if (t >= total) return;
        const int f = (int)(t % cout);
        long long r = t / cout;
        const int c = (int)(r % cin); r /= cin;
        const int blk = (int)(r % 4); r /= 4;
        const int bj = (int)(r % 3), bi = (int)(r / 3);
        const int a = 2 * bi + (blk >> 1) - 1, b = 2 * bj + (blk & 1) - 1;
        w3[t] = (a >= 0 && a < 5 && b >= 0 && b < 5) ? w5[(((long long)a * 5 + b) * cin + c) * cout + f] : 0.f;
    } else {               // dw5 <- gather(dw3), one thread per w5 element
        const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
        const long long total = 25LL * cin * cout;
        if (t >= total) return;
        const int f = (int)(t % cout);
        long long r = t / cout;
        const int c = (int)(r % cin); r /= cin;
        const int b = (int)(r % 5), a = (int)(r / 5);
        const int bi = (a + 1) >> 1, di = (a + 1) & 1, bj = (b + 1) >> 1, dj = (b + 1) & 1;
        const_cast<float*>(w5)[t] = w3[((((long long)bi * 3 + bj) * 4 + di * 2 + dj) * cin + c) * cout + f];
    }
}
// inverse = 0: w3 (3, 3, 4*cin, cout) <- w5 (5, 5, cin, cout); inverse = 1: w5 <- the 5x5 taps of w3 (weight-gradient way back)
extern "C" int ni_s2conv_weights(float* w5, float* w3, int cin, int cout, int inverse, cudaStream_t st) {
    NI_REQUIRE(w5 && w3 && cin > 0 && cout > 0, "ni_s2conv_weights: invalid arguments");
    const long long total = inverse ? 25LL * cin * cout : 36LL * cin * cout;
    s2conv_weights_kernel<<<ni_cdiv(total, kT), kT, 0, st>>>(w5, w3, cin, cout, inverse);
    NI_LAUNCH_CHECK(); NI_COUNT_LAUNCH(1);
    return NI_OK;
}

// Keras Dropout(rate) in training mode (models/forensics.py:88: after each hidden dense layer when the model is CALLED with
// training=True; the reference's training steps call it without, so the layer is inactive there): y = x * keep / (1 - rate), keep ~
// Bernoulli(1 - rate) from a counter-based generator (splitmix64 of seed and element index; TensorFlow's stream cannot be reproduced).
__global__ void dropout_kernel(const float* __restrict__ x, float* __restrict__ y, long long n, float rate, unsigned long long seed) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(i + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    const float u = (float)(z >> 40) * (1.f / 16777216.f);        // uniform [0, 1)
    y[i] = u >= rate ? x[i] / (1.f - rate) : 0.f;
}
extern "C" int ni_dropout(const float* x, float* y, long long n, float rate, unsigned long long seed, cudaStream_t st) {
    NI_REQUIRE(x && y && n >= 0 && rate >= 0.f && rate < 1.f, "ni_dropout: invalid arguments");
    if (n == 0) return NI_OK;
    dropout_kernel<<<ni_cdiv(n, kT), kT, 0, st>>>(x, y, n, rate, seed);
    NI_LAUNCH_CHECK(); NI_COUNT_LAUNCH(1);
    return NI_OK;
}

// Decisions + confusion matrix of a validation pass on the device (reference training/validation.py:163-203 builds it on the host with
// an n_classes^2 Python loop per batch of 10): pred[i] = argmax_k probs[i, k] (first maximum, like numpy.argmax), conf[label, pred] += 1.
__global__ void confusion_kernel(const float* __restrict__ probs, const int* __restrict__ labels, int* __restrict__ conf,
                                 int* __restrict__ pred, int m, int c) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const float* p = probs + (size_t)i * c;
    int best = 0;
    float bv = p[0];
    for (int k = 1; k < c; ++k) {
        float v = p[k];
        if (v > bv) { bv = v; best = k; }
    }
    if (pred) pred[i] = best;
    if (conf && labels) atomicAdd(conf + labels[i] * c + best, 1);
}
extern "C" int ni_confusion_accumulate(const float* probs, const int* labels, int* conf, int* pred, int m, int c, cudaStream_t st) {
    NI_REQUIRE(probs && m >= 0 && c > 0 && (pred || (conf && labels)), "ni_confusion_accumulate: invalid arguments");
    if (m == 0) return NI_OK;
    confusion_kernel<<<ni_cdiv(m, kT), kT, 0, st>>>(probs, labels, conf, pred, m, c);
    NI_LAUNCH_CHECK(); NI_COUNT_LAUNCH(1);
    return NI_OK;
}

// *acc += sum_i term_i (caller zeroes and divides by n). kind: 0 = L2 (mse of 255-scaled images), 1 = L1.
extern "C" int ni_image_loss(const float* a, const float* b, float* acc, long long n, int kind, cudaStream_t st) {
    NI_REQUIRE(a && b && acc && n >= 0 && (kind == 0 || kind == 1), "ni_image_loss: invalid arguments");
    if (n == 0) return NI_OK;
    int grid = ni_cdiv(n, kT * 8);
    if (grid > 8 * ni_num_sms()) grid = 8 * ni_num_sms();
    image_loss_kernel<<<grid, kT, 0, st>>>(a, b, acc, n, kind);
    NI_LAUNCH_CHECK(); NI_COUNT_LAUNCH(1);
    return NI_OK;
}
extern "C" int ni_image_loss_grad(const float* a, const float* b, float* da, long long n, int kind, float scale, int accumulate,
                                  cudaStream_t st) {
    NI_REQUIRE(a && b && da && n >= 0 && (kind == 0 || kind == 1), "ni_image_loss_grad: invalid arguments");
    if (n == 0) return NI_OK;
    image_loss_grad_kernel<<<grid_for(n), kT, 0, st>>>(a, b, da, n, kind, scale, accumulate);
    NI_LAUNCH_CHECK(); NI_COUNT_LAUNCH(1);
    return NI_OK;
}

extern "C" int ni_constrained_filter_fwd(const float* k, float* nf, int ksize, int channels, float strength, cudaStream_t st) {
    NI_REQUIRE(k && nf && ksize > 0 && (ksize & 1) && channels > 0 && channels <= 16, "ni_constrained_filter_fwd: invalid arguments");
    constrained_filter_fwd_kernel<<<1, 256, 0, st>>>(k, nf, ksize, channels, strength);
    NI_LAUNCH_CHECK(); NI_COUNT_LAUNCH(1);
    return NI_OK;
}
extern "C" int ni_constrained_filter_bwd(const float* k, const float* dnf, float* dk, int ksize, int channels, float strength,
                                         cudaStream_t st) {
    NI_REQUIRE(k && dnf && dk && ksize > 0 && (ksize & 1) && channels > 0 && channels <= 16, "ni_constrained_filter_bwd: invalid arguments");
    constrained_filter_bwd_kernel<<<1, 256, 0, st>>>(k, dnf, dk, ksize, channels, strength);
    NI_LAUNCH_CHECK(); NI_COUNT_LAUNCH(1);
    return NI_OK;
}

extern "C" int ni_pad_fold(const float* dpad, float* dx, int n, int h, int w, int c, int pad, int mode, int accumulate,
                           cudaStream_t st) {
    NI_REQUIRE(dpad && dx && n >= 0 && h > 0 && w > 0 && c > 0 && pad >= 0, "ni_pad_fold: invalid arguments");
    NI_REQUIRE(mode == NI_PAD_SYMMETRIC || mode == NI_PAD_REFLECT, "ni_pad_fold: mode must be symmetric (1) or reflect (2)");
    NI_REQUIRE(pad <= (mode == NI_PAD_SYMMETRIC ? h : h - 1) && pad <= (mode == NI_PAD_SYMMETRIC ? w : w - 1), "ni_pad_fold: pad too large");
    if (n == 0) return NI_OK;
    pad_fold_kernel<<<grid_for((long long)n * h * w * c), kT, 0, st>>>(dpad, dx, n, h, w, c, pad, mode, accumulate);
    NI_LAUNCH_CHECK(); NI_COUNT_LAUNCH(1);
    return NI_OK;
}

extern "C" int ni_adam_keras(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                             float eps, long long step, float gscale, int* nonfinite_flag, cudaStream_t st) {
    NI_REQUIRE(p && g && m && v && n >= 0 && step >= 1, "ni_adam_keras: invalid arguments");
    if (n == 0) return NI_OK;
    const double lr_t = (double)lr * sqrt(1.0 - pow((double)beta2, (double)step)) / (1.0 - pow((double)beta1, (double)step));
    adam_kernel<<<grid_for(n), kT, 0, st>>>(p, g, m, v, n, (float)lr_t, nullptr, beta1, beta2, eps, gscale, nonfinite_flag);
    NI_LAUNCH_CHECK(); NI_COUNT_LAUNCH(1);
    return NI_OK;
}

// Same update with the bias-corrected step size lr * sqrt(1 - beta2^t) / (1 - beta1^t) supplied in device memory.
extern "C" int ni_adam_keras_dev(float* p, const float* g, float* m, float* v, long long n, const float* lr_t_dev, float beta1,
                                 float beta2, float eps, float gscale, int* nonfinite_flag, cudaStream_t st) {
    NI_REQUIRE(p && g && m && v && lr_t_dev && n >= 0, "ni_adam_keras_dev: invalid arguments");
    if (n == 0) return NI_OK;
    adam_kernel<<<grid_for(n), kT, 0, st>>>(p, g, m, v, n, 0.f, lr_t_dev, beta1, beta2, eps, gscale, nonfinite_flag);
    NI_LAUNCH_CHECK(); NI_COUNT_LAUNCH(1);
    return NI_OK;
}

extern "C" int ni_fill(float* p, float value, long long n, cudaStream_t st) {
    NI_REQUIRE(p && n >= 0, "ni_fill: invalid arguments");
    if (n == 0) return NI_OK;
    fill_kernel<<<grid_for(n), kT, 0, st>>>(p, value, n);
    NI_LAUNCH_CHECK(); NI_COUNT_LAUNCH(1);
    return NI_OK;
}

extern "C" int ni_affine(const float* x, float* y, float a, float b, int clip, long long n, cudaStream_t st) {
    NI_REQUIRE(x && y && n >= 0, "ni_affine: invalid arguments");
    if (n == 0) return NI_OK;
    affine_kernel<<<grid_for(n), kT, 0, st>>>(x, y, a, b, clip, n);
    NI_LAUNCH_CHECK(); NI_COUNT_LAUNCH(1);
    return NI_OK;
}
