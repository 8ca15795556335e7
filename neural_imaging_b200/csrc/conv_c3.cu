// The constrained residual filter of the forensic network (reference models/layers.py:36-57 ConstrainedConv2D.call): SYMMETRIC pad 2,
// VALID 5x5 convolution 3 -> 3 channels, no bias — forward, input gradient (with the transpose of the mirrored pad folded in) and filter
// gradient. 1280 images of 128x128x3 per step: 0.25 GB in, 0.25 GB out, 225 FMAs per pixel. The generic small-channel kernels ran these
// three launches at 9-12 TFLOP/s (0.8 + 1.0 + 0.43 (pad fold) + 1.0 ms per step at B = 256); here every thread owns a 2x2 pixel quad x 3
// channels, streams the shared 6x6 window through registers with 64-bit shared loads, and takes the 225 weights as CONSTANT-BANK operands of
// its FMAs (staged into __constant__ memory by a stream-ordered device-to-device copy): with the weights as 128-bit shared-memory broadcasts the
// kernels issued one LDS per 4.4 FMAs, which is the shared-memory pipe's limit, not the FMA pipe's.
#include "ni_common.cuh"
#include "tile3.cuh"

namespace {

constexpr int kTS = 32;            // tile edge: 16 x 16 quads, one per thread
constexpr int kThreads = 256;

__device__ __forceinline__ int symm_i(int u, int n) {
    if (u < 0) u = -u - 1;
    if (u >= n) u = 2 * n - 1 - u;
    return u < 0 ? 0 : (u >= n ? n - 1 : u);
}

// weights (5, 5, ci, co) -> [tap][ci][4] (co padded to 4) for the forward kernel or, TRANSPOSED, [tap][co][4] (ci padded) for the input
// gradient, in __constant__ memory: every weight index of the unrolled kernels is a compile-time constant, so the FMAs read c[bank][offset]
// directly. The filter lives in device memory (it is renormalised on the device every step), hence prep kernel -> staging buffer ->
// cudaMemcpyToSymbolAsync(device to device) on the caller's stream (stream-ordered, capturable). One filter per direction at a time per
// process: launches that use different filters must be ordered on one stream (they are: the forensic network owns one such layer).
__constant__ float c_wf[25 * 3 * 4];
__constant__ float c_wb[25 * 3 * 4];
__device__ float g_wstage[2][25 * 3 * 4];

__global__ void cconv5_prep_weights_kernel(const float* __restrict__ w, float* __restrict__ dst, int transposed) {
    const int t = threadIdx.x;
    if (t >= 25 * 3 * 4) return;
    const int e = t & 3, m = (t >> 2) % 3, tap = t / 12;
    float v = 0.f;
    if (e < 3) v = transposed ? __ldg(w + (tap * 3 + e) * 3 + m) : __ldg(w + (tap * 3 + m) * 3 + e);
    dst[t] = v;
}

int stage_weights(const float* w, int transposed, cudaStream_t st) {
    static float* stage = nullptr;
    if (!stage) NI_CUDA(cudaGetSymbolAddress(reinterpret_cast<void**>(&stage), g_wstage));
    float* dst = stage + transposed * (25 * 3 * 4);
    cconv5_prep_weights_kernel<<<1, 320, 0, st>>>(w, dst, transposed);
    NI_LAUNCH_CHECK(); NI_COUNT_LAUNCH(1);
    if (transposed) NI_CUDA(cudaMemcpyToSymbolAsync(c_wb, dst, sizeof(float) * 25 * 3 * 4, 0, cudaMemcpyDeviceToDevice, st));
    else NI_CUDA(cudaMemcpyToSymbolAsync(c_wf, dst, sizeof(float) * 25 * 3 * 4, 0, cudaMemcpyDeviceToDevice, st));
    return NI_OK;
}

// ------------------------------------------------------------------------------------------------------------ forward
constexpr int kFH = 2, kFX = 4, kFC = 40, kFRS = kFC * 3;

// Thread = 2 rows x 4 columns x 3 channels (24 accumulators): the 128 uniform-register loads of the 225 weights and the per-thread overheads
// are paid once per 8 pixels (a 2 x 2 quad per thread was issue-bound with 58 % of its issued instructions being FMAs: ncu, issue slots 77 % busy):
// 0.302 -> 0.272 ms. The input gradient keeps 2 x 2 quads: with 2 x 4 it needs 96 registers and lost more in occupancy (0.519 -> 0.603 ms).
constexpr int kQX = 4;                              // pixel columns per thread
constexpr int kThreadsQ = (kTS / 2) * (kTS / kQX);  // 128 threads per 32 x 32 tile

__global__ void __launch_bounds__(kThreadsQ, 6)
cconv5_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int H, int W) {
    __shared__ __align__(16) float tile[(kTS + 2 * kFH) * kFRS];
    const int n = blockIdx.z, y0 = blockIdx.y * kTS, x0 = blockIdx.x * kTS;
    load_tile3<kTS, kFH, kFX, kFC, TILE_SYMMETRIC, kThreadsQ>(tile, x + (size_t)n * H * W * 3, H, W, y0, x0);
    __syncthreads();
    const int ty = threadIdx.x / (kTS / kQX), tx = threadIdx.x % (kTS / kQX);
    const int py = y0 + 2 * ty, px = x0 + kQX * tx;
    if (py >= H || px >= W) return;
    float acc[2][kQX][3];
#pragma unroll
    for (int qa = 0; qa < 2; ++qa)
#pragma unroll
        for (int qb = 0; qb < kQX; ++qb) acc[qa][qb][0] = acc[qa][qb][1] = acc[qa][qb][2] = 0.f;
#pragma unroll
    for (int r = 0; r < 6; ++r) {
        float row[(kQX + 4) * 3];     // window row r: padded-image row py - 2 + r, columns px - 2 .. px + kQX + 1
        const float2* p = reinterpret_cast<const float2*>(tile + (2 * ty + r) * kFRS + (kQX * tx + kFX - 2) * 3);
#pragma unroll
        for (int v = 0; v < (kQX + 4) * 3 / 2; ++v) { const float2 t2 = p[v]; row[2 * v] = t2.x; row[2 * v + 1] = t2.y; }
#pragma unroll
        for (int qa = 0; qa < 2; ++qa) {
            const int a = r - qa;
            if (a < 0 || a > 4) continue;
#pragma unroll
            for (int b = 0; b < 5; ++b)
#pragma unroll
                for (int ci = 0; ci < 3; ++ci) {
                    const int wi = ((a * 5 + b) * 3 + ci) * 4;          // compile-time after unrolling: constant-bank operands
#pragma unroll
                    for (int qb = 0; qb < kQX; ++qb) {
                        const float xv = row[(qb + b) * 3 + ci];
                        acc[qa][qb][0] = fmaf(c_wf[wi], xv, acc[qa][qb][0]);
                        acc[qa][qb][1] = fmaf(c_wf[wi + 1], xv, acc[qa][qb][1]);
                        acc[qa][qb][2] = fmaf(c_wf[wi + 2], xv, acc[qa][qb][2]);
                    }
                }
        }
    }
#pragma unroll
    for (int qa = 0; qa < 2; ++qa) {
        if (py + qa >= H) continue;
        float* o = y + (((size_t)n * H + py + qa) * W + px) * 3;
        if (px + kQX <= W && (W & 3) == 0) {
            float4* o4 = reinterpret_cast<float4*>(o);       // 12 consecutive floats, 16-byte aligned (px and W multiples of 4)
            o4[0] = make_float4(acc[qa][0][0], acc[qa][0][1], acc[qa][0][2], acc[qa][1][0]);
            o4[1] = make_float4(acc[qa][1][1], acc[qa][1][2], acc[qa][2][0], acc[qa][2][1]);
            o4[2] = make_float4(acc[qa][2][2], acc[qa][3][0], acc[qa][3][1], acc[qa][3][2]);
        } else {
#pragma unroll
            for (int qb = 0; qb < kQX; ++qb)
                if (px + qb < W) { o[qb * 3] = acc[qa][qb][0]; o[qb * 3 + 1] = acc[qa][qb][1]; o[qb * 3 + 2] = acc[qa][qb][2]; }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------ input gradient
// dxp[u, ci] = sum_{a,b,co} w[a,b,ci,co] dy[u + 2 - (a,b), co] on the padded domain u in [-2, H+1] x [-2, W+1] (dy = 0 outside the image),
// dx[p] = sum of dxp over every u that SYMMETRIC padding maps onto p (the transpose of tf.pad): u = p, -p - 1 (p <= 1), 2n - 1 - p (p >= n - 2).
constexpr int kBH = 4, kBX = 8, kBC = 48, kBRS = kBC * 3;
constexpr int kRN = kTS + 4;                     // padded-domain cells of a tile incl. the 2-cell ring

__device__ __forceinline__ int symm_aliases(int p, int n, int (&u)[3]) {
    int c = 0;
    u[c++] = p;
    if (p <= 1) u[c++] = -p - 1;
    if (p >= n - 2) u[c++] = 2 * n - 1 - p;
    return c;
}

__global__ void __launch_bounds__(kThreads, 3)
cconv5_bwd_data_kernel(const float* __restrict__ dy, float* __restrict__ dx, int H, int W, int accumulate) {
    __shared__ __align__(16) float tile[(kTS + 2 * kBH) * kBRS];      // dy, rows y0 - 4 .., columns x0 - 8 .., zero outside the image
    __shared__ float ring[kRN * kRN * 3];                             // dxp of the cells OUTSIDE the image (border tiles only)
    const int n = blockIdx.z, y0 = blockIdx.y * kTS, x0 = blockIdx.x * kTS;
    load_tile3<kTS, kBH, kBX, kBC, TILE_ZERO, kThreads>(tile, dy + (size_t)n * H * W * 3, H, W, y0, x0);
    __syncthreads();
    const bool border = (y0 < 2) || (x0 < 2) || (y0 + kTS + 2 > H) || (x0 + kTS + 2 > W);
    if (border) {
        for (int t = threadIdx.x; t < kRN * kRN; t += kThreads) {
            const int i = t / kRN, j = t - i * kRN;
            const int uy = y0 - 2 + i, ux = x0 - 2 + j;
            if (uy >= 0 && uy < H && ux >= 0 && ux < W) continue;          // inside the image: owned by a tile's interior pass
            if (uy < -2 || uy > H + 1 || ux < -2 || ux > W + 1) continue;
            float a0 = 0.f, a1 = 0.f, a2 = 0.f;
            for (int a = 0; a < 5; ++a)
                for (int b = 0; b < 5; ++b) {
                    // dy[u + 2 - (a, b)] in tile coordinates (row 0 = y0 - 4, column 0 = x0 - 8)
                    const float* g = tile + (i + 4 - a) * kBRS + (j + 8 - b) * 3;
#pragma unroll
                    for (int co = 0; co < 3; ++co) {
                        const float* wv = c_wb + ((a * 5 + b) * 3 + co) * 4;      // run-time index: constant-cache loads (border tiles only)
                        a0 = fmaf(wv[0], g[co], a0); a1 = fmaf(wv[1], g[co], a1); a2 = fmaf(wv[2], g[co], a2);
                    }
                }
            ring[t * 3] = a0; ring[t * 3 + 1] = a1; ring[t * 3 + 2] = a2;
        }
    }
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    const int py = y0 + 2 * ty, px = x0 + 2 * tx;
    float acc[2][2][3];
#pragma unroll
    for (int qa = 0; qa < 2; ++qa)
#pragma unroll
        for (int qb = 0; qb < 2; ++qb) acc[qa][qb][0] = acc[qa][qb][1] = acc[qa][qb][2] = 0.f;
#pragma unroll
    for (int r = 0; r < 6; ++r) {
        float row[18];     // window row r: dy row py - 2 + r, columns px - 2 .. px + 3
        const float2* p = reinterpret_cast<const float2*>(tile + (2 * ty + kBH - 2 + r) * kBRS + (2 * tx + kBX - 2) * 3);
#pragma unroll
        for (int v = 0; v < 9; ++v) { const float2 t2 = p[v]; row[2 * v] = t2.x; row[2 * v + 1] = t2.y; }
#pragma unroll
        for (int qa = 0; qa < 2; ++qa) {
            const int a = 4 + qa - r;          // dy row (py + qa) + 2 - a == py - 2 + r
            if (a < 0 || a > 4) continue;
#pragma unroll
            for (int b = 0; b < 5; ++b)
#pragma unroll
                for (int co = 0; co < 3; ++co) {
                    const int wi = ((a * 5 + b) * 3 + co) * 4;          // compile-time after unrolling: constant-bank operands
#pragma unroll
                    for (int qb = 0; qb < 2; ++qb) {
                        const float g = row[(4 + qb - b) * 3 + co];
                        acc[qa][qb][0] = fmaf(c_wb[wi], g, acc[qa][qb][0]);
                        acc[qa][qb][1] = fmaf(c_wb[wi + 1], g, acc[qa][qb][1]);
                        acc[qa][qb][2] = fmaf(c_wb[wi + 2], g, acc[qa][qb][2]);
                    }
                }
        }
    }
    if (border) __syncthreads();
    if (py >= H || px >= W) return;
#pragma unroll
    for (int qa = 0; qa < 2; ++qa) {
        if (py + qa >= H) continue;
#pragma unroll
        for (int qb = 0; qb < 2; ++qb) {
            if (!border || px + qb >= W) continue;
            int uy[3], ux[3];
            const int ny = symm_aliases(py + qa, H, uy), nx = symm_aliases(px + qb, W, ux);
            for (int iy = 0; iy < ny; ++iy)
                for (int ix = 0; ix < nx; ++ix) {
                    if (iy == 0 && ix == 0) continue;           // u = p itself is the interior result in registers
                    const float* g = ring + ((uy[iy] - (y0 - 2)) * kRN + (ux[ix] - (x0 - 2))) * 3;
                    acc[qa][qb][0] += g[0]; acc[qa][qb][1] += g[1]; acc[qa][qb][2] += g[2];
                }
        }
        float* o = dx + (((size_t)n * H + py + qa) * W + px) * 3;
        float v[6] = {acc[qa][0][0], acc[qa][0][1], acc[qa][0][2], acc[qa][1][0], acc[qa][1][1], acc[qa][1][2]};
        if (px + 1 < W) {
            float2* o2 = reinterpret_cast<float2*>(o);
            if (accumulate) {
                const float2 v0 = o2[0], v1 = o2[1], v2 = o2[2];
                v[0] += v0.x; v[1] += v0.y; v[2] += v1.x; v[3] += v1.y; v[4] += v2.x; v[5] += v2.y;
            }
            o2[0] = make_float2(v[0], v[1]); o2[1] = make_float2(v[2], v[3]); o2[2] = make_float2(v[4], v[5]);
        } else {
            if (accumulate) { v[0] += o[0]; v[1] += o[1]; v[2] += o[2]; }
            o[0] = v[0]; o[1] = v[1]; o[2] = v[2];
        }
    }
}

// ------------------------------------------------------------------------------------------------------------ filter gradient
// dw[a, b, ci, co] = sum_{n, p} xpad[p + (a, b) - 2, ci] dy[p, co]. Thread role = (a, ci) with 15 accumulators (3 co x 5 b); 16 row-splits per
// role (two tile rows each); planar tiles so that a 4-pixel step is two 128-bit loads of x and three of dy feeding 60 FMAs (the first
// version's role = (a, ci, co) fed 20 FMAs from three loads and sat on the shared-memory pipe). Row pitch 36 and a plane pitch = 3 mod 8
// 16-byte slots put the eight (a, ci) rows of a quarter-warp in eight different bank slots. Persistent over tiles; one atomicAdd per output
// per CTA.
constexpr int kWX = 36;                          // planar x tile: 36 rows x 36 columns (x0 - 2 ..)
constexpr int kWP = (kTS + 4) * kWX + 28;        // x plane pitch: 331 slots of 16 bytes = 3 (mod 8)
constexpr int kDX = 36;                          // planar dy tile: 32 rows x 32 columns, row pitch 36
constexpr int kRoles = 15, kSplits = 16;

__global__ void __launch_bounds__(kThreads)
cconv5_bwd_filter_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dw, int N, int H, int W,
                         int tiles_x, int tiles_y) {
    __shared__ __align__(16) float sx[3 * kWP];
    __shared__ __align__(16) float sd[3 * kTS * kDX];
    const int role = threadIdx.x % kRoles, split = threadIdx.x / kRoles;      // threads 240 .. 255 only help loading
    const bool active = threadIdx.x < kRoles * kSplits;
    const int a = role / 3, ci = role % 3;
    float acc[3][5];
#pragma unroll
    for (int co = 0; co < 3; ++co)
#pragma unroll
        for (int bb = 0; bb < 5; ++bb) acc[co][bb] = 0.f;
    const int total = tiles_x * tiles_y * N;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int n = tile / (tiles_x * tiles_y), tr = tile - n * tiles_x * tiles_y;
        const int y0 = (tr / tiles_x) * kTS, x0 = (tr % tiles_x) * kTS;
        const float* xi = x + (size_t)n * H * W * 3;
        const float* di = dy + (size_t)n * H * W * 3;
        __syncthreads();                                   // previous tile fully consumed
        // interleaved rows -> planar tiles: one 128-bit load per 4 floats where the slot lies inside the image row
        {
            // x: columns x0 - 2 .. x0 + 33 start at an even, not 16-byte aligned float: element-wise through the index map
            for (int t = threadIdx.x; t < (kTS + 4) * 36 * 3; t += kThreads) {
                const int ch = t % 3, col = (t / 3) % 36, r = t / 108;
                sx[ch * kWP + r * kWX + col] = __ldg(xi + ((size_t)symm_i(y0 - 2 + r, H) * W + symm_i(x0 - 2 + col, W)) * 3 + ch);
            }
            const bool vec = (W & 3) == 0 && x0 + kTS <= W;
            if (vec) {
                constexpr int DV = kTS * 3 / 4;                          // 24 slots per row, 16-byte aligned (x0 multiple of 32)
                for (int t = threadIdx.x; t < kTS * DV; t += kThreads) {
                    const int r = t / DV, v = t - r * DV;
                    const int gy = y0 + r;
                    float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (gy < H) val = ni_ldg4(di + ((size_t)gy * W + x0) * 3 + 4 * v);
                    const float e[4] = {val.x, val.y, val.z, val.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int f = 4 * v + k, col = f / 3, ch = f - col * 3;
                        sd[(ch * kTS + r) * kDX + col] = e[k];
                    }
                }
            } else {
                for (int t = threadIdx.x; t < kTS * kTS * 3; t += kThreads) {
                    const int ch = t % 3, col = (t / 3) % kTS, r = t / (3 * kTS);
                    const int gy = y0 + r, gx = x0 + col;
                    sd[(ch * kTS + r) * kDX + col] = (gy < H && gx < W) ? __ldg(di + ((size_t)gy * W + gx) * 3 + ch) : 0.f;
                }
            }
        }
        __syncthreads();
        if (!active) continue;
#pragma unroll
        for (int rr = 0; rr < kTS / kSplits; ++rr) {
            const int yy = split * (kTS / kSplits) + rr;
            const float* xr = sx + ci * kWP + (yy + a) * kWX;
            const float* dr = sd + yy * kDX;
#pragma unroll
            for (int xq = 0; xq < kTS; xq += 4) {
                const float4 x0v = *reinterpret_cast<const float4*>(xr + xq), x1v = *reinterpret_cast<const float4*>(xr + xq + 4);
                const float xv[8] = {x0v.x, x0v.y, x0v.z, x0v.w, x1v.x, x1v.y, x1v.z, x1v.w};
#pragma unroll
                for (int co = 0; co < 3; ++co) {
                    const float4 dv = *reinterpret_cast<const float4*>(dr + co * kTS * kDX + xq);
                    const float g[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e)
#pragma unroll
                        for (int bb = 0; bb < 5; ++bb) acc[co][bb] = fmaf(xv[e + bb], g[e], acc[co][bb]);
                }
            }
        }
    }
    // reduce the row-splits through shared memory, then one atomicAdd per output per CTA
    __syncthreads();
    float* red = sx;                                       // (kSplits - 1) * 15 roles * 15 accumulators = 3375 floats <= 3 * kWP
    if (active && split > 0) {
#pragma unroll
        for (int co = 0; co < 3; ++co)
#pragma unroll
            for (int bb = 0; bb < 5; ++bb) red[((split - 1) * kRoles + role) * 15 + co * 5 + bb] = acc[co][bb];
    }
    __syncthreads();
    if (active && split == 0) {
#pragma unroll
        for (int co = 0; co < 3; ++co)
#pragma unroll
            for (int bb = 0; bb < 5; ++bb) {
                float v = acc[co][bb];
                for (int sp = 0; sp < kSplits - 1; ++sp) v += red[(sp * kRoles + role) * 15 + co * 5 + bb];
                atomicAdd(dw + ((a * 5 + bb) * 3 + ci) * 3 + co, v);
            }
    }
}

}  // namespace

extern "C" int ni_cconv5_fwd(const float* x, const float* w, float* y, int n, int h, int w_, cudaStream_t st) {
    NI_REQUIRE(x && w && y && n >= 0 && h >= 4 && w_ >= 4, "ni_cconv5_fwd: invalid arguments");
    if (n == 0) return NI_OK;
    int rc = stage_weights(w, 0, st);
    if (rc) return rc;
    dim3 grid(ni_cdiv(w_, kTS), ni_cdiv(h, kTS), n);
    cconv5_fwd_kernel<<<grid, kThreadsQ, 0, st>>>(x, y, h, w_);
    NI_LAUNCH_CHECK(); NI_COUNT_LAUNCH(1);
    return NI_OK;
}

extern "C" int ni_cconv5_bwd_data(const float* dy, const float* w, float* dx, int n, int h, int w_, int accumulate, cudaStream_t st) {
    NI_REQUIRE(dy && w && dx && n >= 0 && h >= 4 && w_ >= 4, "ni_cconv5_bwd_data: invalid arguments");
    if (n == 0) return NI_OK;
    int rc = stage_weights(w, 1, st);
    if (rc) return rc;
    dim3 grid(ni_cdiv(w_, kTS), ni_cdiv(h, kTS), n);
    cconv5_bwd_data_kernel<<<grid, kThreads, 0, st>>>(dy, dx, h, w_, accumulate);
    NI_LAUNCH_CHECK(); NI_COUNT_LAUNCH(1);
    return NI_OK;
}

extern "C" int ni_cconv5_bwd_filter(const float* x, const float* dy, float* dw, int n, int h, int w_, cudaStream_t st) {
    NI_REQUIRE(x && dy && dw && n >= 0 && h >= 4 && w_ >= 4, "ni_cconv5_bwd_filter: invalid arguments");
    NI_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * 225, st));
    if (n == 0) return NI_OK;
    const int tx = ni_cdiv(w_, kTS), ty = ni_cdiv(h, kTS);
    const long long total = (long long)tx * ty * n;
    const int grid = (int)(total < 4LL * ni_num_sms() ? total : 4LL * ni_num_sms());
    cconv5_bwd_filter_kernel<<<grid, kThreads, 0, st>>>(x, dy, dw, n, h, w_, tx, ty);
    NI_LAUNCH_CHECK(); NI_COUNT_LAUNCH(1);
    return NI_OK;
}
