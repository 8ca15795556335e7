// Register-blocked direct FP32 convolutions for the hot non-dense layers of the joint step (FAN front end 5x5 3->32 and its
// input gradient 32->3, the U-Net input 3x3 4->32 and output 3x3 32->12 + depth_to_space; reference models/forensics.py:62-68,
// models/pipelines.py:190,215-218). They replace the thread-per-pixel stencils of conv_small.cu for these shapes: those issued
// one shared-memory load per 3-4 FMAs (measured 12-26 TFLOP/s); here every thread owns 4 pixels x 3..16 channels so that a
// shared-memory operand feeds 10-16 FMAs.
//
//   fewin  (CIN <= 4)   : thread = 4 pixels of a row x COUT/2 output channels; planar input tile, weights broadcast.
//   manyin (CIN % 4 = 0): thread = 4 pixels of a COLUMN x all COUT (3..12) channels; input tile stored channel-quad-major so
//                         that neighbouring lanes read neighbouring 16-byte words (conflict-free LDS.128). Also serves as
//                         dgrad of the few-input layers (flipped, channel-swapped filter).
//   wgrad               : thread = (filter row, ci, group of COG output channels) with K x COG accumulators, sweeping the
//                         pixels of 8x32 tiles four at a time; register accumulation across tiles, one atomicAdd per output.
#include "conv_desc.h"
#include "ni_common.cuh"
#include "views.cuh"

namespace {

struct DirectParams {
    TensorView src, dst;
    int n, pad_t, pad_l, pad_mode;
    int act, bias_mod; float alpha; int accumulate;
    // POOL variant of the few-input kernel (conv + bias + activation + 2x2 max-pool in one pass): pooled output (n, H/2, W/2, COUT) and one
    // code byte per pooled element = argmax position (bits 0-1, scan order) | (value > 0) << 2 — what the backward needs instead of the
    // full-resolution activation. dst is then only a shape.
    float* pool; unsigned char* code;
};

__device__ __forceinline__ int mirror_idx(int u, int n, int mode) {
    if (mode == NI_PAD_SYMMETRIC) { if (u < 0) u = -u - 1; if (u >= n) u = 2 * n - 1 - u; }
    else { if (u < 0) u = -u; if (u >= n) u = 2 * (n - 1) - u; }
    return u;
}

__device__ __forceinline__ float act_direct(float v, int act, float alpha) {
    switch (act) {
        case NI_ACT_LEAKY_RELU: return v > 0.f ? v : alpha * v;
        case NI_ACT_RELU: return fmaxf(v, 0.f);
        case NI_ACT_TANH: return tanhf(v);
        case NI_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
        case NI_ACT_CLIP01: return ni_clamp01(v);
        default: return v;
    }
}

// Asynchronous global -> shared copies with zero fill (src-size 0) for out-of-range taps: a whole tile is in flight at once,
// whereas a load -> store loop serialises one memory round trip per iteration (measured: 31 us of loads for 5 us of math per tile).
__device__ __forceinline__ void cp_async4_zfill(float* smem, const float* gmem, bool valid) {
    const unsigned sa = static_cast<unsigned>(__cvta_generic_to_shared(smem));
    const int sz = valid ? 4 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(sa), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async16_zfill(float* smem, const float* gmem, bool valid) {
    const unsigned sa = static_cast<unsigned>(__cvta_generic_to_shared(smem));
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(sa), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

__device__ __forceinline__ bool view_vec4(const TensorView& v) { return v.mode == NI_MODE_PLAIN && !(v.pitch & 3) && !(v.coff & 3); }

// ------------------------------------------------------------------------------------------------ few input channels
// Note on the operand path: shared-memory operands cost register write-back bandwidth (128 B/clk/SM = 1 byte per FMA per thread
// at full FMA rate) whether they are broadcasts or not; these kernels move 0.5 - 1.5 B per FMA and sit at 30 - 34 TFLOP/s.
// Measured alternative: weights in __constant__ memory indexed by the (rolled) filter-row loop compile to per-thread LDC
// loads, which were no faster (3.06 vs 2.99 ms for the FAN 3->32 fprop) and add a global staging buffer -- not kept.
constexpr int FW = 64, FH = 8;     // output tile, 256 threads = 64 pixel groups (2 rows x 4 columns) x 4 channel quarters
constexpr int FXS = FW + 4;        // input tile row stride (floats): room for the 8-wide window of the last pixel group

// Persistent: a CTA loads the filter once and walks over tiles with the NEXT tile's input in flight (cp.async into the other buffer)
// while it computes the current one. The one-tile-per-CTA version spent ~1/3 of its life in the load phase (filter + tile through 4-byte
// cp.async, then wait_all: issue slots 63 % busy, long-scoreboard the top stall, ncu r2_prof_fan_first). Thread = 2 x 4 pixels x COUT/4
// channels: 14 LDS.128 per 320 FMAs (was 22 with 1 x 4 pixels x COUT/2), and a 2x2 pooling window lives inside one thread.
template <int CIN, int COUT, int K, bool POOL>
__global__ void __launch_bounds__(256, 2)
conv_fewin_kernel(DirectParams p, const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ y) {
    static_assert(K <= 5 && COUT % 16 == 0, "window of 4 + K - 1 <= 8 input columns; four channel quarters of whole float4s");
    constexpr int COG = COUT / 4, IH = FH + K - 1, XT = CIN * IH * FXS;
    extern __shared__ __align__(16) float smem[];
    float* sw = smem;                              // [K*K][CIN][COUT]
    float* sx0 = smem + K * K * CIN * COUT;        // 2 x [CIN][IH][FXS]
    const int tid = threadIdx.x;
    const int tiles_x = (p.dst.W + FW - 1) / FW, tiles_y = (p.dst.H + FH - 1) / FH;
    const int tiles_total = tiles_x * tiles_y * p.n;

    // fast path of the tile copy (plain NHWC input with exactly CIN channels, zero padding): a tile row is one contiguous run of
    // FXS * CIN floats; indices advance incrementally (the generic path below spends ~80 instructions per element on divisions and
    // 64-bit view arithmetic, 17 % of the kernel's issue slots)
    const bool fast_in = p.src.mode == NI_MODE_PLAIN && p.src.pitch == CIN && p.src.coff == 0 && p.pad_mode == NI_PAD_ZERO;
    auto load_tile = [&](int tile, float* sx) {
        const int tx0 = (tile % tiles_x) * FW, ty0 = ((tile / tiles_x) % tiles_y) * FH, n = tile / (tiles_x * tiles_y);
        if (fast_in) {
            constexpr int RL = FXS * CIN;
            const float* img = x + (long long)n * p.src.H * p.src.W * CIN;
            int pxc = tid % RL, py = tid / RL;
#pragma unroll 1
            while (py < IH) {
                const int px = pxc / CIN, c = pxc - px * CIN;
                const int sy = ty0 + py - p.pad_t, sxx = tx0 + px - p.pad_l;
                const bool in = (unsigned)sy < (unsigned)p.src.H && (unsigned)sxx < (unsigned)p.src.W;
                cp_async4_zfill(sx + (c * IH + py) * FXS + px, in ? img + (sy * p.src.W + tx0 - p.pad_l) * CIN + pxc : x, in);
                pxc += 256 % RL; py += 256 / RL;
                if (pxc >= RL) { pxc -= RL; ++py; }
            }
            return;
        }
        for (int i = tid; i < XT; i += 256) {
            const int c = i % CIN, px = (i / CIN) % FXS, py = i / (CIN * FXS);
            int sy = ty0 + py - p.pad_t, sxx = tx0 + px - p.pad_l;
            if (p.pad_mode != NI_PAD_ZERO) { sy = mirror_idx(sy, p.src.H, p.pad_mode); sxx = mirror_idx(sxx, p.src.W, p.pad_mode); }
            const bool in = sy >= 0 && sy < p.src.H && sxx >= 0 && sxx < p.src.W;
            cp_async4_zfill(sx + (c * IH + py) * FXS + px, in ? x + view_addr(p.src, n, sy, sxx, c) : x, in);
        }
    };

    for (int i = tid; i < K * K * CIN * COUT; i += 256) cp_async4_zfill(sw + i, w + i, true);   // w: slice of the flat parameter buffer (4-byte aligned)
    if ((int)blockIdx.x < tiles_total) load_tile(blockIdx.x, sx0);

    const int pg = tid & 63, cog = tid >> 6;       // warp-uniform channel quarter: weight loads are pure broadcasts
    const int gx = pg & 15, gy = (pg >> 4) * 2;    // first of the thread's two rows
    float bv[COG];
#pragma unroll
    for (int j = 0; j < COG; ++j) {
        const int co = cog * COG + j;
        bv[j] = bias ? __ldg(bias + (p.bias_mod > 0 ? co % p.bias_mod : co)) : 0.f;
    }
    const bool vec = view_vec4(p.dst);

    int buf = 0;
    for (int tile = blockIdx.x; tile < tiles_total; tile += gridDim.x, buf ^= 1) {
        cp_async_wait();
        __syncthreads();        // this tile (and the filter) landed; every thread is done with the buffer the next copy overwrites
        if (tile + (int)gridDim.x < tiles_total) load_tile(tile + gridDim.x, sx0 + (buf ^ 1) * XT);
        const float* sx = sx0 + buf * XT;
        const int tx0 = (tile % tiles_x) * FW, ty0 = ((tile / tiles_x) % tiles_y) * FH, n = tile / (tiles_x * tiles_y);

        float acc[2][4][COG];
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int j = 0; j < COG; ++j) acc[r][q][j] = 0.f;
#pragma unroll 1
        for (int a = 0; a < K; ++a) {
#pragma unroll
            for (int ci = 0; ci < CIN; ++ci) {
                float xr[2][8];
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const float4* xp = reinterpret_cast<const float4*>(sx + (ci * IH + gy + r + a) * FXS + gx * 4);
                    const float4 x0 = xp[0], x1 = xp[1];
                    xr[r][0] = x0.x; xr[r][1] = x0.y; xr[r][2] = x0.z; xr[r][3] = x0.w;
                    xr[r][4] = x1.x; xr[r][5] = x1.y; xr[r][6] = x1.z; xr[r][7] = x1.w;
                }
#pragma unroll
                for (int b = 0; b < K; ++b) {
                    const float4* wp = reinterpret_cast<const float4*>(sw + ((a * K + b) * CIN + ci) * COUT + cog * COG);
#pragma unroll
                    for (int j4 = 0; j4 < COG / 4; ++j4) {
                        const float4 wv = wp[j4];
#pragma unroll
                        for (int r = 0; r < 2; ++r)
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                acc[r][q][4 * j4] = fmaf(xr[r][q + b], wv.x, acc[r][q][4 * j4]);
                                acc[r][q][4 * j4 + 1] = fmaf(xr[r][q + b], wv.y, acc[r][q][4 * j4 + 1]);
                                acc[r][q][4 * j4 + 2] = fmaf(xr[r][q + b], wv.z, acc[r][q][4 * j4 + 2]);
                                acc[r][q][4 * j4 + 3] = fmaf(xr[r][q + b], wv.w, acc[r][q][4 * j4 + 3]);
                            }
                    }
                }
            }
        }
        const int oy = ty0 + gy, ox0 = tx0 + gx * 4;
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int j = 0; j < COG; ++j) acc[r][q][j] = act_direct(acc[r][q][j] + bv[j], p.act, p.alpha);
        if constexpr (POOL) {
            // 2x2 max-pool inside the thread: pooled pixels (columns 0,1 | 2,3) of its row pair; first maximum in scan order
            // (row 0 col 0, (0,1), (1,0), (1,1)) wins, as in ni_maxpool2_fwd / tf.nn.max_pool's gradient routing
            if (oy + 1 >= p.dst.H) continue;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (ox0 + 2 * h + 1 >= p.dst.W) continue;
                float pv[COG];
                unsigned int cd[COG / 4];
#pragma unroll
                for (int j4 = 0; j4 < COG / 4; ++j4) cd[j4] = 0u;
#pragma unroll
                for (int j = 0; j < COG; ++j) {
                    float m = acc[0][2 * h][j];
                    unsigned int arg = 0u;
                    if (acc[0][2 * h + 1][j] > m) { m = acc[0][2 * h + 1][j]; arg = 1u; }
                    if (acc[1][2 * h][j] > m) { m = acc[1][2 * h][j]; arg = 2u; }
                    if (acc[1][2 * h + 1][j] > m) { m = acc[1][2 * h + 1][j]; arg = 3u; }
                    pv[j] = m;
                    cd[j >> 2] |= (arg | (m > 0.f ? 4u : 0u)) << (8 * (j & 3));
                }
                const long long o = (((long long)n * (p.dst.H >> 1) + (oy >> 1)) * (p.dst.W >> 1) + (ox0 >> 1) + h) * COUT + cog * COG;
                float4* po = reinterpret_cast<float4*>(p.pool + o);
#pragma unroll
                for (int j4 = 0; j4 < COG / 4; ++j4) po[j4] = make_float4(pv[4 * j4], pv[4 * j4 + 1], pv[4 * j4 + 2], pv[4 * j4 + 3]);
                unsigned int* co4 = reinterpret_cast<unsigned int*>(p.code + o);
#pragma unroll
                for (int j4 = 0; j4 < COG / 4; ++j4) co4[j4] = cd[j4];
            }
        } else {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            if (oy + r >= p.dst.H) continue;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int ox = ox0 + q;
                if (ox >= p.dst.W) continue;
                if (vec) {
                    float4* o = reinterpret_cast<float4*>(y + view_addr(p.dst, n, oy + r, ox, cog * COG));
#pragma unroll
                    for (int j4 = 0; j4 < COG / 4; ++j4) {
                        float4 v = make_float4(acc[r][q][4 * j4], acc[r][q][4 * j4 + 1], acc[r][q][4 * j4 + 2], acc[r][q][4 * j4 + 3]);
                        if (p.accumulate) { const float4 old = o[j4]; v.x += old.x; v.y += old.y; v.z += old.z; v.w += old.w; }
                        o[j4] = v;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < COG; ++j) {
                        float* o = y + view_addr(p.dst, n, oy + r, ox, cog * COG + j);
                        *o = p.accumulate ? *o + acc[r][q][j] : acc[r][q][j];
                    }
                }
            }
        }
        }
    }
}

// ------------------------------------------------------------------------------------------------ many input channels
constexpr int MW = 32, MH = 16;    // output tile, 128 threads: lane = column, warp = group of 4 rows

// weights re-ordered for the column-blocked kernel: [b][c4][a][ci % 4][co]  (one contiguous run per (b, c4))
__global__ void reorder_manyin_weights_kernel(const float* __restrict__ w, float* __restrict__ wr, int k, int cin, int cout, int flip) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= k * k * cin * cout) return;
    const int co = i % cout, e = (i / cout) % 4, a = (i / (cout * 4)) % k, c4 = (i / (cout * 4 * k)) % (cin / 4), b = i / (cout * 4 * k * (cin / 4));
    const int ci = c4 * 4 + e;
    // flip: dgrad = forward conv of dy with the spatially flipped filter and swapped channel axes (w is (k,k,cout_fwd=cin... see caller)
    wr[i] = flip ? w[(((k - 1 - a) * k + (k - 1 - b)) * cout + co) * cin + ci] : w[((a * k + b) * cin + ci) * cout + co];
}

template <int CIN, int COUT, int K>
__global__ void __launch_bounds__(128, 2)
conv_manyin_kernel(DirectParams p, const float* __restrict__ x, const float* __restrict__ wr, const float* __restrict__ bias, float* __restrict__ y) {
    static_assert(CIN % 4 == 0, "channel quads");
    constexpr int C4 = CIN / 4, IH = MH + K - 1, IW = MW + K - 1, R = 4 + K - 1;
    constexpr int PLANE = IH * IW * 4 + 4;         // +16 bytes: the 8 quads of a pixel land in 8 different bank groups
    extern __shared__ __align__(16) float smem[];
    float* sw = smem;                              // [K][C4][K][4][COUT]
    float* sx = smem + K * K * CIN * COUT;         // [C4][IH][IW][4]
    const int tid = threadIdx.x;
    const int tiles_x = (p.dst.W + MW - 1) / MW, tiles_y = (p.dst.H + MH - 1) / MH;
    const int tx0 = (blockIdx.x % tiles_x) * MW, ty0 = ((blockIdx.x / tiles_x) % tiles_y) * MH, n = blockIdx.x / (tiles_x * tiles_y);

    for (int i = tid; i < K * K * CIN * COUT / 4; i += 128) cp_async16_zfill(sw + 4 * i, wr + 4 * i, true);
    const bool vec_in = view_vec4(p.src);
    for (int i = tid; i < IH * IW * C4; i += 128) {
        const int c4 = i % C4, pxl = i / C4, px = pxl % IW, py = pxl / IW;
        int sy = ty0 + py - p.pad_t, sxx = tx0 + px - p.pad_l;
        if (p.pad_mode != NI_PAD_ZERO) { sy = mirror_idx(sy, p.src.H, p.pad_mode); sxx = mirror_idx(sxx, p.src.W, p.pad_mode); }
        const bool in = sy >= 0 && sy < p.src.H && sxx >= 0 && sxx < p.src.W;
        float* dst = sx + c4 * PLANE + pxl * 4;
        if (vec_in) cp_async16_zfill(dst, in ? x + view_addr(p.src, n, sy, sxx, c4 * 4) : x, in);
        else {
#pragma unroll
            for (int e = 0; e < 4; ++e) cp_async4_zfill(dst + e, in ? x + view_addr(p.src, n, sy, sxx, c4 * 4 + e) : x, in);
        }
    }
    cp_async_wait();
    __syncthreads();

    const int lane = tid & 31, rg = tid >> 5;
    float acc[4][COUT];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int j = 0; j < COUT; ++j) acc[q][j] = 0.f;
#pragma unroll 1
    for (int b = 0; b < K; ++b) {
#pragma unroll 1
        for (int c4 = 0; c4 < C4; ++c4) {
            float4 xv[R];
#pragma unroll
            for (int r = 0; r < R; ++r) xv[r] = *reinterpret_cast<const float4*>(sx + c4 * PLANE + ((rg * 4 + r) * IW + lane + b) * 4);
            const float* wp = sw + (b * C4 + c4) * (K * 4 * COUT);             // broadcast loads
#pragma unroll
            for (int a = 0; a < K; ++a)
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float xe[4] = {xv[q + a].x, xv[q + a].y, xv[q + a].z, xv[q + a].w};
#pragma unroll
                    for (int e = 0; e < 4; ++e)
#pragma unroll
                        for (int j = 0; j < COUT; ++j) acc[q][j] = fmaf(xe[e], wp[(a * 4 + e) * COUT + j], acc[q][j]);
                }
        }
    }
    const int ox = tx0 + lane;
    if (ox >= p.dst.W) return;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int oy = ty0 + rg * 4 + q;
        if (oy >= p.dst.H) continue;
#pragma unroll
        for (int j = 0; j < COUT; ++j) {
            float t = acc[q][j];
            if (bias) t += __ldg(bias + (p.bias_mod > 0 ? j % p.bias_mod : j));
            t = act_direct(t, p.act, p.alpha);
            float* o = y + view_addr(p.dst, n, oy, ox, j);
            *o = p.accumulate ? *o + t : t;
        }
    }
}

// ------------------------------------------------------------------------------------------------ many input channels, <= 4 outputs
// The input gradient of the FAN front end (5x5, 32 -> 3 at 1280 x 128 x 128: 100 GFLOP with three outputs per pixel). With so few
// outputs per pixel a thread needs many pixels to amortise a weight operand: thread = 8 rows x 2 columns x COUT accumulators, the K x 4 x
// COUT weights of a (filter column, channel quad) live in registers while the thread streams its (8 + K - 1)-row input column through
// them: 39 LDS.128 per 960 FMAs (0.65 B per FMA; conv_manyin_kernel: 1.5). Persistent CTA per SM, 64 x 64-pixel tiles taken one channel
// quad at a time (74 KB), the next quad / tile in flight (cp.async into the other buffer) while this one is computed.
constexpr int M3W = 64, M3H = 64;

template <int CIN, int COUT, int K>
__global__ void __launch_bounds__(256, 1)
conv_manyin3_kernel(DirectParams p, const float* __restrict__ x, const float* __restrict__ wr, const float* __restrict__ bias, float* __restrict__ y) {
    static_assert(CIN % 4 == 0 && COUT <= 4 && (K * 4 * COUT) % 4 == 0, "channel quads; whole float4 weight runs");
    constexpr int C4 = CIN / 4, IH = M3H + K - 1, IW = M3W + K - 1, R = 8 + K - 1, WN = K * 4 * COUT;
    constexpr int PLANE = IH * IW * 4;
    extern __shared__ __align__(16) float smem[];
    float* sw = smem;                              // [K][C4][K][4][COUT]
    float* sx0 = smem + K * K * CIN * COUT;        // 2 x [IH][IW][4]
    const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
    const int tiles_x = (p.dst.W + M3W - 1) / M3W, tiles_y = (p.dst.H + M3H - 1) / M3H;
    const int tiles_total = tiles_x * tiles_y * p.n;
    const int my_tiles = (int)blockIdx.x < tiles_total ? (tiles_total - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int steps = my_tiles * C4;
    const bool vec_in = view_vec4(p.src);

    auto load_step = [&](int step, float* buf) {
        const int tile = blockIdx.x + (step / C4) * gridDim.x, c4 = step % C4;
        const int tx0 = (tile % tiles_x) * M3W, ty0 = ((tile / tiles_x) % tiles_y) * M3H, n = tile / (tiles_x * tiles_y);
        if (vec_in && p.pad_mode == NI_PAD_ZERO) {     // incremental indices, 32-bit offsets inside the image (see conv_fewin_kernel)
            const float* img = x + (long long)n * p.src.H * p.src.W * p.src.pitch + p.src.coff + c4 * 4;
            int px = tid % IW, py = tid / IW;
#pragma unroll 1
            while (py < IH) {
                const int sy = ty0 + py - p.pad_t, sxx = tx0 + px - p.pad_l;
                const bool in = (unsigned)sy < (unsigned)p.src.H && (unsigned)sxx < (unsigned)p.src.W;
                cp_async16_zfill(buf + (py * IW + px) * 4, in ? img + (sy * p.src.W + sxx) * p.src.pitch : x, in);
                px += 256 % IW; py += 256 / IW;
                if (px >= IW) { px -= IW; ++py; }
            }
            return;
        }
        for (int i = tid; i < IH * IW; i += 256) {
            const int px = i % IW, py = i / IW;
            int sy = ty0 + py - p.pad_t, sxx = tx0 + px - p.pad_l;
            if (p.pad_mode != NI_PAD_ZERO) { sy = mirror_idx(sy, p.src.H, p.pad_mode); sxx = mirror_idx(sxx, p.src.W, p.pad_mode); }
            const bool in = sy >= 0 && sy < p.src.H && sxx >= 0 && sxx < p.src.W;
            float* dst = buf + i * 4;
            if (vec_in) cp_async16_zfill(dst, in ? x + view_addr(p.src, n, sy, sxx, c4 * 4) : x, in);
            else {
#pragma unroll
                for (int e = 0; e < 4; ++e) cp_async4_zfill(dst + e, in ? x + view_addr(p.src, n, sy, sxx, c4 * 4 + e) : x, in);
            }
        }
    };

    for (int i = tid; i < K * K * CIN * COUT / 4; i += 256) cp_async16_zfill(sw + 4 * i, wr + 4 * i, true);
    if (steps > 0) load_step(0, sx0);

    float acc[8][2][COUT];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int j = 0; j < COUT; ++j) acc[r][h][j] = 0.f;

    for (int step = 0; step < steps; ++step) {
        cp_async_wait();
        __syncthreads();        // this quad (and the filter) landed; everybody is done with the buffer the next copy overwrites
        if (step + 1 < steps) load_step(step + 1, sx0 + ((step + 1) & 1) * PLANE);
        const float* sx = sx0 + (step & 1) * PLANE;
        const int c4 = step % C4;
#pragma unroll 1
        for (int b = 0; b < K; ++b) {
            float wv[WN];
            const float4* wp = reinterpret_cast<const float4*>(sw + (b * C4 + c4) * WN);      // broadcast loads
#pragma unroll
            for (int i = 0; i < WN / 4; ++i) { const float4 t = wp[i]; wv[4 * i] = t.x; wv[4 * i + 1] = t.y; wv[4 * i + 2] = t.z; wv[4 * i + 3] = t.w; }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                float4 xv[R];
#pragma unroll
                for (int r = 0; r < R; ++r) xv[r] = *reinterpret_cast<const float4*>(sx + ((wrp * 8 + r) * IW + lane + 32 * h + b) * 4);
#pragma unroll
                for (int a = 0; a < K; ++a)
#pragma unroll
                    for (int e = 0; e < 4; ++e)
#pragma unroll
                        for (int r = 0; r < 8; ++r) {
                            const float xe = e == 0 ? xv[r + a].x : (e == 1 ? xv[r + a].y : (e == 2 ? xv[r + a].z : xv[r + a].w));
#pragma unroll
                            for (int j = 0; j < COUT; ++j) acc[r][h][j] = fmaf(xe, wv[(a * 4 + e) * COUT + j], acc[r][h][j]);
                        }
            }
        }
        if (c4 != C4 - 1) continue;
        const int tile = blockIdx.x + (step / C4) * gridDim.x;
        const int tx0 = (tile % tiles_x) * M3W, ty0 = ((tile / tiles_x) % tiles_y) * M3H, n = tile / (tiles_x * tiles_y);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int oy = ty0 + wrp * 8 + r;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int ox = tx0 + lane + 32 * h;
#pragma unroll
                for (int j = 0; j < COUT; ++j) {
                    float t = acc[r][h][j];
                    acc[r][h][j] = 0.f;
                    if (oy >= p.dst.H || ox >= p.dst.W) continue;
                    if (bias) t += __ldg(bias + (p.bias_mod > 0 ? j % p.bias_mod : j));
                    t = act_direct(t, p.act, p.alpha);
                    float* o = y + view_addr(p.dst, n, oy, ox, j);
                    *o = p.accumulate ? *o + t : t;
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ wgrad
constexpr int WH = 8, WW = 32, WXS = 36;   // pixel tile; input tile row stride (8-wide windows starting at multiples of 4)

struct DirectWgradParams {
    TensorView xin, dyv;
    int n, pad_t, pad_l, pad_mode;
    int tiles_x, tiles_y, tiles_total;
};

template <int CIN, int COUT, int K, int COG>
__global__ void __launch_bounds__(256)
conv_direct_wgrad_kernel(DirectWgradParams p, const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dw) {
    static_assert(K <= 5 && COUT % COG == 0, "window of 4 + K - 1 <= 8 input columns");
    constexpr int XH = WH + K - 1;
    constexpr int TPS = K * CIN * (COUT / COG);                 // threads per pixel split
    constexpr int S = (256 / TPS) < WH ? (256 / TPS) : WH;      // splits: each takes rows s, s + S, ...
    static_assert(S >= 1, "too many (row, ci, co-group) combinations for one CTA");
    extern __shared__ __align__(16) float smem[];
    constexpr int XPL = XH * WXS + 4;                  // channel-plane stride (+4 floats: neighbouring channels on different banks)
    float* sx = smem;                                  // [CIN][XH][WXS]
    float* sd = smem + CIN * XPL;                      // [WH][WW][COUT]
    const int tid = threadIdx.x;
    const int s = tid / TPS, r = tid - s * TPS;
    const bool active = s < S;
    const int cog = r % (COUT / COG), ci = (r / (COUT / COG)) % CIN, a = r / ((COUT / COG) * CIN);
    float acc[K][COG];
#pragma unroll
    for (int b = 0; b < K; ++b)
#pragma unroll
        for (int j = 0; j < COG; ++j) acc[b][j] = 0.f;

    for (int tile = blockIdx.x; tile < p.tiles_total; tile += gridDim.x) {
        const int tx0 = (tile % p.tiles_x) * WW, ty0 = ((tile / p.tiles_x) % p.tiles_y) * WH, n = tile / (p.tiles_x * p.tiles_y);
        __syncthreads();
        if constexpr (CIN % 4 == 0) if (view_vec4(p.xin)) {
            // many channels: 128-bit loads (8 in flight per thread), scattered to the channel planes (4-byte async copies of
            // 32 channels per pixel were slower: 1.58 vs 1.09 ms for the U-Net 32->12 wgrad)
            for (int i = tid; i < XH * WXS; i += 256) {
                const int px = i % WXS, py = i / WXS;
                int sy = ty0 + py - p.pad_t, sxx = tx0 + px - p.pad_l;
                if (p.pad_mode != NI_PAD_ZERO) { sy = mirror_idx(sy, p.xin.H, p.pad_mode); sxx = mirror_idx(sxx, p.xin.W, p.pad_mode); }
                const bool in = px < WW + K - 1 && sy >= 0 && sy < p.xin.H && sxx >= 0 && sxx < p.xin.W;
                float4 v[CIN / 4 > 0 ? CIN / 4 : 1];
#pragma unroll
                for (int c4 = 0; c4 < CIN / 4; ++c4)
                    v[c4] = in ? __ldg(reinterpret_cast<const float4*>(x + view_addr(p.xin, n, sy, sxx, c4 * 4))) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int c4 = 0; c4 < CIN / 4; ++c4) {
                    float* d = sx + (c4 * 4) * XPL + py * WXS + px;
                    d[0] = v[c4].x; d[XPL] = v[c4].y; d[2 * XPL] = v[c4].z; d[3 * XPL] = v[c4].w;
                }
            }
        }
        if (!(CIN % 4 == 0 && view_vec4(p.xin)))
        for (int i = tid; i < XH * WXS * CIN; i += 256) {        // lanes = channels: coalesced reads of the NHWC rows
            const int c = i % CIN, px = (i / CIN) % WXS, py = i / (CIN * WXS);
            int sy = ty0 + py - p.pad_t, sxx = tx0 + px - p.pad_l;
            if (p.pad_mode != NI_PAD_ZERO) { sy = mirror_idx(sy, p.xin.H, p.pad_mode); sxx = mirror_idx(sxx, p.xin.W, p.pad_mode); }
            const bool in = px < WW + K - 1 && sy >= 0 && sy < p.xin.H && sxx >= 0 && sxx < p.xin.W;
            cp_async4_zfill(sx + c * XPL + py * WXS + px, in ? x + view_addr(p.xin, n, sy, sxx, c) : x, in);
        }
        constexpr int CO4 = COUT / 4 > 0 ? COUT / 4 : 1;
        if (COUT % 4 == 0 && view_vec4(p.dyv)) {
            for (int i = tid; i < WH * WW * CO4; i += 256) {
                const int c4 = i % CO4, px = (i / CO4) % WW, py = i / (CO4 * WW);
                const int oy = ty0 + py, ox = tx0 + px;
                const bool in = oy < p.dyv.H && ox < p.dyv.W;
                cp_async16_zfill(sd + 4 * i, in ? dy + view_addr(p.dyv, n, oy, ox, c4 * 4) : dy, in);
            }
        } else {
            for (int i = tid; i < WH * WW * COUT; i += 256) {
                const int c = i % COUT, px = (i / COUT) % WW, py = i / (COUT * WW);
                const int oy = ty0 + py, ox = tx0 + px;
                const bool in = oy < p.dyv.H && ox < p.dyv.W;
                cp_async4_zfill(sd + i, in ? dy + view_addr(p.dyv, n, oy, ox, c) : dy, in);
            }
        }
        cp_async_wait();
        __syncthreads();
        if (!active) continue;
#pragma unroll 1
        for (int yy = s; yy < WH; yy += S) {
            const float* xr = sx + ci * XPL + (yy + a) * WXS;
            const float* dr = sd + (yy * WW) * COUT + cog * COG;
#pragma unroll 2
            for (int xq = 0; xq < WW; xq += 4) {
                const float4 x0 = *reinterpret_cast<const float4*>(xr + xq), x1 = *reinterpret_cast<const float4*>(xr + xq + 4);
                const float xv[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    float g[COG];
                    if (COG % 4 == 0 && COUT % 4 == 0) {
#pragma unroll
                        for (int j4 = 0; j4 < COG / 4; ++j4) {
                            const float4 f = *reinterpret_cast<const float4*>(dr + (xq + e) * COUT + j4 * 4);
                            g[4 * j4] = f.x; g[4 * j4 + 1] = f.y; g[4 * j4 + 2] = f.z; g[4 * j4 + 3] = f.w;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < COG; ++j) g[j] = dr[(xq + e) * COUT + j];
                    }
#pragma unroll
                    for (int b = 0; b < K; ++b)
#pragma unroll
                        for (int j = 0; j < COG; ++j) acc[b][j] = fmaf(xv[e + b], g[j], acc[b][j]);
                }
            }
        }
    }
    // reduce the S pixel splits inside the CTA (through the idle tile memory), then ONE atomicAdd per output per CTA: with
    // per-thread atomics the 3->3 filter (225 addresses) took 1.6 M contended L2 atomics (0.41 ms for 0.04 ms of math)
    __syncthreads();
    float* red = smem;                                 // [S - 1][K * COG][TPS]
    if (active && s > 0) {
#pragma unroll
        for (int b = 0; b < K; ++b)
#pragma unroll
            for (int j = 0; j < COG; ++j) red[((s - 1) * (K * COG) + b * COG + j) * TPS + r] = acc[b][j];
    }
    __syncthreads();
    if (!active || s > 0) return;
#pragma unroll
    for (int b = 0; b < K; ++b)
#pragma unroll
        for (int j = 0; j < COG; ++j) {
            float v = acc[b][j];
            for (int t = 0; t < S - 1; ++t) v += red[(t * (K * COG) + b * COG + j) * TPS + r];
            atomicAdd(dw + (((a * K + b) * CIN) + ci) * COUT + cog * COG + j, v);
        }
}

template <int CIN, int COUT, int K, bool POOL = false>
int launch_fewin(const DirectParams& p, const float* x, const float* w, const float* bias, float* y, cudaStream_t st) {
    const size_t smem = sizeof(float) * (K * K * CIN * COUT + 2 * CIN * (FH + K - 1) * FXS);
    NI_CUDA(cudaFuncSetAttribute(conv_fewin_kernel<CIN, COUT, K, POOL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int tiles = ((p.dst.W + FW - 1) / FW) * ((p.dst.H + FH - 1) / FH) * p.n;
    const int grid = tiles < 2 * ni_num_sms() ? tiles : 2 * ni_num_sms();           // persistent: two CTAs per SM (<= 128 registers)
    conv_fewin_kernel<CIN, COUT, K, POOL><<<grid, 256, smem, st>>>(p, x, w, bias, y);
    NI_LAUNCH_CHECK(); NI_COUNT_LAUNCH(1);
    return NI_OK;
}

template <int CIN, int COUT, int K>
int launch_manyin(const DirectParams& p, const float* x, const float* wr, const float* bias, float* y, cudaStream_t st) {
    if constexpr (COUT <= 4) {
        // 64 x 64 tiles, persistent (conv_manyin3_kernel) when the tile grid is mostly real pixels. Measured and NOT used: layers with 12
        // outputs as three groups of 4 output channels: 3 x 0.92 ms against 1.53 ms of conv_manyin_kernel for the U-Net output layer --
        // with K = 3 a channel quad is consumed so fast that the 128-byte pixel rows it was cut from leave L2 before the next quad needs
        // them (each line fetched up to 8 times).
        const long long covered = (long long)((p.dst.W + M3W - 1) / M3W) * M3W * ((p.dst.H + M3H - 1) / M3H) * M3H;
        if (4 * (long long)p.dst.W * p.dst.H >= 3 * covered) {
            const size_t smem3 = sizeof(float) * (K * K * CIN * COUT + 2 * (M3H + K - 1) * (M3W + K - 1) * 4);
            NI_CUDA(cudaFuncSetAttribute(conv_manyin3_kernel<CIN, COUT, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3));
            const int tiles3 = ((p.dst.W + M3W - 1) / M3W) * ((p.dst.H + M3H - 1) / M3H) * p.n;
            const int grid = tiles3 < ni_num_sms() ? tiles3 : ni_num_sms();
            conv_manyin3_kernel<CIN, COUT, K><<<grid, 256, smem3, st>>>(p, x, wr, bias, y);
            NI_LAUNCH_CHECK(); NI_COUNT_LAUNCH(1);
            return NI_OK;
        }
    }
    const size_t smem = sizeof(float) * (K * K * CIN * COUT + (CIN / 4) * ((MH + K - 1) * (MW + K - 1) * 4 + 4));
    NI_CUDA(cudaFuncSetAttribute(conv_manyin_kernel<CIN, COUT, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int tiles = ((p.dst.W + MW - 1) / MW) * ((p.dst.H + MH - 1) / MH) * p.n;
    conv_manyin_kernel<CIN, COUT, K><<<tiles, 128, smem, st>>>(p, x, wr, bias, y);
    NI_LAUNCH_CHECK(); NI_COUNT_LAUNCH(1);
    return NI_OK;
}

template <int CIN, int COUT, int K, int COG>
int launch_wgrad(DirectWgradParams p, const float* x, const float* dy, float* dw, cudaStream_t st) {
    constexpr int TPS = K * CIN * (COUT / COG), S = (256 / TPS) < WH ? (256 / TPS) : WH;
    static_assert((S - 1) * K * COG * TPS <= CIN * ((WH + K - 1) * WXS + 4) + WH * WW * COUT, "split reduction fits the tile memory");
    const size_t smem = sizeof(float) * (CIN * ((WH + K - 1) * WXS + 4) + WH * WW * COUT);
    NI_CUDA(cudaFuncSetAttribute(conv_direct_wgrad_kernel<CIN, COUT, K, COG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    p.tiles_x = (p.dyv.W + WW - 1) / WW; p.tiles_y = (p.dyv.H + WH - 1) / WH; p.tiles_total = p.tiles_x * p.tiles_y * p.n;
    int per_sm = (int)((200u << 10) / (smem + 1024));
    if (per_sm > 6) per_sm = 6;
    if (per_sm < 1) per_sm = 1;
    const int grid = p.tiles_total < per_sm * ni_num_sms() ? p.tiles_total : per_sm * ni_num_sms();
    conv_direct_wgrad_kernel<CIN, COUT, K, COG><<<grid, 256, smem, st>>>(p, x, dy, dw);
    NI_LAUNCH_CHECK(); NI_COUNT_LAUNCH(1);
    return NI_OK;
}

// (cin, cout, k) of the logical FORWARD convolution each kernel family is instantiated for
#define NI_FEWIN_SHAPES(X) X(3, 32, 5) X(4, 32, 3)
#define NI_MANYIN_SHAPES(X) X(32, 3, 5) X(32, 12, 3) X(32, 4, 3) X(64, 12, 3)
#define NI_DWGRAD_SHAPES(X) X(3, 32, 5, 8) X(4, 32, 3, 8) X(32, 12, 3, 12) X(3, 3, 5, 3) X(64, 12, 3, 12)

}  // namespace

int ni_get_scratch2(size_t bytes, float** out);

// op: 0 fprop, 1 dgrad, 2 wgrad. Only stride 1, square filters, plain input addressing.
extern "C" int ni_conv2d_direct_supported(const ni_conv_desc* d, int op) {
    if (!d || d->n <= 0 || d->stride != 1 || d->kh != d->kw || d->in_mode != NI_MODE_PLAIN) return 0;
#define X(a, b, c) if (op == 0 && d->cin == a && d->cout == b && d->kh == c) return 1;
    NI_FEWIN_SHAPES(X) NI_MANYIN_SHAPES(X)
#undef X
#define X(a, b, c) if (op == 1 && d->pad_mode == NI_PAD_ZERO && d->cout == a && d->cin == b && d->kh == c) return 1;   /* dgrad: roles swapped */
    NI_MANYIN_SHAPES(X)
#undef X
#define X(a, b, c, g) if (op == 2 && d->cin == a && d->cout == b && d->kh == c) return 1;
    NI_DWGRAD_SHAPES(X)
#undef X
    return 0;
}

static int reorder_weights(const float* w, int k, int cin, int cout, int flip, float** out, cudaStream_t st) {
    const int total = k * k * cin * cout;
    int rc = ni_get_scratch2(sizeof(float) * (size_t)total, out);
    if (rc) return rc;
    reorder_manyin_weights_kernel<<<ni_cdiv(total, 256), 256, 0, st>>>(w, *out, k, cin, cout, flip);
    NI_LAUNCH_CHECK(); NI_COUNT_LAUNCH(1);
    return NI_OK;
}

extern "C" int ni_conv2d_fprop_direct(const ni_conv_desc* d, const float* x, const float* w, const float* bias, float* y, cudaStream_t st) {
    NI_REQUIRE(ni_conv2d_direct_supported(d, 0) && x && w && y, "ni_conv2d_fprop_direct: unsupported problem or null pointer");
    DirectParams p;
    p.src = TensorView{d->h, d->w, d->cin, d->in_pitch, d->in_coff, d->in_mode};
    p.dst = TensorView{d->oh, d->ow, d->cout, d->out_pitch, d->out_coff, d->out_mode};
    p.n = d->n; p.pad_t = d->pad_t; p.pad_l = d->pad_l; p.pad_mode = d->pad_mode;
    p.act = d->act; p.bias_mod = d->bias_mod; p.alpha = d->act_alpha; p.accumulate = d->accumulate;
    p.pool = nullptr; p.code = nullptr;
#define X(a, b, c) if (d->cin == a && d->cout == b && d->kh == c) return launch_fewin<a, b, c>(p, x, w, bias, y, st);
    NI_FEWIN_SHAPES(X)
#undef X
    float* wr = nullptr;
    int rc = reorder_weights(w, d->kh, d->cin, d->cout, 0, &wr, st);
    if (rc) return rc;
#define X(a, b, c) if (d->cin == a && d->cout == b && d->kh == c) return launch_manyin<a, b, c>(p, x, wr, bias, y, st);
    NI_MANYIN_SHAPES(X)
#undef X
    return NI_ERR_UNSUPPORTED;
}

// Conv2D(few input channels) + bias + activation + MaxPool2D(2x2) in one kernel (models/forensics.py:66-69, first block of the FAN):
// writes the pooled activation and one code byte per pooled element instead of the full-resolution activation (2.7 GB at 1280 images
// of 128x128x32 that the stand-alone pooling kernel read back and the pooling backward read again).
extern "C" int ni_conv2d_pool2_supported(const ni_conv_desc* d) {
    if (!d || d->n <= 0 || d->stride != 1 || d->kh != d->kw || d->in_mode != NI_MODE_PLAIN || d->out_mode != NI_MODE_PLAIN) return 0;
    if ((d->oh & 1) || (d->ow & 1) || d->accumulate || d->bias_mod || d->out_coff || d->out_pitch != d->cout) return 0;
    if (d->act != NI_ACT_LEAKY_RELU && d->act != NI_ACT_RELU) return 0;           // the code byte stores the slope as "value > 0"
#define X(a, b, c) if (d->cin == a && d->cout == b && d->kh == c) return 1;
    NI_FEWIN_SHAPES(X)
#undef X
    return 0;
}
extern "C" int ni_conv2d_pool2_fwd(const ni_conv_desc* d, const float* x, const float* w, const float* bias, float* pooled,
                                   unsigned char* code, cudaStream_t st) {
    NI_REQUIRE(ni_conv2d_pool2_supported(d) && x && w && pooled && code, "ni_conv2d_pool2_fwd: unsupported problem or null pointer");
    DirectParams p;
    p.src = TensorView{d->h, d->w, d->cin, d->in_pitch, d->in_coff, d->in_mode};
    p.dst = TensorView{d->oh, d->ow, d->cout, d->out_pitch, d->out_coff, d->out_mode};
    p.n = d->n; p.pad_t = d->pad_t; p.pad_l = d->pad_l; p.pad_mode = d->pad_mode;
    p.act = d->act; p.bias_mod = 0; p.alpha = d->act_alpha; p.accumulate = 0;
    p.pool = pooled; p.code = code;
#define X(a, b, c) if (d->cin == a && d->cout == b && d->kh == c) return launch_fewin<a, b, c, true>(p, x, w, bias, nullptr, st);
    NI_FEWIN_SHAPES(X)
#undef X
    return NI_ERR_UNSUPPORTED;
}

// dx = forward conv of dy (cout channels) with the flipped / channel-swapped filter (stride 1, zero padding K-1-pad).
extern "C" int ni_conv2d_dgrad_direct(const ni_conv_desc* d, const float* dy, const float* w, float* dx, cudaStream_t st) {
    NI_REQUIRE(ni_conv2d_direct_supported(d, 1) && dy && w && dx, "ni_conv2d_dgrad_direct: unsupported problem or null pointer");
    const int k = d->kh;
    float* wr = nullptr;
    // w is (k, k, cin, cout); the "forward" conv of dgrad has cin' = cout, cout' = cin
    int rc = reorder_weights(w, k, d->cout, d->cin, 1, &wr, st);
    if (rc) return rc;
    DirectParams p;
    p.src = TensorView{d->oh, d->ow, d->cout, d->out_pitch, d->out_coff, d->out_mode};
    p.dst = TensorView{d->h, d->w, d->cin, d->in_pitch, d->in_coff, d->in_mode};
    p.n = d->n; p.pad_t = k - 1 - d->pad_t; p.pad_l = k - 1 - d->pad_l; p.pad_mode = NI_PAD_ZERO;
    p.act = NI_ACT_NONE; p.bias_mod = 0; p.alpha = 0.f; p.accumulate = d->accumulate;
    p.pool = nullptr; p.code = nullptr;
#define X(a, b, c) if (d->cout == a && d->cin == b && d->kh == c) return launch_manyin<a, b, c>(p, dy, wr, nullptr, dx, st);
    NI_MANYIN_SHAPES(X)
#undef X
    return NI_ERR_UNSUPPORTED;
}

extern "C" int ni_conv2d_wgrad_direct(const ni_conv_desc* d, const float* x, const float* dy, float* dw, cudaStream_t st) {
    NI_REQUIRE(ni_conv2d_direct_supported(d, 2) && x && dy && dw, "ni_conv2d_wgrad_direct: unsupported problem or null pointer");
    if (!d->accumulate) NI_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)d->kh * d->kw * d->cin * d->cout, st));
    DirectWgradParams p;
    p.xin = TensorView{d->h, d->w, d->cin, d->in_pitch, d->in_coff, d->in_mode};
    p.dyv = TensorView{d->oh, d->ow, d->cout, d->out_pitch, d->out_coff, d->out_mode};
    p.n = d->n; p.pad_t = d->pad_t; p.pad_l = d->pad_l; p.pad_mode = d->pad_mode;
#define X(a, b, c, g) if (d->cin == a && d->cout == b && d->kh == c) return launch_wgrad<a, b, c, g>(p, x, dy, dw, st);
    NI_DWGRAD_SHAPES(X)
#undef X
    return NI_ERR_UNSUPPORTED;
}
