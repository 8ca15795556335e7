// Fused differentiable-JPEG kernels (forward and backward) for sm_100a.
//
// Replaces the whole of DifferentiableJPEG.call (reference models/jpeg.py:91-159): colour transform ->
// 8x8 blocking -> DCT (literal 4-decimal matrix, jpeg.py:78-85) -> /Q -> Quantization (models/layers.py:118-128)
// -> *Q -> IDCT -> inverse blocking -> colour transform -> /255 -> clip, in ONE kernel (the reference materialises
// >= 25 full-size temporaries). HBM traffic = read x + write y (24 B/pixel) forward; read x, dy + write dx backward
// (36 B/pixel, X/Q is recomputed instead of saved).
//
// Work decomposition: a tile = 32 8x8 pixel blocks (24.5 KB of interleaved RGB staged in shared memory); one thread owns one
// (block, Y/Cb/Cr channel) pair (warp = channel, lane = block) and keeps the whole 8x8 coefficient block in registers, so the 2-D DCT /
// quantisation / IDCT need no inter-thread exchange. The 1-D transforms use the even/odd symmetry that the literal rows preserve (36
// instead of 64 FMA-class ops per 8-point transform) with the literal coefficients as FFMA immediates. The colour transforms (which mix
// channels) run in cooperative per-pixel passes over the shared-memory tile.
// Forward: generation 4 (djpeg_fwd4_kernel) — PERSISTENT CTAs (4 per SM) whose tiles arrive as one TMA box each (cp.async.bulk.tensor
// into a dense stage, mbarrier completion), issued by a producer warp while the other warps transform the previous tile; generation 3
// (djpeg_fwd3_kernel: one tile per CTA, 16-byte cp.async staging) remains for widths whose block grid does not tile into rectangular
// 32-block boxes (W / 8 not a multiple of 4). Backward / table gradient: generation 3 structure.
#include "ni_common.cuh"
#include "tc_common.cuh"

int ni_encode_tiled_sw(CUtensorMap* tm, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes, const cuuint32_t* box,
                       int swizzle);   // conv_tc.cu

namespace {

constexpr int kTileBlocks = 32;                 // 8x8 pixel blocks per CTA tile
constexpr int kBlockFloats = 196;               // 8*8*3 = 192 payload + 4 pad: the 16-byte bank group rotates from block to
                                                // block, which removes the 11-way shared-memory conflicts of a 192-word stride
constexpr int kThreads = kTileBlocks * 3;       // one thread per (block, channel)
constexpr int kTileFloats = kTileBlocks * kBlockFloats;
constexpr int kUnitsPerBlock = 16;              // 4-pixel (3 x float4) units per block in the per-pixel passes

struct DjpegTables {
    float q[2][64];   // [0] luma, [1] chroma (reference compression/jpeg_helpers.py:264-305), row-major [k][l]
    float rq[2][64];  // reciprocals
};

// DCT literals, reference models/jpeg.py:78-85.
#define C0 0.3536f
#define C1 0.4904f
#define C2 0.4619f
#define C3 0.4157f
#define C5 0.2778f
#define C6 0.1913f
#define C7 0.0975f

// out[k] = sum_j F[k][j] a[j]
__device__ __forceinline__ void dct8_fwd(float& a0, float& a1, float& a2, float& a3, float& a4, float& a5, float& a6,
                                         float& a7) {
    const float s0 = a0 + a7, s1 = a1 + a6, s2 = a2 + a5, s3 = a3 + a4;
    const float d0 = a0 - a7, d1 = a1 - a6, d2 = a2 - a5, d3 = a3 - a4;
    const float ss0 = s0 + s3, ss1 = s1 + s2, sd0 = s0 - s3, sd1 = s1 - s2;
    a0 = C0 * (ss0 + ss1);
    a4 = C0 * (ss0 - ss1);
    a2 = fmaf(C2, sd0, C6 * sd1);
    a6 = fmaf(C6, sd0, -C2 * sd1);
    a1 = fmaf(C1, d0, fmaf(C3, d1, fmaf(C5, d2, C7 * d3)));
    a3 = fmaf(C3, d0, fmaf(-C7, d1, fmaf(-C1, d2, -C5 * d3)));
    a5 = fmaf(C5, d0, fmaf(-C1, d1, fmaf(C7, d2, C3 * d3)));
    a7 = fmaf(C7, d0, fmaf(-C5, d1, fmaf(C3, d2, -C1 * d3)));
}

// a[j] = sum_k F[k][j] X[k]
__device__ __forceinline__ void dct8_inv(float& x0, float& x1, float& x2, float& x3, float& x4, float& x5, float& x6,
                                         float& x7) {
    const float p = C0 * (x0 + x4), m = C0 * (x0 - x4);
    const float t = fmaf(C2, x2, C6 * x6), u = fmaf(C6, x2, -C2 * x6);
    const float e0 = p + t, e3 = p - t, e1 = m + u, e2 = m - u;
    const float o0 = fmaf(C1, x1, fmaf(C3, x3, fmaf(C5, x5, C7 * x7)));
    const float o1 = fmaf(C3, x1, fmaf(-C7, x3, fmaf(-C1, x5, -C5 * x7)));
    const float o2 = fmaf(C5, x1, fmaf(-C1, x3, fmaf(C7, x5, C3 * x7)));
    const float o3 = fmaf(C7, x1, fmaf(-C5, x3, fmaf(C3, x5, -C1 * x7)));
    x0 = e0 + o0; x7 = e0 - o0;
    x1 = e1 + o1; x6 = e1 - o1;
    x2 = e2 + o2; x5 = e2 - o2;
    x3 = e3 + o3; x4 = e3 - o3;
}

// X = F r F^T  (rows then columns), in place.
__device__ __forceinline__ void dct2d_fwd(float (&v)[8][8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) dct8_fwd(v[i][0], v[i][1], v[i][2], v[i][3], v[i][4], v[i][5], v[i][6], v[i][7]);
#pragma unroll
    for (int j = 0; j < 8; ++j) dct8_fwd(v[0][j], v[1][j], v[2][j], v[3][j], v[4][j], v[5][j], v[6][j], v[7][j]);
}
// x = F^T X F, in place.
__device__ __forceinline__ void dct2d_inv(float (&v)[8][8]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) dct8_inv(v[0][j], v[1][j], v[2][j], v[3][j], v[4][j], v[5][j], v[6][j], v[7][j]);
#pragma unroll
    for (int i = 0; i < 8; ++i) dct8_inv(v[i][0], v[i][1], v[i][2], v[i][3], v[i][4], v[i][5], v[i][6], v[i][7]);
}

// Quantisation modes (reference models/layers.py:118-136; 'harmonic' only ever has its first term active, see
// SURVEY 8a a11). 0 = soft, 1 = sin, 2 = harmonic.
constexpr float kTwoPi = 6.2831855f;  // float32(2*np.pi), as TF casts the python scalar
constexpr float kPi = 3.1415927f;

// sin(2 pi z) through the period-1 identity: r = z - rint(z) is exact in float32 and lies in [-0.5, 0.5], where the
// special-function unit is accurate to < 5e-7 (the plain sinf(2 pi z) loses ulp(2 pi z) and takes a ~40-instruction path).
__device__ __forceinline__ float sin_2pi(float z) { return __sinf(kTwoPi * (z - ni_round_he(z))); }
template <int MODE>
__device__ __forceinline__ float quant_fwd(float z) {
    if (MODE == 0) return ni_round_he(z);
    if (MODE == 1) return z - sin_2pi(z) * (1.f / kTwoPi);
    return z - sin_2pi(z) * (1.f / kPi);
}
// Colour constants. Forward rows of _color_F with the 255 input scale and the -127 level shift folded in; inverse
// rows of _color_I with the +127 shift and the /255 folded in (reference models/jpeg.py:74-75,99-105,154-156).
__device__ __forceinline__ void color_fwd_coeffs(int c, float& k0, float& kr, float& kg, float& kb) {
    if (c == 0) { k0 = 0.f - 127.f; kr = 255.f * 0.299f; kg = 255.f * 0.587f; kb = 255.f * 0.114f; }
    else if (c == 1) { k0 = 128.f - 127.f; kr = 255.f * -0.168736f; kg = 255.f * -0.331264f; kb = 255.f * 0.5f; }
    else { k0 = 128.f - 127.f; kr = 255.f * 0.5f; kg = 255.f * -0.418688f; kb = 255.f * -0.081312f; }
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const unsigned sa = static_cast<unsigned>(__cvta_generic_to_shared(smem));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// Float offset of (block g, row 0, col 0) in an (N,H,W,3) tensor; computed once per block per tile (32-bit divisions: nblk < 2^31).
__device__ __forceinline__ long long block_base_offset(int g, int nbw, int nbh, int W) {
    const int bx = g % nbw, t = g / nbw;
    const int by = t % nbh, n = t / nbh;
    return ((((long long)n * nbh + by) * 8) * (long long)W + bx * 8) * 3;
}
// Stage a tile (block-major [blk][row][24], block stride 196) global -> shared with 16-byte async copies. A thread owns
// (block, 16-byte column) pairs and walks down the 8 rows with pointer increments (3 instructions per copy instead of
// ~15 of index arithmetic); consecutive threads still copy consecutive 16-byte chunks of one image row.
__device__ __forceinline__ void tile_load(float* tile, const float* __restrict__ src, const long long* base, int rowf) {
    for (int pr = threadIdx.x; pr < kTileBlocks * 6; pr += kThreads) {
        const int blk = pr / 6, f4 = pr - blk * 6;
        const long long b = base[blk];
        if (b < 0) continue;
        const float* g = src + b + f4 * 4;
        float* d = tile + blk * kBlockFloats + f4 * 4;
#pragma unroll
        for (int row = 0; row < 8; ++row) cp_async16(d + row * 24, g + (long long)row * rowf);
    }
}
// Load this thread's channel of its block (level-shifted YCbCr) from the interleaved RGB tile.
__device__ __forceinline__ void load_channel(const float* bp, int c, float (&v)[8][8]) {
    float k0, kr, kg, kb;
    color_fwd_coeffs(c, k0, kr, kg, kb);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float f[24];
#pragma unroll
        for (int m = 0; m < 6; ++m) {
            const float4 p = *reinterpret_cast<const float4*>(bp + i * 24 + m * 4);
            f[m * 4 + 0] = p.x; f[m * 4 + 1] = p.y; f[m * 4 + 2] = p.z; f[m * 4 + 3] = p.w;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) v[i][j] = fmaf(kb, f[3 * j + 2], fmaf(kg, f[3 * j + 1], fmaf(kr, f[3 * j], k0)));
    }
}
// ---------------------------------------------------------------------------------------------------- forward, generation 3
// ncu on generation 2 (profiles/r1_djpeg_v2_ncu.txt): 113 thread-instructions per pixel, issue-active 51 %, 16.6 warps/SM,
// stalls = short scoreboard (shared-memory loads: 64 table LDS.64 + 48 tile LDS.128 + 64 STS.32 per thread), barrier, MIO
// throttle; 17 % of the shared wavefronts were bank conflicts (store_channel). Changes:
//  * warp = channel (Y | Cb | Cr), lane = 8x8 block: the quantisation table is warp-uniform, so Q and 1/Q come from the
//    kernel-parameter constant bank as immediate operands of the FMULs (no table loads at all);
//  * the IDCT result goes back to shared memory PLANAR ([block][channel][8][8], 16 conflict-free STS.128 instead of 64
//    STS.32) and the colour pass reads three LDS.128 per 4 pixels;
//  * inverse colour transform with the +127 / -k*128 / 1/255 constants folded: 7 FFMA(.SAT) per pixel instead of 16 ops;
//  * 80 registers -> 8 CTAs (24 warps) per SM instead of 6.
// Measured and dropped: cp.async.bulk.prefetch.L2 of the tile a later CTA will read (0.117 -> 0.159 ms), packed FFMA2 arithmetic
// (half issue rate on sm_100a: tools/probes/ffma2_probe.cu).
template <int MODE, int CC>
__device__ __forceinline__ void quantise_block(float (&v)[8][8], const DjpegTables& tab) {
#pragma unroll
    for (int k = 0; k < 8; ++k)
#pragma unroll
        for (int l = 0; l < 8; ++l) v[k][l] = quant_fwd<MODE>(v[k][l] * tab.rq[CC][k * 8 + l]) * tab.q[CC][k * 8 + l];
}

// ypre for 4 pixels of one row from planar level-shifted Y / Cb / Cr, clipped to [0,1] (reference models/jpeg.py:154-157)
__device__ __forceinline__ void color_inv_clip4(const float4 Y, const float4 B, const float4 R, float (&f)[12]) {
    constexpr float s = 1.f / 255.f;
    constexpr float kr = (127.f - 1.402f) * s, kg = (127.f + 1.058272f) * s, kb = (127.f - 1.772f) * s;
    const float yy[4] = {Y.x, Y.y, Y.z, Y.w}, cb[4] = {B.x, B.y, B.z, B.w}, cr[4] = {R.x, R.y, R.z, R.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        f[3 * j] = __saturatef(fmaf(cr[j], 1.402f * s, fmaf(yy[j], s, kr)));
        f[3 * j + 1] = __saturatef(fmaf(cr[j], -0.714136f * s, fmaf(cb[j], -0.344136f * s, fmaf(yy[j], s, kg))));
        f[3 * j + 2] = __saturatef(fmaf(cb[j], 1.772f * s, fmaf(yy[j], s, kb)));
    }
}

template <int MODE, bool WRITE_X>
__global__ void __launch_bounds__(kThreads, 8)
djpeg_fwd3_kernel(const float* __restrict__ x, float* __restrict__ y, float* __restrict__ Xd, int H, int W, int nblk,
                  const __grid_constant__ DjpegTables tab) {
    extern __shared__ __align__(16) float smem[];
    float* tile = smem;
    long long* base = reinterpret_cast<long long*>(smem + kTileFloats);
    const int nbw = W / 8, nbh = H / 8, rowf = W * 3;
    const int g0 = blockIdx.x * kTileBlocks;
    const int c = threadIdx.x >> 5, blk = threadIdx.x & 31;

    if (c == 0) {
        const int g = g0 + blk;
        base[blk] = g < nblk ? block_base_offset(g, nbw, nbh, W) : -1;
    }
    __syncthreads();
    tile_load(tile, x, base, rowf);
    cp_async_wait_all();
    __syncthreads();

    float* bp = tile + blk * kBlockFloats;
    float v[8][8];
    load_channel(bp, c, v);
    __syncthreads();  // all three channel warps have read the RGB data before the tile is overwritten

    dct2d_fwd(v);
    if (c == 0) quantise_block<MODE, 0>(v, tab); else quantise_block<MODE, 1>(v, tab);
    if (WRITE_X) {
        // de-quantised coefficients, reference block order ((n*3+c)*nb + by*nbw + bx, k, l)  (models/jpeg.py:105-114,159)
        const int g = g0 + blk;
        if (g < nblk) {
            const int nb = nbw * nbh;
            const long long n = g / nb, r = g % nb;
            float4* o = reinterpret_cast<float4*>(Xd + ((n * 3 + c) * nb + r) * 64);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                o[k * 2 + 0] = make_float4(v[k][0], v[k][1], v[k][2], v[k][3]);
                o[k * 2 + 1] = make_float4(v[k][4], v[k][5], v[k][6], v[k][7]);
            }
        }
    }
    dct2d_inv(v);
    {   // planar store: [block][channel][row][col]
        float4* pp = reinterpret_cast<float4*>(bp + c * 64);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            pp[2 * i] = make_float4(v[i][0], v[i][1], v[i][2], v[i][3]);
            pp[2 * i + 1] = make_float4(v[i][4], v[i][5], v[i][6], v[i][7]);
        }
    }
    __syncthreads();

    // colour pass: unit u = (block u/16, row (u%16)/2, half u%2) = 4 pixels = 48 contiguous output bytes
    for (int u = threadIdx.x; u < kTileBlocks * kUnitsPerBlock; u += kThreads) {
        const int b = u >> 4, r = u & 15;
        const long long bo = base[b];
        if (bo < 0) continue;
        const float4* p = reinterpret_cast<const float4*>(tile + b * kBlockFloats + r * 4);
        float f[12];
        color_inv_clip4(p[0], p[16], p[32], f);
        float4* o = reinterpret_cast<float4*>(y + bo + (long long)(r >> 1) * rowf + (r & 1) * 12);
        __stcs(o, make_float4(f[0], f[1], f[2], f[3]));
        __stcs(o + 1, make_float4(f[4], f[5], f[6], f[7]));
        __stcs(o + 2, make_float4(f[8], f[9], f[10], f[11]));
    }
}

__device__ __forceinline__ void store_plane(float* plane, const float (&v)[8][8]) {
    float4* pp = reinterpret_cast<float4*>(plane);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        pp[2 * i] = make_float4(v[i][0], v[i][1], v[i][2], v[i][3]);
        pp[2 * i + 1] = make_float4(v[i][4], v[i][5], v[i][6], v[i][7]);
    }
}
__device__ __forceinline__ void load_plane(const float* plane, float (&v)[8][8]) {
    const float4* pp = reinterpret_cast<const float4*>(plane);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float4 a = pp[2 * i], b = pp[2 * i + 1];
        v[i][0] = a.x; v[i][1] = a.y; v[i][2] = a.z; v[i][3] = a.w;
        v[i][4] = b.x; v[i][5] = b.y; v[i][6] = b.z; v[i][7] = b.w;
    }
}

// ---------------------------------------------------------------------------------------------------- forward, generation 4
// Persistent CTAs fed by TMA boxes (BASELINE north_star: "TMA staging into shared memory"). The image tensor is described to the copy
// engine as a 3-D tensor (24 K floats = the 8-pixel rows of K neighbouring blocks | W / 8 / K | N*H pixel rows; K = 8 or 4, see
// ni_djpeg_fwd); a tile of TBX x TBY = 32 blocks is ONE cp.async.bulk.tensor box that lands densely ([pixel row][block][24]) in a stage
// and completes on an mbarrier. One thread issues the box of the CTA's NEXT tile as soon as the current stage has been consumed, so the
// load runs under the DCT / quantisation / IDCT / colour work of the current tile (generation 3 relies on 8 co-resident single-tile CTAs
// for that overlap and spends instructions on issuing 16-byte cp.async copies).
// The dense stage cannot be read block-per-lane without bank conflicts (block pitch 96 B: only even 16-byte slots), so a colour
// pre-pass converts it to the planar level-shifted Y | Cb | Cr layout of generation 3 (block pitch 196 words): a quarter-warp takes
// 4 blocks x 2 half-rows with the row staggered per lane (i0 below) so that BOTH its 48-byte dense reads (slots 3u mod 8) and its planar
// stores (slots block + 2*row + half mod 8) are conflict-free. Arithmetic and its order are those of generation 3: bit-identical output.
constexpr int kDenseFloats = kTileBlocks * 192;
constexpr uint32_t kDenseBytes = kDenseFloats * sizeof(float);
struct Fwd4Geom {
    int tbx, log_tbx, tby;      // blocks per tile along x (power of two, 4..32) and y (32 / tbx)
    int tiles_x, ntiles;        // tiles per image row of blocks, total
    int nbr, nbh, nbw, rowf;    // block rows of the whole tensor (N * H / 8), per image, blocks per row, floats per pixel row
    int cx_mul;                 // tensor-map coordinate of a tile along x = tile_x * cx_mul (K neighbouring blocks are one tensor-map element row, see ni_djpeg_fwd)
};

// CTA = four warps: warps 0-2 own the Y | Cb | Cr channels of the tile's 32 blocks in the DCT phase (warp = channel, lane = block);
// warp 3 is the producer (its lane 0 issues the next tile's box as soon as the stage has been consumed: an issue costs the issuing
// thread several hundred cycles, tools/hw_probes.py, and now falls into the DCT phase in which warp 3 has nothing else to do). All FOUR
// warps share the two latency-bound passes — colour pre-pass (16 steps = 4 per warp) and colour output (512 four-pixel units = 4 per
// thread) — which shortens them by a quarter to a third compared with three warps doing 6/5/5 steps.
// Measured and dropped (profiles/r2_djpeg4_variants.json): 2 stages (3 CTAs per SM fit: slower), two or three compute groups sharing one
// stage and one producer (18 compute warps per SM, but a stage's load + pre-pass serialise: 0.16 - 0.21 ms against 0.12).
constexpr int kThreads4 = 128;

template <int MODE, bool WRITE_X>
__global__ void __launch_bounds__(kThreads4, 4)
djpeg_fwd4_kernel(const __grid_constant__ CUtensorMap tmX, float* __restrict__ y, float* __restrict__ Xd, const Fwd4Geom g,
                  const __grid_constant__ DjpegTables tab) {
    extern __shared__ __align__(16) float smem[];
    // 128-byte aligned stage base by OFFSET arithmetic (a uintptr_t round trip loses the shared address space: generic LD.E / ST.E)
    float* stage = smem + (((128u - (tc::smem_u32(smem) & 127u)) & 127u) >> 2);
    float* planar = stage + kDenseFloats;
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(planar + kTileFloats);
    const int c = threadIdx.x >> 5, lane = threadIdx.x & 31;

    auto issue = [&](int tx_, int ty_) {
        tc::mbar_expect_tx(bar_full, kDenseBytes);
        tc::tma_load_3d(stage, &tmX, bar_full, 0, tx_ * g.cx_mul, ty_ * g.tby * 8);
    };
    int tile = blockIdx.x;
    int tx = tile % g.tiles_x, ty = tile / g.tiles_x;
    const int step_x = gridDim.x % g.tiles_x, step_y = gridDim.x / g.tiles_x;
    if (threadIdx.x == 96) {
        tc::mbar_init(bar_full, 1);
        tc::fence_barrier_init();
        tc::tma_prefetch_desc(&tmX);
        if (tile < g.ntiles) issue(tx, ty);
    }
    __syncthreads();

    // this thread's block in the DCT phase
    const int bxl = lane & (g.tbx - 1), byl = lane >> g.log_tbx;
    float* bp = planar + lane * kBlockFloats + c * 64;
    // Index arithmetic is tile-invariant: computed once per thread, in 32 bits.
    // pre-pass role: lane = (block p_bl of a 16-block half tile, half row); step t = c + 4 tt covers half tile tt >> 1 at row
    // (i0 + t) & 7 with the start row i0 staggered per lane (conflict-free dense reads AND planar stores, see above)
    const int p_half = lane & 1, p_bl = lane >> 1;
    const int p_i0 = (p_half ? ((p_bl & 1) ? 1 : 2) : 0) + c;
    const int row_f = 24 << g.log_tbx;                                   // floats per dense pixel row
    int p_src[2], p_dst[2];
#pragma unroll
    for (int h2 = 0; h2 < 2; ++h2) {
        const int b = h2 * 16 + p_bl;
        p_src[h2] = (((b >> g.log_tbx) * 8) << g.log_tbx) * 24 + (b & (g.tbx - 1)) * 24 + p_half * 12;
        p_dst[h2] = b * kBlockFloats + p_half * 4;
    }
    // colour-pass role: unit u = tid + 128 it -> block (tid >> 4) + 8 it, row / half (tid & 15) fixed per thread
    const int o_r = threadIdx.x & 15, o_b0 = threadIdx.x >> 4;
    const float* o_pl = planar + o_b0 * kBlockFloats + o_r * 4;
    const int o_thr = (o_r >> 1) * g.rowf + (o_r & 1) * 12;              // float offset of the thread's 4 pixels inside its block
    const int brow_f = 8 * g.rowf;                                       // floats per block row of the image tensor

    for (int k = 0; tile < g.ntiles; tile += gridDim.x, ++k) {
        const uint32_t ph = (uint32_t)(k & 1);
        if (lane == 0) tc::mbar_wait(bar_full, ph, 20);
        __syncwarp();
        tc::mbar_wait(bar_full, ph, 21);              // every lane observes the completed phase itself (acquire), first try succeeds

        // ---- colour pre-pass: dense interleaved RGB -> planar level-shifted Y | Cb | Cr
#pragma unroll
        for (int tt = 0; tt < 4; ++tt) {
            const int i = (p_i0 + 4 * tt) & 7;
            const float4* src = reinterpret_cast<const float4*>(stage + p_src[tt >> 1] + i * row_f);
            const float4 q0 = src[0], q1 = src[1], q2 = src[2];
            const float f[12] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w};
            float4* dst = reinterpret_cast<float4*>(planar + p_dst[tt >> 1] + i * 8);
#pragma unroll
            for (int cc = 0; cc < 3; ++cc) {
                float k0, kr, kg, kb;
                color_fwd_coeffs(cc, k0, kr, kg, kb);
                float o[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) o[j] = fmaf(kb, f[3 * j + 2], fmaf(kg, f[3 * j + 1], fmaf(kr, f[3 * j], k0)));
                dst[cc * 16] = make_float4(o[0], o[1], o[2], o[3]);
            }
        }
        __syncthreads();   // planar tile complete; every register loaded from the stage has been consumed by the stores above

        const int rows_valid = g.nbr - ty * g.tby;      // block rows of this tile inside the tensor (>= tby except in the last tile row)
        int ntx = tx + step_x, nty = ty + step_y;       // the CTA's next tile
        if (ntx >= g.tiles_x) { ntx -= g.tiles_x; ++nty; }
        if (c == 3) {
            // ---- producer: the next tile's box goes into the stage while the other three warps transform this tile
            if (lane == 0 && tile + (int)gridDim.x < g.ntiles) {
                tc::fence_proxy_async_smem();
                issue(ntx, nty);
            }
        } else {
            // ---- per (block, channel): DCT -> quantisation -> IDCT in registers
            float v[8][8];
            load_plane(bp, v);
            dct2d_fwd(v);
            if (c == 0) quantise_block<MODE, 0>(v, tab); else quantise_block<MODE, 1>(v, tab);
            if (WRITE_X) {
                if (byl < rows_valid) {
                    const int gbr = ty * g.tby + byl;   // block row of this thread's block in the (N * H / 8) x nbw block grid
                    const long long n = gbr / g.nbh, by = gbr % g.nbh, nb = (long long)g.nbw * g.nbh;
                    float4* o = reinterpret_cast<float4*>(Xd + ((n * 3 + c) * nb + by * g.nbw + (tx << g.log_tbx) + bxl) * 64);
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk) {
                        o[kk * 2 + 0] = make_float4(v[kk][0], v[kk][1], v[kk][2], v[kk][3]);
                        o[kk * 2 + 1] = make_float4(v[kk][4], v[kk][5], v[kk][6], v[kk][7]);
                    }
                }
            }
            dct2d_inv(v);
            store_plane(bp, v);
        }
        __syncthreads();

        // ---- colour pass: unit u = (block u/16, row (u%16)/2, half u%2) = 4 pixels = 48 contiguous output bytes; all planar loads of a
        // thread's four units first, each into registers of its own (a load into a register that an earlier global store still has to read
        // waits on that store)
        float* const y_tile = y + ((long long)ty * g.tby * brow_f + (long long)(tx << g.log_tbx) * 24 + o_thr);
        float4 pl[4][3];
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const float4* pp = reinterpret_cast<const float4*>(o_pl + it * 8 * kBlockFloats);
            pl[it][0] = pp[0]; pl[it][1] = pp[16]; pl[it][2] = pp[32];
        }
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const int bb = o_b0 + it * 8;
            const int brl = bb >> g.log_tbx;
            if (brl < rows_valid) {
                float f[12];
                color_inv_clip4(pl[it][0], pl[it][1], pl[it][2], f);
                float4* o = reinterpret_cast<float4*>(y_tile + (brl * brow_f + (bb & (g.tbx - 1)) * 24));
                __stcs(o, make_float4(f[0], f[1], f[2], f[3]));
                __stcs(o + 1, make_float4(f[4], f[5], f[6], f[7]));
                __stcs(o + 2, make_float4(f[8], f[9], f[10], f[11]));
            }
        }
        __syncthreads();   // the planar tile is rewritten by the next pre-pass
        tx = ntx; ty = nty;
    }
}

// ---------------------------------------------------------------------------------------------------- backward, generation 3
// dx = J^T dy with everything recomputed from x. Chain (per block-channel):
//   r = C_F[1,255x]-127 ; Z = (F r F^T)/Q ; Zq = q(Z) ; xi = F^T (Zq*Q) F ; ypre = (C_I[1,xi+127])/255 ; y = clip(ypre)
//   g_ypre = dy * 1[0<=ypre<=1] ; g_xi = C_I[:,1:]^T g_ypre / 255 ; G = F g_xi F^T ; g_Z-path: G*Q*q'(Z)/Q = G*q'(Z)
//   g_r = F^T (G q') F ; dx = 255 * C_F[:,1:]^T g_r
// Same mapping as djpeg_fwd3_kernel (warp = channel, lane = block, tables from the constant bank, planar exchange through the
// x tile). q'(Z) uses the period-1 identity cos(2 pi z) = cos(2 pi (z - rint(z))): the reduction is exact in float32, so the
// special-function unit is accurate (|err| < 5e-7) whatever |z| is, and the ~40-instruction cosf() slow path disappears.
template <int MODE>
__device__ __forceinline__ float quant_grad_fast(float z) {
    const float r = z - ni_round_he(z);                   // exact, in [-0.5, 0.5]
    const float cs = __cosf(kTwoPi * r);
    return MODE == 2 ? fmaf(-2.f, cs, 1.f) : 1.f - cs;
}
template <int MODE, int CC>
__device__ __forceinline__ void quantise_block_grad(float (&v)[8][8], float (&qg)[8][8], const DjpegTables& tab) {
#pragma unroll
    for (int k = 0; k < 8; ++k)
#pragma unroll
        for (int l = 0; l < 8; ++l) {
            const float z = v[k][l] * tab.rq[CC][k * 8 + l];
            qg[k][l] = quant_grad_fast<MODE>(z);
            v[k][l] = quant_fwd<MODE>(z) * tab.q[CC][k * 8 + l];
        }
}
template <int MODE>
__global__ void __launch_bounds__(kThreads, 4)
djpeg_bwd3_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx, int H, int W, int nblk,
                  const __grid_constant__ DjpegTables tab) {
    extern __shared__ __align__(16) float smem[];
    float* tx = smem;                 // x tile (interleaved) -> xi -> g_xi -> g_r (planar)
    float* tg = smem + kTileFloats;   // dy tile (interleaved)
    long long* base = reinterpret_cast<long long*>(smem + 2 * kTileFloats);
    const int nbw = W / 8, nbh = H / 8, rowf = W * 3;
    const int g0 = blockIdx.x * kTileBlocks;
    const int c = threadIdx.x >> 5, blk = threadIdx.x & 31;

    if (c == 0) {
        const int g = g0 + blk;
        base[blk] = g < nblk ? block_base_offset(g, nbw, nbh, W) : -1;
    }
    __syncthreads();
    tile_load(tx, x, base, rowf);
    tile_load(tg, dy, base, rowf);
    cp_async_wait_all();
    __syncthreads();

    float* bp = tx + blk * kBlockFloats;
    float v[8][8];
    float qg[8][8];
    load_channel(bp, c, v);
    __syncthreads();
    dct2d_fwd(v);
    if (c == 0) quantise_block_grad<MODE, 0>(v, qg, tab); else quantise_block_grad<MODE, 1>(v, qg, tab);
    dct2d_inv(v);
    store_plane(bp + c * 64, v);
    __syncthreads();

    // pixel pass 1: clip mask from the recomputed ypre, g_xi = C_I[:,1:]^T (mask * dy) / 255, planar, in place over xi
    for (int u = threadIdx.x; u < kTileBlocks * kUnitsPerBlock; u += kThreads) {
        const int b = u >> 4, r = u & 15;
        float4* p = reinterpret_cast<float4*>(tx + b * kBlockFloats + r * 4);
        const float4* pg = reinterpret_cast<const float4*>(tg + b * kBlockFloats + (r >> 1) * 24 + (r & 1) * 12);
        const float4 Y = p[0], B = p[16], R = p[32];
        const float4 ga = pg[0], gb4 = pg[1], gc = pg[2];
        const float yy[4] = {Y.x, Y.y, Y.z, Y.w}, cb[4] = {B.x, B.y, B.z, B.w}, cr[4] = {R.x, R.y, R.z, R.w};
        const float g[12] = {ga.x, ga.y, ga.z, ga.w, gb4.x, gb4.y, gb4.z, gb4.w, gc.x, gc.y, gc.z, gc.w};
        constexpr float s = 1.f / 255.f;
        constexpr float kr = (127.f - 1.402f) * s, kg = (127.f + 1.058272f) * s, kb = (127.f - 1.772f) * s;
        float oy[4], ob[4], orr[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float pr = fmaf(cr[j], 1.402f * s, fmaf(yy[j], s, kr));
            const float pgn = fmaf(cr[j], -0.714136f * s, fmaf(cb[j], -0.344136f * s, fmaf(yy[j], s, kg)));
            const float pb = fmaf(cb[j], 1.772f * s, fmaf(yy[j], s, kb));
            const float gr = (pr >= 0.f && pr <= 1.f) ? g[3 * j] * s : 0.f;
            const float ggn = (pgn >= 0.f && pgn <= 1.f) ? g[3 * j + 1] * s : 0.f;
            const float gbl = (pb >= 0.f && pb <= 1.f) ? g[3 * j + 2] * s : 0.f;
            oy[j] = gr + ggn + gbl;                               // d/dY
            ob[j] = fmaf(-0.344136f, ggn, 1.772f * gbl);          // d/dCb
            orr[j] = fmaf(1.402f, gr, -0.714136f * ggn);          // d/dCr
        }
        p[0] = make_float4(oy[0], oy[1], oy[2], oy[3]);
        p[16] = make_float4(ob[0], ob[1], ob[2], ob[3]);
        p[32] = make_float4(orr[0], orr[1], orr[2], orr[3]);
    }
    __syncthreads();

    load_plane(bp + c * 64, v);
    dct2d_fwd(v);
#pragma unroll
    for (int k = 0; k < 8; ++k)
#pragma unroll
        for (int l = 0; l < 8; ++l) v[k][l] *= qg[k][l];
    dct2d_inv(v);
    store_plane(bp + c * 64, v);   // same thread, same addresses as the load above
    __syncthreads();

    // pixel pass 2: dx = 255 * C_F[:,1:]^T g_r, written straight to global memory
    for (int u = threadIdx.x; u < kTileBlocks * kUnitsPerBlock; u += kThreads) {
        const int b = u >> 4, r = u & 15;
        const long long bo = base[b];
        if (bo < 0) continue;
        const float4* p = reinterpret_cast<const float4*>(tx + b * kBlockFloats + r * 4);
        const float4 Y = p[0], B = p[16], R = p[32];
        const float gy[4] = {Y.x, Y.y, Y.z, Y.w}, gb[4] = {B.x, B.y, B.z, B.w}, gr[4] = {R.x, R.y, R.z, R.w};
        float f[12];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            f[3 * j] = fmaf(255.f * 0.299f, gy[j], fmaf(255.f * -0.168736f, gb[j], (255.f * 0.5f) * gr[j]));
            f[3 * j + 1] = fmaf(255.f * 0.587f, gy[j], fmaf(255.f * -0.331264f, gb[j], (255.f * -0.418688f) * gr[j]));
            f[3 * j + 2] = fmaf(255.f * 0.114f, gy[j], fmaf(255.f * 0.5f, gb[j], (255.f * -0.081312f) * gr[j]));
        }
        float4* o = reinterpret_cast<float4*>(dx + bo + (long long)(r >> 1) * rowf + (r & 1) * 12);
        __stcs(o, make_float4(f[0], f[1], f[2], f[3]));
        __stcs(o + 1, make_float4(f[4], f[5], f[6], f[7]));
        __stcs(o + 2, make_float4(f[8], f[9], f[10], f[11]));
    }
}

// ---------------------------------------------------------------------------------------------------- table gradient (trainable Q)
// DifferentiableJPEG(trainable=True) (models/jpeg.py:58-62): the two 8x8 quantisation tables are model weights. With Z = D / Q,
// X = q(Z) * Q:  dX/dQ = q(Z) - Z q'(Z)  ('soft': forward value round(Z), gradient 1 - cos 2 pi Z), so
//   dL/dQ[k][l] = sum over the blocks that use the table of G[k][l] * (q(Z) - Z q'(Z)),   G = F g_xi F^T = dL/dX.
// Same tile / thread mapping as the backward kernel (warp = channel, lane = block); the per-block terms are summed over the warp's
// 32 blocks with a butterfly (lane l ends with entries l and 32 + l) and added to the 2 x 64 output with atomics. Not a hot path.
template <int MODE>
__global__ void __launch_bounds__(kThreads, 2)
djpeg_dq_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dq, int H, int W, int nblk,
                const __grid_constant__ DjpegTables tab) {
    extern __shared__ __align__(16) float smem[];
    float* tx = smem;
    float* tg = smem + kTileFloats;
    long long* base = reinterpret_cast<long long*>(smem + 2 * kTileFloats);
    const int nbw = W / 8, nbh = H / 8, rowf = W * 3;
    const int g0 = blockIdx.x * kTileBlocks;
    const int c = threadIdx.x >> 5, blk = threadIdx.x & 31;
    if (c == 0) {
        const int g = g0 + blk;
        base[blk] = g < nblk ? block_base_offset(g, nbw, nbh, W) : -1;
    }
    for (int i = threadIdx.x; i < 2 * kTileFloats; i += kThreads) smem[i] = 0.f;     // blocks past the end contribute exactly zero
    __syncthreads();
    tile_load(tx, x, base, rowf);
    tile_load(tg, dy, base, rowf);
    cp_async_wait_all();
    __syncthreads();

    float* bp = tx + blk * kBlockFloats;
    float v[8][8];
    float coef[8][8];
    load_channel(bp, c, v);
    __syncthreads();
    dct2d_fwd(v);
    const int cc = c == 0 ? 0 : 1;
#pragma unroll
    for (int k = 0; k < 8; ++k)
#pragma unroll
        for (int l = 0; l < 8; ++l) {
            const float z = v[k][l] * tab.rq[cc][k * 8 + l];
            const float qf = quant_fwd<MODE>(z);
            coef[k][l] = qf - z * quant_grad_fast<MODE>(z);
            v[k][l] = qf * tab.q[cc][k * 8 + l];
        }
    dct2d_inv(v);
    store_plane(bp + c * 64, v);
    __syncthreads();
    // clip mask from the recomputed ypre, g_xi = C_I[:,1:]^T (mask * dy) / 255 (pixel pass 1 of the backward kernel)
    for (int u = threadIdx.x; u < kTileBlocks * kUnitsPerBlock; u += kThreads) {
        const int b = u >> 4, r = u & 15;
        float4* p = reinterpret_cast<float4*>(tx + b * kBlockFloats + r * 4);
        const float4* pg = reinterpret_cast<const float4*>(tg + b * kBlockFloats + (r >> 1) * 24 + (r & 1) * 12);
        const float4 Y = p[0], B = p[16], R = p[32];
        const float4 ga = pg[0], gb4 = pg[1], gc = pg[2];
        const float yy[4] = {Y.x, Y.y, Y.z, Y.w}, cb[4] = {B.x, B.y, B.z, B.w}, cr[4] = {R.x, R.y, R.z, R.w};
        const float g[12] = {ga.x, ga.y, ga.z, ga.w, gb4.x, gb4.y, gb4.z, gb4.w, gc.x, gc.y, gc.z, gc.w};
        constexpr float s = 1.f / 255.f;
        constexpr float kr = (127.f - 1.402f) * s, kg = (127.f + 1.058272f) * s, kb = (127.f - 1.772f) * s;
        float oy[4], ob[4], orr[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float pr = fmaf(cr[j], 1.402f * s, fmaf(yy[j], s, kr));
            const float pgn = fmaf(cr[j], -0.714136f * s, fmaf(cb[j], -0.344136f * s, fmaf(yy[j], s, kg)));
            const float pb = fmaf(cb[j], 1.772f * s, fmaf(yy[j], s, kb));
            const float gr = (pr >= 0.f && pr <= 1.f) ? g[3 * j] * s : 0.f;
            const float ggn = (pgn >= 0.f && pgn <= 1.f) ? g[3 * j + 1] * s : 0.f;
            const float gbl = (pb >= 0.f && pb <= 1.f) ? g[3 * j + 2] * s : 0.f;
            oy[j] = gr + ggn + gbl;
            ob[j] = fmaf(-0.344136f, ggn, 1.772f * gbl);
            orr[j] = fmaf(1.402f, gr, -0.714136f * ggn);
        }
        p[0] = make_float4(oy[0], oy[1], oy[2], oy[3]);
        p[16] = make_float4(ob[0], ob[1], ob[2], ob[3]);
        p[32] = make_float4(orr[0], orr[1], orr[2], orr[3]);
    }
    __syncthreads();
    load_plane(bp + c * 64, v);
    dct2d_fwd(v);                           // G = dL/dX
    const bool live = base[blk] >= 0;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        float r[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const int e = half * 32 + j;
            r[j] = live ? v[e >> 3][e & 7] * coef[e >> 3][e & 7] : 0.f;
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
            const bool upper = (blk & off) != 0;
#pragma unroll
            for (int j = 0; j < off; ++j) {
                const float send = upper ? r[j] : r[j + off];
                const float keep = upper ? r[j + off] : r[j];
                r[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
            }
        }
        atomicAdd(dq + cc * 64 + half * 32 + blk, r[0]);
    }
}

int fill_tables(DjpegTables& t, const float* q_luma, const float* q_chroma) {
    for (int i = 0; i < 64; ++i) {
        if (!(q_luma[i] > 0.f) || !(q_chroma[i] > 0.f)) return -1;
        t.q[0][i] = q_luma[i]; t.rq[0][i] = 1.f / q_luma[i];
        t.q[1][i] = q_chroma[i]; t.rq[1][i] = 1.f / q_chroma[i];
    }
    return 0;
}

constexpr size_t kBwd3Smem = 2 * kTileFloats * sizeof(float) + kTileBlocks * sizeof(long long);
constexpr size_t kFwd3Smem = kTileFloats * sizeof(float) + kTileBlocks * sizeof(long long);
template <typename K>
int set_smem(K kernel, size_t bytes) {
    NI_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return NI_OK;
}

}  // namespace

extern "C" int ni_djpeg_fwd(const float* x, float* y, float* x_deq, int n, int h, int w, const float* q_luma,
                            const float* q_chroma, int mode, cudaStream_t stream) {
    if (n == 0) return NI_OK;
    NI_REQUIRE(x && y && q_luma && q_chroma, "ni_djpeg_fwd: null pointer");
    NI_REQUIRE(n >= 0 && h > 0 && w > 0 && h % 8 == 0 && w % 8 == 0,
               "ni_djpeg_fwd: H and W must be positive multiples of 8 (got %d x %d)", h, w);
    NI_REQUIRE(mode >= 0 && mode <= 2, "ni_djpeg_fwd: mode must be 0 (soft), 1 (sin) or 2 (harmonic), got %d", mode);
    if (n == 0) return NI_OK;
    DjpegTables tab;
    NI_REQUIRE(fill_tables(tab, q_luma, q_chroma) == 0, "ni_djpeg_fwd: quantisation tables must be positive");
    const long long nblk = (long long)n * (h / 8) * (w / 8);
    NI_REQUIRE(nblk < (1ll << 31) - 64 * kTileBlocks, "ni_djpeg_fwd: tensor too large (%lld blocks)", nblk);
    // generation 4 (persistent CTAs, TMA ring) where the block grid tiles into rectangular 32-block boxes; generation 3 otherwise
    const int nbw = w / 8;
    int variant = 4, ctas = 4;
#ifdef NI_DEV
    if (const char* e = getenv("NI_DJPEG_FWD")) variant = e[0] - '0';       // 3 | 4
    if (const char* e = getenv("NI_DJPEG_CTAS")) ctas = atoi(e) > 0 ? atoi(e) : 4;
#endif
    if (variant == 4 && nbw % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (long long)n * h < (1ll << 31)) {
        Fwd4Geom g;
        g.tbx = 4; g.log_tbx = 2;
        while (g.tbx < 32 && nbw % (g.tbx * 2) == 0) { g.tbx *= 2; ++g.log_tbx; }
        g.tby = kTileBlocks / g.tbx;
        g.nbw = nbw; g.nbh = h / 8; g.nbr = n * g.nbh; g.rowf = w * 3;
        g.tiles_x = nbw / g.tbx;
        const long long nt = (long long)g.tiles_x * ((g.nbr + g.tby - 1) / g.tby);
        NI_REQUIRE(nt < (1ll << 31), "ni_djpeg_fwd: tensor too large (%lld tiles)", nt);
        g.ntiles = (int)nt;
        // The copy engine works row by row (a few cycles per box row whatever its length, tools/hw_probes.py), and the 24 floats of one
        // block row are only 96 bytes: K = 8 (or 4) neighbouring blocks of an image row are contiguous in memory, so the innermost
        // tensor dimension is 24 K floats (768 bytes) and a tile is 8 * tby * tbx / K box rows instead of 256. Same bytes, same dense
        // shared-memory image ([pixel row][block][24]).
        const int kmerge = g.tbx % 8 == 0 ? 8 : 4;
        g.cx_mul = g.tbx / kmerge;
        CUtensorMap tm;
        const cuuint64_t dims[3] = {(cuuint64_t)24 * kmerge, (cuuint64_t)(nbw / kmerge), (cuuint64_t)n * h};
        const cuuint64_t strides[2] = {(cuuint64_t)96 * kmerge, (cuuint64_t)w * 12};
        const cuuint32_t box[3] = {(cuuint32_t)24 * kmerge, (cuuint32_t)(g.tbx / kmerge), (cuuint32_t)g.tby * 8};
        int rc = ni_encode_tiled_sw(&tm, x, 3, dims, strides, box, 0);
        if (rc) return rc;
        const size_t smem_bytes = 128 + kDenseBytes + kTileFloats * sizeof(float) + 2 * sizeof(uint64_t);
        const int grid4 = (int)(nt < (long long)ctas * ni_num_sms() ? nt : (long long)ctas * ni_num_sms());
#define NI_FWD4(MODE, WX)                                                                                        \
    {                                                                                                             \
        rc = set_smem(djpeg_fwd4_kernel<MODE, WX>, smem_bytes);                                                   \
        if (rc) return rc;                                                                                        \
        djpeg_fwd4_kernel<MODE, WX><<<grid4, kThreads4, smem_bytes, stream>>>(tm, y, x_deq, g, tab);              \
    }
        if (x_deq) {
            if (mode == 0) NI_FWD4(0, true) else if (mode == 1) NI_FWD4(1, true) else NI_FWD4(2, true)
        } else {
            if (mode == 0) NI_FWD4(0, false) else if (mode == 1) NI_FWD4(1, false) else NI_FWD4(2, false)
        }
#undef NI_FWD4
        NI_LAUNCH_CHECK();
        NI_COUNT_LAUNCH(1);
        return NI_OK;
    }
    const int grid = ni_cdiv(nblk, kTileBlocks);
#define NI_FWD(MODE, WX)                                                                                      \
    {                                                                                                          \
        int rc = set_smem(djpeg_fwd3_kernel<MODE, WX>, kFwd3Smem);                                             \
        if (rc) return rc;                                                                                     \
        djpeg_fwd3_kernel<MODE, WX><<<grid, kThreads, kFwd3Smem, stream>>>(x, y, x_deq, h, w, (int)nblk, tab); \
    }
    if (x_deq) {
        if (mode == 0) NI_FWD(0, true) else if (mode == 1) NI_FWD(1, true) else NI_FWD(2, true)
    } else {
        if (mode == 0) NI_FWD(0, false) else if (mode == 1) NI_FWD(1, false) else NI_FWD(2, false)
    }
#undef NI_FWD
    NI_LAUNCH_CHECK();
    NI_COUNT_LAUNCH(1);
    return NI_OK;
}

extern "C" int ni_djpeg_bwd(const float* x, const float* dy, float* dx, int n, int h, int w, const float* q_luma,
                            const float* q_chroma, int mode, cudaStream_t stream) {
    if (n == 0) return NI_OK;
    NI_REQUIRE(x && dy && dx && q_luma && q_chroma, "ni_djpeg_bwd: null pointer");
    NI_REQUIRE(n >= 0 && h > 0 && w > 0 && h % 8 == 0 && w % 8 == 0,
               "ni_djpeg_bwd: H and W must be positive multiples of 8 (got %d x %d)", h, w);
    NI_REQUIRE(mode >= 0 && mode <= 2, "ni_djpeg_bwd: mode must be 0 (soft), 1 (sin) or 2 (harmonic), got %d", mode);
    if (n == 0) return NI_OK;
    DjpegTables tab;
    NI_REQUIRE(fill_tables(tab, q_luma, q_chroma) == 0, "ni_djpeg_bwd: quantisation tables must be positive");
    const long long nblk = (long long)n * (h / 8) * (w / 8);
    NI_REQUIRE(nblk < (1ll << 31) - 64 * kTileBlocks, "ni_djpeg_bwd: tensor too large (%lld blocks)", nblk);
    const int grid = ni_cdiv(nblk, kTileBlocks);
#define NI_BWD(MODE)                                                                                  \
    {                                                                                                  \
        int rc = set_smem(djpeg_bwd3_kernel<MODE>, kBwd3Smem);                                         \
        if (rc) return rc;                                                                             \
        djpeg_bwd3_kernel<MODE><<<grid, kThreads, kBwd3Smem, stream>>>(x, dy, dx, h, w, (int)nblk, tab); \
    }
    if (mode == 0) NI_BWD(0) else if (mode == 1) NI_BWD(1) else NI_BWD(2)
#undef NI_BWD
    NI_LAUNCH_CHECK();
    NI_COUNT_LAUNCH(1);
    return NI_OK;
}

// dq: 2 x 64 floats (luma table, chroma table), overwritten. Gradient of the loss w.r.t. the quantisation tables of a trainable
// DifferentiableJPEG (models/jpeg.py:58-62) given dy = dL/dy; everything is recomputed from x like ni_djpeg_bwd.
extern "C" int ni_djpeg_bwd_tables(const float* x, const float* dy, float* dq, int n, int h, int w, const float* q_luma,
                                   const float* q_chroma, int mode, cudaStream_t stream) {
    NI_REQUIRE(x && dy && dq && q_luma && q_chroma, "ni_djpeg_bwd_tables: null pointer");
    NI_REQUIRE(n >= 0 && h > 0 && w > 0 && h % 8 == 0 && w % 8 == 0,
               "ni_djpeg_bwd_tables: H and W must be positive multiples of 8 (got %d x %d)", h, w);
    NI_REQUIRE(mode >= 0 && mode <= 2, "ni_djpeg_bwd_tables: mode must be 0 (soft), 1 (sin) or 2 (harmonic), got %d", mode);
    NI_CUDA(cudaMemsetAsync(dq, 0, sizeof(float) * 128, stream));
    if (n == 0) return NI_OK;
    DjpegTables tab;
    NI_REQUIRE(fill_tables(tab, q_luma, q_chroma) == 0, "ni_djpeg_bwd_tables: quantisation tables must be positive");
    const long long nblk = (long long)n * (h / 8) * (w / 8);
    NI_REQUIRE(nblk < (1ll << 31) - 64 * kTileBlocks, "ni_djpeg_bwd_tables: tensor too large (%lld blocks)", nblk);
    const int grid = ni_cdiv(nblk, kTileBlocks);
#define NI_DQ(MODE)                                                                                  \
    {                                                                                                  \
        int rc = set_smem(djpeg_dq_kernel<MODE>, kBwd3Smem);                                           \
        if (rc) return rc;                                                                             \
        djpeg_dq_kernel<MODE><<<grid, kThreads, kBwd3Smem, stream>>>(x, dy, dq, h, w, (int)nblk, tab); \
    }
    if (mode == 0) NI_DQ(0) else if (mode == 1) NI_DQ(1) else NI_DQ(2)
#undef NI_DQ
    NI_LAUNCH_CHECK();
    NI_COUNT_LAUNCH(1);
    return NI_OK;
}
