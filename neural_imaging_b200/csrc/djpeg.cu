// Fused differentiable-JPEG kernels (forward and backward) for sm_100a.
//
// Replaces the whole of DifferentiableJPEG.call (reference models/jpeg.py:91-159): colour transform ->
// 8x8 blocking -> DCT (literal 4-decimal matrix, jpeg.py:78-85) -> /Q -> Quantization (models/layers.py:118-128)
// -> *Q -> IDCT -> inverse blocking -> colour transform -> /255 -> clip, in ONE kernel (the reference materialises
// >= 25 full-size temporaries). HBM traffic = read x + write y (24 B/pixel) forward; read x, dy + write dx backward
// (36 B/pixel, X/Q is recomputed instead of saved).
//
// Work decomposition: a CTA owns a tile of 64 8x8 pixel blocks (48 KB of interleaved RGB staged in shared
// memory); one thread owns one (block, Y/Cb/Cr channel) pair and keeps the whole 8x8 coefficient block in
// registers, so the 2-D DCT / quantisation / IDCT need no inter-thread exchange. The 1-D transforms use the
// even/odd symmetry that the literal rows preserve (36 instead of 64 FMA-class ops per 8-point transform) with
// the literal coefficients as FFMA immediates. The colour transforms (which mix channels) run in cooperative
// per-pixel passes over the shared-memory tile.
#include "ni_common.cuh"

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

template <int MODE>
__device__ __forceinline__ float quant_fwd(float z) {
    if (MODE == 0) return ni_round_he(z);
    if (MODE == 1) return z - sinf(kTwoPi * z) / kTwoPi;
    return z - sinf(kTwoPi * z) / kPi;
}
template <int MODE>
__device__ __forceinline__ float quant_grad(float z) {
    if (MODE == 2) return 1.f - 2.f * cosf(kTwoPi * z);
    return 1.f - cosf(kTwoPi * z);  // soft: backward of the sin approximation; sin: its own derivative
}

// Colour constants. Forward rows of _color_F with the 255 input scale and the -127 level shift folded in; inverse
// rows of _color_I with the +127 shift and the /255 folded in (reference models/jpeg.py:74-75,99-105,154-156).
__device__ __forceinline__ void color_fwd_coeffs(int c, float& k0, float& kr, float& kg, float& kb) {
    if (c == 0) { k0 = 0.f - 127.f; kr = 255.f * 0.299f; kg = 255.f * 0.587f; kb = 255.f * 0.114f; }
    else if (c == 1) { k0 = 128.f - 127.f; kr = 255.f * -0.168736f; kg = 255.f * -0.331264f; kb = 255.f * 0.5f; }
    else { k0 = 128.f - 127.f; kr = 255.f * 0.5f; kg = 255.f * -0.418688f; kb = 255.f * -0.081312f; }
}

// ypre (before clip) for one pixel from level-shifted Y, Cb, Cr (xi, i.e. without the +127).
__device__ __forceinline__ void color_inv(float yy, float cb, float cr, float& r, float& g, float& b) {
    const float Y = yy + 127.f, B = cb + 127.f, R = cr + 127.f;
    const float s = 1.f / 255.f;
    r = (fmaf(1.402f, R, Y) + (-1.402f * 128.f)) * s;
    g = (fmaf(-0.714136f, R, fmaf(-0.344136f, B, Y)) + (1.058272f * 128.f)) * s;
    b = (fmaf(1.772f, B, Y) + (-1.772f * 128.f)) * s;
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const unsigned sa = static_cast<unsigned>(__cvta_generic_to_shared(smem));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// Float offset of (block g, row 0, col 0) in an (N,H,W,3) tensor; computed once per block per tile.
__device__ __forceinline__ long long block_base_offset(long long g, int nbw, int nbh, int W) {
    const int bx = (int)(g % nbw);
    const long long t = g / nbw;
    const int by = (int)(t % nbh);
    const long long n = t / nbh;
    return (((n * nbh + by) * 8) * (long long)W + bx * 8) * 3;
}
__device__ __forceinline__ void fill_base_table(long long* base, long long g0, long long nblk, int nbw, int nbh, int W) {
    if (threadIdx.x < kTileBlocks) {
        const long long g = g0 + threadIdx.x;
        base[threadIdx.x] = g < nblk ? block_base_offset(g, nbw, nbh, W) : -1;
    }
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
// global float offset of the 4-pixel unit u (see unit_offset) or -1 for blocks past the end
__device__ __forceinline__ long long unit_global(int u, const long long* base, int rowf) {
    const long long b = base[u / kUnitsPerBlock];
    const int r = u % kUnitsPerBlock;
    return b < 0 ? -1 : b + (long long)(r >> 1) * rowf + (r & 1) * 12;
}
// address of 4-pixel unit u of the per-pixel passes
__device__ __forceinline__ int unit_offset(int u) { return (u / kUnitsPerBlock) * kBlockFloats + (u % kUnitsPerBlock) * 12; }

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
__device__ __forceinline__ void store_channel(float* bp, int c, const float (&v)[8][8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) bp[i * 24 + 3 * j + c] = v[i][j];
}
__device__ __forceinline__ void load_channel_raw(const float* bp, int c, float (&v)[8][8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) v[i][j] = bp[i * 24 + 3 * j + c];
}

__device__ __forceinline__ void load_tables(float2* sq, const DjpegTables& tab) {
    // layout: sq[idx*2 + cc] = {Q, 1/Q}; adjacent banks for the two tables -> conflict-free mixed-channel reads.
    for (int i = threadIdx.x; i < 128; i += kThreads) {
        const int cc = i & 1, idx = i >> 1;
        sq[i] = make_float2(tab.q[cc][idx], tab.rq[cc][idx]);
    }
}

// ---------------------------------------------------------------------------------------------------- forward
template <int MODE, bool WRITE_X>
__global__ void __launch_bounds__(kThreads, 6)
djpeg_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, float* __restrict__ Xd, int H, int W,
                 long long nblk, DjpegTables tab) {
    extern __shared__ __align__(16) float smem[];
    float* tile = smem;
    float2* sq = reinterpret_cast<float2*>(smem + kTileFloats);
    long long* base = reinterpret_cast<long long*>(sq + 128);
    const int nbw = W / 8, nbh = H / 8, rowf = W * 3;
    const long long g0 = (long long)blockIdx.x * kTileBlocks;

    fill_base_table(base, g0, nblk, nbw, nbh, W);
    load_tables(sq, tab);
    __syncthreads();
    tile_load(tile, x, base, rowf);
    cp_async_wait_all();
    __syncthreads();

    const int blk = threadIdx.x / 3, c = threadIdx.x % 3, cc = c ? 1 : 0;
    float* bp = tile + blk * kBlockFloats;
    float v[8][8];
    load_channel(bp, c, v);
    __syncthreads();  // all three channel threads of a block have read the RGB data before it is overwritten

    dct2d_fwd(v);
#pragma unroll
    for (int k = 0; k < 8; ++k)
#pragma unroll
        for (int l = 0; l < 8; ++l) {
            const float2 qr = sq[(k * 8 + l) * 2 + cc];
            v[k][l] = quant_fwd<MODE>(v[k][l] * qr.y) * qr.x;
        }
    if (WRITE_X) {
        // de-quantised coefficients, reference block order ((n*3+c)*nb + by*nbw + bx, k, l)  (models/jpeg.py:105-114,159)
        const long long g = g0 + blk;
        if (g < nblk) {
            const long long nb = (long long)nbw * nbh;
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
    store_channel(bp, c, v);
    __syncthreads();

    // per-pixel inverse colour transform + /255 + clip, 4 pixels (3 float4 = 48 contiguous bytes) per step, written straight
    // to global memory (saves a shared-memory round trip and a barrier; L2 merges the partial-sector writes of a warp)
    for (int u = threadIdx.x; u < kTileBlocks * kUnitsPerBlock; u += kThreads) {
        const long long go = unit_global(u, base, rowf);
        if (go < 0) continue;
        const float4* p = reinterpret_cast<const float4*>(tile + unit_offset(u));
        float4 a = p[0], b = p[1], d = p[2];
        float f[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, d.x, d.y, d.z, d.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float r, g, bl;
            color_inv(f[3 * j], f[3 * j + 1], f[3 * j + 2], r, g, bl);
            f[3 * j] = ni_clamp01(r); f[3 * j + 1] = ni_clamp01(g); f[3 * j + 2] = ni_clamp01(bl);
        }
        float4* o = reinterpret_cast<float4*>(y + go);
        __stcs(o, make_float4(f[0], f[1], f[2], f[3]));
        __stcs(o + 1, make_float4(f[4], f[5], f[6], f[7]));
        __stcs(o + 2, make_float4(f[8], f[9], f[10], f[11]));
    }
}

// ---------------------------------------------------------------------------------------------------- backward
// dx = J^T dy with everything recomputed from x. Chain (per block-channel):
//   r = C_F[1,255x]-127 ; Z = (F r F^T)/Q ; Zq = q(Z) ; xi = F^T (Zq*Q) F ; ypre = (C_I[1,xi+127])/255 ; y = clip(ypre)
//   g_ypre = dy * 1[0<=ypre<=1] ; g_xi = C_I[:,1:]^T g_ypre / 255 ; G = F g_xi F^T ; g_Z-path: G*Q*q'(Z)/Q = G*q'(Z)
//   g_r = F^T (G q') F ; dx = 255 * C_F[:,1:]^T g_r
template <int MODE>
__global__ void __launch_bounds__(kThreads, 4)
djpeg_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx, int H, int W,
                 long long nblk, DjpegTables tab) {
    extern __shared__ __align__(16) float smem[];
    float* tx = smem;                 // x tile -> xi -> g_r -> dx
    float* tg = smem + kTileFloats;   // dy tile -> g_xi
    float2* sq = reinterpret_cast<float2*>(smem + 2 * kTileFloats);
    long long* base = reinterpret_cast<long long*>(sq + 128);
    const int nbw = W / 8, nbh = H / 8, rowf = W * 3;
    const long long g0 = (long long)blockIdx.x * kTileBlocks;

    fill_base_table(base, g0, nblk, nbw, nbh, W);
    load_tables(sq, tab);
    __syncthreads();
    tile_load(tx, x, base, rowf);
    tile_load(tg, dy, base, rowf);
    cp_async_wait_all();
    __syncthreads();

    const int blk = threadIdx.x / 3, c = threadIdx.x % 3, cc = c ? 1 : 0;
    float* bx_ = tx + blk * kBlockFloats;
    float* bg_ = tg + blk * kBlockFloats;
    float v[8][8];
    float qg[8][8];
    load_channel(bx_, c, v);
    __syncthreads();
    dct2d_fwd(v);
#pragma unroll
    for (int k = 0; k < 8; ++k)
#pragma unroll
        for (int l = 0; l < 8; ++l) {
            const float2 qr = sq[(k * 8 + l) * 2 + cc];
            const float z = v[k][l] * qr.y;
            qg[k][l] = quant_grad<MODE>(z);
            v[k][l] = quant_fwd<MODE>(z) * qr.x;
        }
    dct2d_inv(v);
    store_channel(bx_, c, v);
    __syncthreads();

    // pixel pass 1: clip mask from recomputed ypre, g_xi = C_I[:,1:]^T (mask*dy)/255, written over the dy tile
    for (int u = threadIdx.x; u < kTileBlocks * kUnitsPerBlock; u += kThreads) {
        const float4* px = reinterpret_cast<const float4*>(tx + unit_offset(u));
        float4* pg = reinterpret_cast<float4*>(tg + unit_offset(u));
        float4 a = px[0], b = px[1], d = px[2];
        float f[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, d.x, d.y, d.z, d.w};
        a = pg[0]; b = pg[1]; d = pg[2];
        float g[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, d.x, d.y, d.z, d.w};
        const float s = 1.f / 255.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float r, gg, bl;
            color_inv(f[3 * j], f[3 * j + 1], f[3 * j + 2], r, gg, bl);
            const float gr = (r >= 0.f && r <= 1.f) ? g[3 * j] * s : 0.f;
            const float ggn = (gg >= 0.f && gg <= 1.f) ? g[3 * j + 1] * s : 0.f;
            const float gb = (bl >= 0.f && bl <= 1.f) ? g[3 * j + 2] * s : 0.f;
            g[3 * j] = gr + ggn + gb;                                   // d/dY
            g[3 * j + 1] = fmaf(-0.344136f, ggn, 1.772f * gb);          // d/dCb
            g[3 * j + 2] = fmaf(1.402f, gr, -0.714136f * ggn);          // d/dCr
        }
        pg[0] = make_float4(g[0], g[1], g[2], g[3]);
        pg[1] = make_float4(g[4], g[5], g[6], g[7]);
        pg[2] = make_float4(g[8], g[9], g[10], g[11]);
    }
    __syncthreads();

    load_channel_raw(bg_, c, v);
    dct2d_fwd(v);
#pragma unroll
    for (int k = 0; k < 8; ++k)
#pragma unroll
        for (int l = 0; l < 8; ++l) v[k][l] *= qg[k][l];
    dct2d_inv(v);
    store_channel(bx_, c, v);   // tx is free: xi was consumed by pixel pass 1
    __syncthreads();

    // pixel pass 2: dx = 255 * C_F[:,1:]^T g_r, written straight to global memory
    for (int u = threadIdx.x; u < kTileBlocks * kUnitsPerBlock; u += kThreads) {
        const long long go = unit_global(u, base, rowf);
        if (go < 0) continue;
        const float4* p = reinterpret_cast<const float4*>(tx + unit_offset(u));
        float4 a = p[0], b = p[1], d = p[2];
        float f[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, d.x, d.y, d.z, d.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float gy = f[3 * j], gb = f[3 * j + 1], gr = f[3 * j + 2];
            f[3 * j] = fmaf(255.f * 0.299f, gy, fmaf(255.f * -0.168736f, gb, (255.f * 0.5f) * gr));
            f[3 * j + 1] = fmaf(255.f * 0.587f, gy, fmaf(255.f * -0.331264f, gb, (255.f * -0.418688f) * gr));
            f[3 * j + 2] = fmaf(255.f * 0.114f, gy, fmaf(255.f * 0.5f, gb, (255.f * -0.081312f) * gr));
        }
        float4* o = reinterpret_cast<float4*>(dx + go);
        __stcs(o, make_float4(f[0], f[1], f[2], f[3]));
        __stcs(o + 1, make_float4(f[4], f[5], f[6], f[7]));
        __stcs(o + 2, make_float4(f[8], f[9], f[10], f[11]));
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

constexpr size_t kFwdSmem = kTileFloats * sizeof(float) + 128 * sizeof(float2) + kTileBlocks * sizeof(long long);
constexpr size_t kBwdSmem = 2 * kTileFloats * sizeof(float) + 128 * sizeof(float2) + kTileBlocks * sizeof(long long);

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
    const int grid = ni_cdiv(nblk, kTileBlocks);
#define NI_FWD(MODE, WX)                                                                                      \
    {                                                                                                          \
        int rc = set_smem(djpeg_fwd_kernel<MODE, WX>, kFwdSmem);                                               \
        if (rc) return rc;                                                                                     \
        djpeg_fwd_kernel<MODE, WX><<<grid, kThreads, kFwdSmem, stream>>>(x, y, x_deq, h, w, nblk, tab);        \
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
    const int grid = ni_cdiv(nblk, kTileBlocks);
#define NI_BWD(MODE)                                                                                  \
    {                                                                                                  \
        int rc = set_smem(djpeg_bwd_kernel<MODE>, kBwdSmem);                                           \
        if (rc) return rc;                                                                             \
        djpeg_bwd_kernel<MODE><<<grid, kThreads, kBwdSmem, stream>>>(x, dy, dx, h, w, nblk, tab);      \
    }
    if (mode == 0) NI_BWD(0) else if (mode == 1) NI_BWD(1) else NI_BWD(2)
#undef NI_BWD
    NI_LAUNCH_CHECK();
    NI_COUNT_LAUNCH(1);
    return NI_OK;
}
