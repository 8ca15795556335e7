// tcgen05 implicit-GEMM convolutions for sm_100a (fprop, dgrad, wgrad), FP32-accurate through 3xTF32.
//
// Why 3xTF32: BASELINE.json asks for 1e-5 relative parity with the reference's FP32 graphs. A single TF32 pass is
// ~3e-4. Each operand is split x = hi + lo (hi = x with the low 13 mantissa bits cleared, exactly representable in
// TF32; lo = x - hi, exact in FP32) and the product is accumulated as hi*hi + hi*lo + lo*hi in the FP32 TMEM
// accumulator (the dropped lo*lo term is ~2^-22 relative). Weights are split once per call by a prep kernel;
// activations are split tile by tile in shared memory by a splitter warp-group, so HBM/L2 traffic is unchanged.
//
// Implicit GEMM without im2col: for filter tap (a, b) the A operand tile (128 output pixels x 32 input channels) is
// ONE TMA box of the NHWC activation tensor shifted by (a - pad, b - pad); out-of-bounds box elements are
// zero-filled by the TMA unit, which IS TensorFlow's SAME zero padding. A 32-channel box row is 128 bytes = one
// SWIZZLE_128B span, so the box lands directly in the canonical K-major UMMA layout. Warp roles per CTA: warp 0 =
// TMA producer, warp 1 = MMA issuer (one elected thread) + TMEM allocator, warps 2-5 = hi/lo splitter during the
// main loop, then epilogue (tcgen05.ld -> bias + activation -> global). 3-stage mbarrier pipeline.
//
// wgrad is the same machinery with both operands MN-major: dW[(tap,ci), co] = sum_pixels X_tap[pixel, ci] dY[pixel, co];
// the pixel-major TMA boxes (32 pixels x 32 channels) are exactly the canonical MN-major SWIZZLE_128B atoms, the
// flattened (tap, ci) index fills the 128-row M tile even for 32-channel layers, split-K over pixels with atomics.
#include <cuda.h>

#include <mutex>

#include "conv_desc.h"
#include "ni_common.cuh"
#include "tc_common.cuh"

namespace {

using namespace tc;

constexpr int kStages = 3;
constexpr int kThreadsTc = 192;
constexpr int kABytes = 16384;  // 128 rows x 128 B (gemm) or 4 atoms x 32 rows x 128 B (wgrad)

__device__ __forceinline__ float apply_act_tc(float v, int act, float alpha) {
    switch (act) {
        case NI_ACT_LEAKY_RELU: return v > 0.f ? v : alpha * v;
        case NI_ACT_RELU: return fmaxf(v, 0.f);
        case NI_ACT_TANH: return tanhf(v);
        case NI_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
        case NI_ACT_CLIP01: return ni_clamp01(v);
        default: return v;
    }
}

// hi/lo split of a shared-memory region (n16 16-byte chunks), hi in place, lo to `lo`. 128 threads.
__device__ __forceinline__ void split_region(float4* hi, float4* lo, int n16, int tid) {
    for (int i = tid; i < n16; i += 128) {
        const float4 v = hi[i];
        float4 h, l;
        h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); l.x = v.x - h.x;
        h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u); l.y = v.y - h.y;
        h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); l.z = v.z - h.z;
        h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u); l.w = v.w - h.w;
        hi[i] = h;
        lo[i] = l;
    }
}

struct TcGemmParams {
    int n, oh, ow;                 // target tensor (the GEMM's M space): batch, height, width
    int bw, bh, bn;                // 128-pixel tile = bn images x bh rows x bw columns
    int tiles_w, tiles_h;          // tiles along w / h (tiles along n = gridDim.x / (tiles_w * tiles_h))
    int kh, kw, off_y0, off_x0, off_sign;   // source box origin = tile origin + off0 + sign * tap index
    int kchunks;                   // contracted channels / 32
    int ntot;                      // total N (output channels)
    int out_pitch, out_coff, out_mode;
    int bias_mod, act, accumulate;
    int nacc;                      // number of hi*hi accumulators (round-robin over k-iterations), 1..3
    float alpha;
    const float* bias;
    float* out;
};

template <int BNT>
__global__ void __launch_bounds__(kThreadsTc, 1)
conv_tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcGemmParams p) {
    constexpr int B_BYTES = BNT * 128;
    constexpr int STAGE_BYTES = 2 * kABytes + 2 * B_BYTES;
    // Two FP32 accumulators: D1 (columns [0, BNT)) takes the hi*hi products, D2 (columns [BNT, 2 BNT)) the small
    // hi*lo + lo*hi corrections. The tensor core truncates addends to the accumulator's exponent, so adding the 2^-11-sized
    // corrections into the big accumulator would throw most of their bits away; D1 + D2 is formed in FP32 in the epilogue.
    // Up to three D1 accumulators are rotated over the k-iterations for deep contractions: the truncation bias of the
    // in-TMEM accumulation grows with the accumulator magnitude (measured 1.2e-5 at K = 4608 with one accumulator).
    const uint32_t want_cols = (uint32_t)(p.nacc + 1) * BNT;
    const uint32_t TMEM_COLS = want_cols <= 32 ? 32 : (want_cols <= 64 ? 64 : (want_cols <= 128 ? 128 : (want_cols <= 256 ? 256 : 512)));
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar_full[kStages], bar_split[kStages], bar_empty[kStages], bar_accum;
    __shared__ uint32_t tmem_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tw = blockIdx.x % p.tiles_w, th = (blockIdx.x / p.tiles_w) % p.tiles_h, tn = blockIdx.x / (p.tiles_w * p.tiles_h);
    const int x0 = tw * p.bw, y0 = th * p.bh, n0 = tn * p.bn;
    const int ntile0 = blockIdx.y * BNT;
    const int taps = p.kh * p.kw;
    const int iters = taps * p.kchunks;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_split[s], 128); mbar_init(&bar_empty[s], 1); }
        mbar_init(&bar_accum, 1);
        fence_barrier_init();
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (warp == 1) tmem_alloc(&tmem_slot, TMEM_COLS);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = tmem_slot;

    auto a_hi = [&](int s) { return smem + s * STAGE_BYTES; };
    auto a_lo = [&](int s) { return smem + s * STAGE_BYTES + kABytes; };
    auto b_hi = [&](int s) { return smem + s * STAGE_BYTES + 2 * kABytes; };
    auto b_lo = [&](int s) { return smem + s * STAGE_BYTES + 2 * kABytes + B_BYTES; };

    if (warp == 0) {
        if (lane == 0) {
            for (int it = 0; it < iters; ++it) {
                const int s = it % kStages, ph = (it / kStages) & 1;
                mbar_wait(&bar_empty[s], ph ^ 1, 0);
                mbar_expect_tx(&bar_full[s], kABytes + 2 * B_BYTES);
                const int tap = it / p.kchunks, kc = it - tap * p.kchunks;
                const int a = tap / p.kw, b = tap - a * p.kw;
                tma_load_4d(a_hi(s), &tmA, &bar_full[s], kc * 32, x0 + p.off_x0 + p.off_sign * b, y0 + p.off_y0 + p.off_sign * a, n0);
                tma_load_3d(b_hi(s), &tmB, &bar_full[s], kc * 32, ntile0, tap);
                tma_load_3d(b_lo(s), &tmB, &bar_full[s], kc * 32, ntile0, taps + tap);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_tf32(128, BNT, 0, 0);
            for (int it = 0; it < iters; ++it) {
                const int s = it % kStages, ph = (it / kStages) & 1;
                mbar_wait(&bar_full[s], ph, 1);
                mbar_wait(&bar_split[s], ph, 2);
                tcgen05_fence_after();
                const uint32_t ah = smem_u32(a_hi(s)), al = smem_u32(a_lo(s)), bh = smem_u32(b_hi(s)), bl = smem_u32(b_lo(s));
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {   // 4 x K=8 (32 bytes) inside the 128-byte swizzle span
                    const uint64_t dah = make_smem_desc_sw128(ah + ks * 32, 16, 1024), dal = make_smem_desc_sw128(al + ks * 32, 16, 1024);
                    const uint64_t dbh = make_smem_desc_sw128(bh + ks * 32, 16, 1024), dbl = make_smem_desc_sw128(bl + ks * 32, 16, 1024);
                    const uint32_t acc = (it > 0 || ks > 0) ? 1u : 0u;
                    const uint32_t d2 = tmem + (uint32_t)p.nacc * BNT, d1 = tmem + (uint32_t)(it % p.nacc) * BNT;
                    umma_tf32(d2, dal, dbh, idesc, acc);
                    umma_tf32(d2, dah, dbl, idesc, 1u);
                    umma_tf32(d1, dah, dbh, idesc, (it >= p.nacc || ks > 0) ? 1u : 0u);
                }
                umma_commit(&bar_empty[s]);
            }
            umma_commit(&bar_accum);
        }
    } else {
        const int tid = threadIdx.x - 64;
        for (int it = 0; it < iters; ++it) {
            const int s = it % kStages, ph = (it / kStages) & 1;
            mbar_wait(&bar_full[s], ph, 3);
            split_region(reinterpret_cast<float4*>(a_hi(s)), reinterpret_cast<float4*>(a_lo(s)), kABytes / 16, tid);
            fence_proxy_async_smem();
            mbar_arrive(&bar_split[s]);
        }
        // ---- epilogue: TMEM -> registers -> bias + activation -> global
        mbar_wait(&bar_accum, 0, 4);
        tcgen05_fence_after();
        const int q = warp & 3, row = q * 32 + lane;
        const int lw = row % p.bw, lh = (row / p.bw) % p.bh, ln = row / (p.bw * p.bh);
        const int ox = x0 + lw, oy = y0 + lh, on = n0 + ln;
        const bool valid = on < p.n && oy < p.oh && ox < p.ow;
#pragma unroll 1
        for (int c = 0; c < BNT / 32; ++c) {
            float v[32], v2[32];
            tmem_ld_32x32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), v);
            for (int a2 = 1; a2 <= p.nacc; ++a2) {
                tmem_ld_32x32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(a2 * BNT + c * 32), v2);
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] += v2[j];
            }
            if (!valid) continue;
            const int co0 = ntile0 + c * 32;
            float* o;
            if (p.out_mode == NI_MODE_PLAIN) {
                o = p.out + (((long long)on * p.oh + oy) * p.ow + ox) * p.out_pitch + p.out_coff + co0;
            } else {
                const int F = p.ntot >> 2, blk = co0 / F, f0 = co0 - blk * F;
                o = p.out + (((long long)on * 2 * p.oh + 2 * oy + (blk >> 1)) * (2 * p.ow) + 2 * ox + (blk & 1)) * p.out_pitch + p.out_coff + f0;
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                float t = v[j];
                if (p.bias) { const int bi = co0 + j; t += __ldg(p.bias + (p.bias_mod > 0 ? bi % p.bias_mod : bi)); }
                v[j] = apply_act_tc(t, p.act, p.alpha);
            }
            float4* o4 = reinterpret_cast<float4*>(o);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float4 w = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                if (p.accumulate) { const float4 old = o4[j]; w.x += old.x; w.y += old.y; w.z += old.z; w.w += old.w; }
                o4[j] = w;
            }
        }
        tcgen05_fence_before();
    }
    __syncthreads();
    if (warp == 1) { tcgen05_fence_after(); tmem_dealloc(tmem, TMEM_COLS); }
}

struct TcWgradParams {
    int n, oh, ow;                 // dY dims
    int bw, bh, bn;                // 32-pixel tile
    int tiles_w, tiles_h;
    int kw, pad_t, pad_l;
    int cin_chunks;                // Cin / 32
    int atoms;                     // taps * cin_chunks  (32-row atoms of the flattened (tap, ci) M index)
    int mtot;                      // taps * Cin
    int cout;
    int steps_total, steps_per_split;
    float* dw;
};

constexpr int kWgStages = 2;
constexpr int kThreadsWg = 64 + 256;   // TMA warp, MMA warp, 8 transpose/split warps (the first 4 also run the epilogue)

// Transpose + hi/lo split of one TMA-landed atom: raw [32 pixels][32 channels] (SWIZZLE_128B, pixel rows) ->
// K-major operand rows [32 channels][32 pixels] (SWIZZLE_128B, channel rows) at row offset `row0` of the hi / lo tiles.
// kind::tf32 only accepts K-major operands (MN-major descriptors return zeros: measured with ni_tc_selftest), so the
// pixel-major activation tiles have to be turned around in shared memory. One float4 read (4 channels of a pixel) ->
// 2 x 4 scalar stores; both sides are bank-conflict free thanks to the 128-byte XOR swizzle.
__device__ __forceinline__ void transpose_split_chunk(const uint8_t* raw, uint8_t* hi, uint8_t* lo, int row0, int r) {
    const int p = r & 31, c4 = r >> 5;
    const float4 v = *reinterpret_cast<const float4*>(raw + p * 128 + ((c4 ^ (p & 7)) << 4));
    const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int m = row0 + c4 * 4 + e;
        const int off = (m >> 3) * 1024 + (m & 7) * 128 + (((p >> 2) ^ (m & 7)) << 4) + (p & 3) * 4;
        const float h = __uint_as_float(__float_as_uint(vv[e]) & 0xFFFFE000u);
        *reinterpret_cast<float*>(hi + off) = h;
        *reinterpret_cast<float*>(lo + off) = vv[e] - h;
    }
}

template <int BNT>
__global__ void __launch_bounds__(kThreadsWg, 1)
conv_tc_wgrad_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY, const TcWgradParams p) {
    constexpr int B_BYTES = BNT * 128;
    constexpr int STAGE_BYTES = 3 * kABytes + 3 * B_BYTES;   // raw A, raw B, A hi, A lo, B hi, B lo
    constexpr uint32_t TMEM_COLS = BNT < 32 ? 32 : BNT;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar_full[kWgStages], bar_split[kWgStages], bar_empty[kWgStages], bar_accum;
    __shared__ uint32_t tmem_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int atom0 = blockIdx.x * 4;
    const int valid_atoms = min(4, p.atoms - atom0);
    const int co0 = blockIdx.y * BNT;
    const int step0 = blockIdx.z * p.steps_per_split;
    const int iters = min(p.steps_per_split, p.steps_total - step0);

    if (threadIdx.x == 0) {
        for (int s = 0; s < kWgStages; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_split[s], 256); mbar_init(&bar_empty[s], 1); }
        mbar_init(&bar_accum, 1);
        fence_barrier_init();
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmDY);
    }
    if (warp == 1) tmem_alloc(&tmem_slot, TMEM_COLS);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = tmem_slot;

    auto raw_a = [&](int s) { return smem + s * STAGE_BYTES; };
    auto raw_b = [&](int s) { return smem + s * STAGE_BYTES + kABytes; };
    auto a_hi = [&](int s) { return smem + s * STAGE_BYTES + kABytes + B_BYTES; };
    auto a_lo = [&](int s) { return smem + s * STAGE_BYTES + 2 * kABytes + B_BYTES; };
    auto b_hi = [&](int s) { return smem + s * STAGE_BYTES + 3 * kABytes + B_BYTES; };
    auto b_lo = [&](int s) { return smem + s * STAGE_BYTES + 3 * kABytes + 2 * B_BYTES; };

    if (warp == 0) {
        if (lane == 0) {
            for (int it = 0; it < iters; ++it) {
                const int s = it % kWgStages, ph = (it / kWgStages) & 1;
                mbar_wait(&bar_empty[s], ph ^ 1, 0);
                mbar_expect_tx(&bar_full[s], valid_atoms * 4096 + B_BYTES);
                const int st = step0 + it;
                const int tw = st % p.tiles_w, th = (st / p.tiles_w) % p.tiles_h, tn = st / (p.tiles_w * p.tiles_h);
                const int x0 = tw * p.bw, y0 = th * p.bh, n0 = tn * p.bn;
                for (int j = 0; j < valid_atoms; ++j) {
                    const int atom = atom0 + j, tap = atom / p.cin_chunks, cc = atom - tap * p.cin_chunks;
                    const int a = tap / p.kw, b = tap - a * p.kw;
                    tma_load_4d(raw_a(s) + j * 4096, &tmX, &bar_full[s], cc * 32, x0 + b - p.pad_l, y0 + a - p.pad_t, n0);
                }
                for (int j = 0; j < BNT / 32; ++j) tma_load_4d(raw_b(s) + j * 4096, &tmDY, &bar_full[s], co0 + j * 32, x0, y0, n0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_tf32(128, BNT, 0, 0);
            for (int it = 0; it < iters; ++it) {
                const int s = it % kWgStages, ph = (it / kWgStages) & 1;
                mbar_wait(&bar_split[s], ph, 2);
                tcgen05_fence_after();
                const uint32_t ah = smem_u32(a_hi(s)), al = smem_u32(a_lo(s)), bh = smem_u32(b_hi(s)), bl = smem_u32(b_lo(s));
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {   // 32 pixels per stage = 4 x K=8
                    const uint64_t dah = make_smem_desc_sw128(ah + ks * 32, 16, 1024), dal = make_smem_desc_sw128(al + ks * 32, 16, 1024);
                    const uint64_t dbh = make_smem_desc_sw128(bh + ks * 32, 16, 1024), dbl = make_smem_desc_sw128(bl + ks * 32, 16, 1024);
                    umma_tf32(tmem, dal, dbh, idesc, (it > 0 || ks > 0) ? 1u : 0u);
                    umma_tf32(tmem, dah, dbl, idesc, 1u);
                    umma_tf32(tmem, dah, dbh, idesc, 1u);
                }
                umma_commit(&bar_empty[s]);
            }
            umma_commit(&bar_accum);
        }
    } else {
        const int tid = threadIdx.x - 64;
        const int chunks_a = valid_atoms * 256, chunks = chunks_a + (BNT / 32) * 256;
        for (int it = 0; it < iters; ++it) {
            const int s = it % kWgStages, ph = (it / kWgStages) & 1;
            mbar_wait(&bar_full[s], ph, 3);
            for (int q = tid; q < chunks; q += 256) {
                if (q < chunks_a) transpose_split_chunk(raw_a(s) + (q >> 8) * 4096, a_hi(s), a_lo(s), (q >> 8) * 32, q & 255);
                else { const int qb = q - chunks_a; transpose_split_chunk(raw_b(s) + (qb >> 8) * 4096, b_hi(s), b_lo(s), (qb >> 8) * 32, qb & 255); }
            }
            fence_proxy_async_smem();
            mbar_arrive(&bar_split[s]);
        }
        if (warp < 6) {
            mbar_wait(&bar_accum, 0, 4);
            tcgen05_fence_after();
            const int q = warp & 3, row = q * 32 + lane;
            const int mm = blockIdx.x * 128 + row;
            const bool valid = mm < p.mtot && iters > 0;
#pragma unroll 1
            for (int c = 0; c < BNT / 32; ++c) {
                float v[32];
                tmem_ld_32x32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), v);
                if (!valid) continue;
                float* o = p.dw + (long long)mm * p.cout + co0 + c * 32;
#pragma unroll
                for (int j = 0; j < 32; ++j) atomicAdd(o + j, v[j]);
            }
            tcgen05_fence_before();
        }
    }
    __syncthreads();
    if (warp == 1) { tcgen05_fence_after(); tmem_dealloc(tmem, TMEM_COLS); }
}

// w (taps, cin, cout) HWIO -> out (2, taps, N, K) K-major hi / lo. transpose: N = cout, K = cin (fprop); else N = cin, K = cout.
__global__ void tc_prep_weights_kernel(const float* __restrict__ w, float* __restrict__ out, int taps, int cin, int cout, int transpose) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    const long long total = (long long)taps * cin * cout;
    if (i >= total) return;
    float v;
    if (transpose) {
        const int ci = (int)(i % cin);
        const long long t = i / cin;
        const int co = (int)(t % cout), tap = (int)(t / cout);
        v = w[((long long)tap * cin + ci) * cout + co];
    } else {
        v = w[i];
    }
    const float h = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    out[i] = h;
    out[total + i] = v - h;
}

// ---------------------------------------------------------------- host side
std::mutex g_scratch_mutex;
float* g_scratch = nullptr;
size_t g_scratch_bytes = 0;

int get_scratch(size_t bytes, float** out) {
    std::lock_guard<std::mutex> lock(g_scratch_mutex);
    if (bytes > g_scratch_bytes) {
        // grown rarely (largest layer first seen); a device sync keeps in-flight users of the old buffer safe
        NI_CUDA(cudaDeviceSynchronize());
        if (g_scratch) NI_CUDA(cudaFree(g_scratch));
        g_scratch = nullptr; g_scratch_bytes = 0;
        size_t want = bytes < (size_t)(32u << 20) ? (size_t)(32u << 20) : bytes;
        NI_CUDA(cudaMalloc(&g_scratch, want));
        g_scratch_bytes = want;
    }
    *out = g_scratch;
    return NI_OK;
}

int encode_map(CUtensorMap* tm, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes, const cuuint32_t* box) {
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = cuTensorMapEncodeTiled(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, const_cast<void*>(base), dims, strides_bytes, box, estr,
                                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        const char* msg = nullptr;
        cuGetErrorString(r, &msg);
        ni_set_error("cuTensorMapEncodeTiled failed: %s", msg ? msg : "?");
        return NI_ERR_CUDA;
    }
    return NI_OK;
}

// NHWC activation view (n, h, w, c) with channel pitch: 4-D map, box (32, bw, bh, bn)
int encode_act_map(CUtensorMap* tm, const float* base, int n, int h, int w, int c, int pitch, int bw, int bh, int bn) {
    cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
    cuuint64_t str[3] = {(cuuint64_t)pitch * 4, (cuuint64_t)w * pitch * 4, (cuuint64_t)h * w * pitch * 4};
    cuuint32_t box[4] = {32, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn};
    return encode_map(tm, base, 4, dims, str, box);
}

bool pick_tile(int h, int w, int pixels, int& bw, int& bh, int& bn) {
    if (w >= pixels) { if (w % pixels) return false; bw = pixels; bh = 1; bn = 1; return true; }
    if (pixels % w) return false;
    bw = w;
    const int rows = pixels / w;
    if (h >= rows) { if (h % rows) return false; bh = rows; bn = 1; return true; }
    if (rows % h) return false;
    bh = h; bn = rows / h;
    return bn <= 256;
}

int pick_bnt(int n) { return n % 128 == 0 ? 128 : (n % 64 == 0 ? 64 : (n % 32 == 0 ? 32 : 0)); }

template <typename K>
int set_dyn_smem(K kern, size_t bytes) {
    NI_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return NI_OK;
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// Shared launcher for fprop (dgrad = false) and dgrad (dgrad = true).
int launch_gemm(const ni_conv_desc* d, bool dgrad, const float* src, const float* w, const float* bias, float* dst, cudaStream_t st) {
    const int K = dgrad ? d->cout : d->cin, N = dgrad ? d->cin : d->cout;
    const int sh = dgrad ? d->oh : d->h, sw = dgrad ? d->ow : d->w;           // source dims
    const int th = dgrad ? d->h : d->oh, tw = dgrad ? d->w : d->ow;           // target dims
    const int spitch = dgrad ? d->out_pitch : d->in_pitch, scoff = dgrad ? d->out_coff : d->in_coff;
    const int dpitch = dgrad ? d->in_pitch : d->out_pitch, dcoff = dgrad ? d->in_coff : d->out_coff;
    const int dmode = dgrad ? d->in_mode : d->out_mode;
    const int taps = d->kh * d->kw;
    TcGemmParams p;
    if (!pick_tile(th, tw, 128, p.bw, p.bh, p.bn)) { ni_set_error("conv_tc: unsupported spatial tile"); return NI_ERR_UNSUPPORTED; }
    const int bnt = pick_bnt(N);
    float* scratch = nullptr;
    const size_t wbytes = (size_t)2 * taps * K * N * sizeof(float);
    int rc = get_scratch(wbytes, &scratch);
    if (rc) return rc;
    const long long total = (long long)taps * K * N;
    tc_prep_weights_kernel<<<ni_cdiv(total, 256), 256, 0, st>>>(w, scratch, taps, d->cin, d->cout, dgrad ? 0 : 1);
    NI_LAUNCH_CHECK();
    CUtensorMap tmA, tmB;
    rc = encode_act_map(&tmA, src + scoff, d->n, sh, sw, K, spitch, p.bw, p.bh, p.bn);
    if (rc) return rc;
    {
        cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)N, (cuuint64_t)(2 * taps)};
        cuuint64_t str[2] = {(cuuint64_t)K * 4, (cuuint64_t)K * N * 4};
        cuuint32_t box[3] = {32, (cuuint32_t)bnt, 1};
        rc = encode_map(&tmB, scratch, 3, dims, str, box);
        if (rc) return rc;
    }
    p.n = d->n; p.oh = th; p.ow = tw;
    p.tiles_w = tw / p.bw; p.tiles_h = th / p.bh;
    const int tiles_n = (d->n + p.bn - 1) / p.bn;
    p.kh = d->kh; p.kw = d->kw;
    p.off_y0 = dgrad ? d->pad_t : -d->pad_t; p.off_x0 = dgrad ? d->pad_l : -d->pad_l; p.off_sign = dgrad ? -1 : 1;
    p.kchunks = K / 32; p.ntot = N;
    p.out_pitch = dpitch; p.out_coff = dcoff; p.out_mode = dmode;
    p.bias_mod = dgrad ? 0 : d->bias_mod; p.act = dgrad ? NI_ACT_NONE : d->act; p.accumulate = d->accumulate; p.alpha = d->act_alpha;
    p.bias = dgrad ? nullptr : bias; p.out = dst;
    {
        // accumulator rotation: deeper contractions get more D1 accumulators, bounded by the 512 TMEM columns
        const int ktot = taps * K;
        int nacc = ktot > 2304 ? 3 : (ktot > 1024 ? 2 : 1);
        while ((nacc + 1) * bnt > 512) --nacc;
        if (nacc > taps * (K / 32)) nacc = taps * (K / 32);
        p.nacc = nacc < 1 ? 1 : nacc;
    }
    dim3 grid((unsigned)(p.tiles_w * p.tiles_h * tiles_n), (unsigned)(N / bnt));
#define NI_TC_GEMM(B)                                                                              \
    {                                                                                              \
        const size_t smem = (size_t)kStages * (2 * kABytes + 2 * B * 128) + 1024;                  \
        rc = set_dyn_smem(conv_tc_gemm_kernel<B>, smem);                                           \
        if (rc) return rc;                                                                         \
        conv_tc_gemm_kernel<B><<<grid, kThreadsTc, smem, st>>>(tmA, tmB, p);                       \
    }
    if (bnt == 128) NI_TC_GEMM(128) else if (bnt == 64) NI_TC_GEMM(64) else NI_TC_GEMM(32)
#undef NI_TC_GEMM
    NI_LAUNCH_CHECK();
    NI_COUNT_LAUNCH(2);
    return NI_OK;
}

}  // namespace

// Second internal scratch (transposed weights of the SIMT dgrad path), same grow-on-demand policy.
int ni_get_scratch2(size_t bytes, float** out) {
    static float* buf = nullptr;
    static size_t cap = 0;
    std::lock_guard<std::mutex> lock(g_scratch_mutex);
    if (bytes > cap) {
        NI_CUDA(cudaDeviceSynchronize());
        if (buf) NI_CUDA(cudaFree(buf));
        buf = nullptr; cap = 0;
        const size_t want = bytes < (size_t)(16u << 20) ? (size_t)(16u << 20) : bytes;
        NI_CUDA(cudaMalloc(&buf, want));
        cap = want;
    }
    *out = buf;
    return NI_OK;
}

// 1 if the tcgen05 path handles this problem. op: 0 fprop, 1 dgrad, 2 wgrad.
extern "C" int ni_conv2d_tc_supported(const ni_conv_desc* d, int op) {
    if (!d || d->n <= 0) return 0;
    if (d->stride != 1 || d->pad_mode != NI_PAD_ZERO) return 0;
    if (d->cin % 32 || d->cout % 32) return 0;
    if (d->kh * d->kw > 64) return 0;
    if ((d->in_pitch % 4) || (d->in_coff % 4) || (d->out_pitch % 4) || (d->out_coff % 4)) return 0;
    int bw, bh, bn;
    if (op == 0) {
        if (d->in_mode != NI_MODE_PLAIN) return 0;
        if (d->out_mode == NI_MODE_BLOCK2 && ((d->cout / 4) % 32)) return 0;
        return pick_tile(d->oh, d->ow, 128, bw, bh, bn) ? 1 : 0;
    }
    if (op == 1) {
        if (d->out_mode != NI_MODE_PLAIN) return 0;                    // dy is the TMA source
        if (d->in_mode == NI_MODE_BLOCK2 && ((d->cin / 4) % 32)) return 0;
        return pick_tile(d->h, d->w, 128, bw, bh, bn) ? 1 : 0;
    }
    if (d->in_mode != NI_MODE_PLAIN || d->out_mode != NI_MODE_PLAIN) return 0;
    if (!pick_tile(d->oh, d->ow, 32, bw, bh, bn)) return 0;
    return (d->n % bn) == 0 ? 1 : 0;
}

extern "C" int ni_conv2d_fprop_tc(const ni_conv_desc* d, const float* x, const float* w, const float* bias, float* y, cudaStream_t st) {
    NI_REQUIRE(ni_conv2d_tc_supported(d, 0), "ni_conv2d_fprop_tc: problem not supported by the tcgen05 path");
    NI_REQUIRE(x && w && y && aligned16(x) && aligned16(y), "ni_conv2d_fprop_tc: null or unaligned pointer");
    return launch_gemm(d, false, x, w, bias, y, st);
}

// w is the HWIO weight tensor itself (no transposed copy needed: (tap, cin, cout) is already K-major for dgrad).
extern "C" int ni_conv2d_dgrad_tc(const ni_conv_desc* d, const float* dy, const float* w, float* dx, cudaStream_t st) {
    NI_REQUIRE(ni_conv2d_tc_supported(d, 1), "ni_conv2d_dgrad_tc: problem not supported by the tcgen05 path");
    NI_REQUIRE(dy && w && dx && aligned16(dy) && aligned16(dx), "ni_conv2d_dgrad_tc: null or unaligned pointer");
    return launch_gemm(d, true, dy, w, nullptr, dx, st);
}

extern "C" int ni_conv2d_wgrad_tc(const ni_conv_desc* d, const float* x, const float* dy, float* dw, cudaStream_t st) {
    NI_REQUIRE(ni_conv2d_tc_supported(d, 2), "ni_conv2d_wgrad_tc: problem not supported by the tcgen05 path");
    NI_REQUIRE(x && dy && dw && aligned16(x) && aligned16(dy), "ni_conv2d_wgrad_tc: null or unaligned pointer");
    const int taps = d->kh * d->kw;
    if (!d->accumulate) NI_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)taps * d->cin * d->cout, st));
    TcWgradParams p;
    pick_tile(d->oh, d->ow, 32, p.bw, p.bh, p.bn);
    CUtensorMap tmX, tmDY;
    int rc = encode_act_map(&tmX, x + d->in_coff, d->n, d->h, d->w, d->cin, d->in_pitch, p.bw, p.bh, p.bn);
    if (rc) return rc;
    rc = encode_act_map(&tmDY, dy + d->out_coff, d->n, d->oh, d->ow, d->cout, d->out_pitch, p.bw, p.bh, p.bn);
    if (rc) return rc;
    p.n = d->n; p.oh = d->oh; p.ow = d->ow;
    p.tiles_w = d->ow / p.bw; p.tiles_h = d->oh / p.bh;
    p.kw = d->kw; p.pad_t = d->pad_t; p.pad_l = d->pad_l;
    p.cin_chunks = d->cin / 32; p.atoms = taps * p.cin_chunks; p.mtot = taps * d->cin; p.cout = d->cout;
    p.steps_total = p.tiles_w * p.tiles_h * (d->n / p.bn);
    p.dw = dw;
    const int bnt = pick_bnt(d->cout);
    const int mtiles = (p.atoms + 3) / 4, ntiles = d->cout / bnt;
    int splits = (2 * ni_num_sms() + mtiles * ntiles - 1) / (mtiles * ntiles);
    const int max_splits = (p.steps_total + 15) / 16;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    p.steps_per_split = (p.steps_total + splits - 1) / splits;
    splits = (p.steps_total + p.steps_per_split - 1) / p.steps_per_split;
    dim3 grid((unsigned)mtiles, (unsigned)ntiles, (unsigned)splits);
#define NI_TC_WGRAD(B)                                                                             \
    {                                                                                              \
        const size_t smem = (size_t)kWgStages * (3 * kABytes + 3 * B * 128) + 1024;                \
        rc = set_dyn_smem(conv_tc_wgrad_kernel<B>, smem);                                          \
        if (rc) return rc;                                                                         \
        conv_tc_wgrad_kernel<B><<<grid, kThreadsWg, smem, st>>>(tmX, tmDY, p);                     \
    }
    if (bnt == 128) NI_TC_WGRAD(128) else if (bnt == 64) NI_TC_WGRAD(64) else NI_TC_WGRAD(32)
#undef NI_TC_WGRAD
    NI_LAUNCH_CHECK();
    NI_COUNT_LAUNCH(1);
    return NI_OK;
}
