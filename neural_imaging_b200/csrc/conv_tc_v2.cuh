// tcgen05 convolution kernels, generation 2: the A operand travels through TENSOR MEMORY.
//
// Generation 1 (conv_tc.cu) kept A in shared memory: per 32-channel k-iteration the 16 KB A tile was written by TMA,
// read + re-written as hi / lo by the splitter (48 KB) and read 12 more times by the MMAs (48 KB): 112+ KB of shared-
// memory traffic per k-iteration against a 128 B/clk port, i.e. 2x (N = 128) to 5x (N = 32) the MMA time.
// Here converter warps read each pixel row ONCE from the TMA-landed tile (8 conflict-free 128-bit loads through the
// XOR swizzle), split it in registers and write hi and lo to TMEM with tcgen05.st; tcgen05.mma then takes A from TMEM
// ([d], [a], b-desc form, validated by ni_tc_selftest) and only B (weights) is read from shared memory.
// Shared-memory traffic per k-iteration drops to 32 KB + B tiles.
#pragma once
#include "conv_desc.h"
#include "ni_common.cuh"
#include "tc_common.cuh"

namespace tcv2 {
using namespace tc;

constexpr int kTmemSlots = 2;    // A (hi | lo) slots in tensor memory, 64 columns each
constexpr int kAraw = 16384;
// warp 0: A producer, 1: MMA issuer, 2 .. 2+NCW-1: converters + epilogue, last: B producer. NCW = 8 for the N = 128 tile
// (+20 %: one warp per scheduler cannot hide its own latencies), NCW = 4 for N <= 64 so that two CTAs share an SM.
template <int BNT> struct Roles {
    static constexpr int NCW = BNT == 128 ? 8 : 4;
    static constexpr int THREADS = (3 + NCW) * 32;
    static constexpr int MIN_CTAS = BNT == 128 ? 1 : 2;   // caps registers at 128 so that two 7-warp CTAs really co-reside
};

struct GemmParams {
    int n, oh, ow;
    int bw, bh, bn;
    int tiles_w, tiles_h;
    int kh, kw, off_y0, off_x0, off_sign;
    int hw, hh, sa, a_stage;       // halo box (bw + kw - 1) x (bh + kh - 1) pixels, A stages and their byte size (1024-aligned)
    int kchunks, ntot;
    int n_valid;                   // real output channels (< ntot when the n tile is zero-padded: scalar masked stores in the epilogue)
    int out_pitch, out_coff, out_mode;
    int bias_mod, act, accumulate, nacc;
    float alpha;
    const float* bias;
    float* out;
    int src_block2_f;              // > 0: the source is read through depth_to_space(2) of an (n, 2h, 2w, F) buffer (5-D tensor map
                                   // (F, 2, w, 2, h*n); only for 1x1 filters = the transposed convolutions), F = src_block2_f
    // dgrad fused with the PRODUCING layer's activation backward (generation 3 only): the result is d(loss)/d(output y of the layer below);
    // the epilogue multiplies it by act'(y) and accumulates that layer's bias gradient (column sums), so that no separate elementwise pass
    // (ni_act_bwd_bias: read y, read + write dy) is needed. dact_y = nullptr: off.
    const float* dact_y;
    int dact_pitch, dact_coff, dact;
    float dact_alpha;
    float* dbias;
    int dbias_mod;
};

__device__ __forceinline__ float act_apply(float v, int act, float alpha) {
    switch (act) {
        case NI_ACT_LEAKY_RELU: return v > 0.f ? v : alpha * v;
        case NI_ACT_RELU: return fmaxf(v, 0.f);
        case NI_ACT_TANH: return tanhf(v);
        case NI_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
        case NI_ACT_CLIP01: return ni_clamp01(v);
        default: return v;
    }
}

__device__ __forceinline__ uint32_t pow2_cols(uint32_t want) {
    return want <= 32 ? 32 : (want <= 64 ? 64 : (want <= 128 ? 128 : (want <= 256 ? 256 : 512)));
}

// Ring depths. The kernel is bound by how many bytes of TMA loads one SM keeps in flight (ncu: converters stalled on the
// "tile landed" barrier 26 % of all samples, L2 at 17 %, tensor pipe at 27 % with a single 3-deep ring): activations stream
// from HBM and need depth, weights are L2-resident and big, so the two operands get their own rings.
// A is staged as a HALO tile: for one 32-channel chunk the (bw + kw - 1) x (bh + kh - 1) input pixels that all kh*kw taps of
// the 128-pixel tile touch are loaded ONCE (one TMA box) and every tap reads its shifted window from shared memory while
// converting to TMEM. Per k-iteration the per-tap design moved 16 KB of activations + 2*BNT*128 B of weights through TMA,
// and TMA ingest (~25 B/clk per CTA, tools/tma_probe.py) bounded the kernel at 2-5x the MMA time; now activations cost
// 16 KB * halo_ratio / taps per iteration.
constexpr int kMaxSA = 2;
template <int BNT> struct Rings {
    static constexpr int SB = 3;                                // B stages, hi + lo = 2 * BNT * 128 bytes each
    static constexpr int B_BYTES = BNT * 128;
    static constexpr int SMEM_B = SB * 2 * B_BYTES;
};

// TMEM map: [0, (nacc+1)*BNT) accumulators (D1_0.. D1_{nacc-1}, D2), then kTmemSlots x 64 columns of A (hi 32 | lo 32).
template <int BNT>
__global__ void __launch_bounds__(Roles<BNT>::THREADS, Roles<BNT>::MIN_CTAS)
conv_tc2_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const float* __restrict__ wtiled, const GemmParams p) {
    constexpr int SA = kMaxSA, SB = Rings<BNT>::SB, B_BYTES = Rings<BNT>::B_BYTES, NCW = Roles<BNT>::NCW;
    const uint32_t acc_cols = (uint32_t)(p.nacc + 1) * BNT;
    const uint32_t TMEM_COLS = pow2_cols(acc_cols + kTmemSlots * 64);
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar_afull[SA], bar_afree[SA], bar_bfull[SB], bar_bfree[SB], bar_tready[kTmemSlots], bar_tfree[kTmemSlots], bar_accum;
    __shared__ uint32_t tmem_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tw = blockIdx.x % p.tiles_w, th = (blockIdx.x / p.tiles_w) % p.tiles_h, tn = blockIdx.x / (p.tiles_w * p.tiles_h);
    const int x0 = tw * p.bw, y0 = th * p.bh, n0 = tn * p.bn;
    const int ntile0 = blockIdx.y * BNT;
    const int taps = p.kh * p.kw;
    const int iters = taps * p.kchunks;

    if (threadIdx.x == 0) {
        for (int s = 0; s < SA; ++s) { mbar_init(&bar_afull[s], 1); mbar_init(&bar_afree[s], NCW * 32); }
        for (int s = 0; s < SB; ++s) { mbar_init(&bar_bfull[s], 1); mbar_init(&bar_bfree[s], 1); }
        for (int t = 0; t < kTmemSlots; ++t) { mbar_init(&bar_tready[t], NCW * 32); mbar_init(&bar_tfree[t], 1); }
        mbar_init(&bar_accum, 1);
        fence_barrier_init();
        tma_prefetch_desc(&tmA);
    }
    if (warp == 1) tmem_alloc(&tmem_slot, TMEM_COLS);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t a_base = tmem + acc_cols;

    auto a_halo = [&](int s) { return smem + s * p.a_stage; };
    auto b_hi = [&](int s) { return smem + p.sa * p.a_stage + s * 2 * B_BYTES; };
    auto b_lo = [&](int s) { return smem + p.sa * p.a_stage + s * 2 * B_BYTES + B_BYTES; };
    // iteration it = kc * taps + tap (channel chunk outer, taps inner: one halo tile serves all taps)

    if (warp == 0) {
        if (lane == 0) {   // ---- A producer: one halo box per 32-channel chunk
            const int bx = x0 + p.off_x0 + (p.off_sign < 0 ? -(p.kw - 1) : 0), by = y0 + p.off_y0 + (p.off_sign < 0 ? -(p.kh - 1) : 0);
            const uint32_t bytes = (uint32_t)(p.hw * p.hh * p.bn) * 128u;
            for (int kc = 0; kc < p.kchunks; ++kc) {
                const int s = kc % p.sa, ph = (kc / p.sa) & 1;
                mbar_wait(&bar_afree[s], ph ^ 1, 0);
                mbar_expect_tx(&bar_afull[s], bytes);
                tma_load_4d(a_halo(s), &tmA, &bar_afull[s], kc * 32, bx, by, n0);
            }
        }
    } else if (warp == 2 + NCW) {
        if (lane == 0) {   // ---- B producer
            for (int it = 0; it < iters; ++it) {
                const int s = it % SB, ph = (it / SB) & 1;
                mbar_wait(&bar_bfree[s], ph ^ 1, 7);
                mbar_expect_tx(&bar_bfull[s], 2 * B_BYTES);
                // weights are pre-tiled by the prep kernel: block (n-tile, kc, tap) = [hi tile | lo tile], already in the swizzled
                // K-major order, so one contiguous bulk copy replaces two 128-row tensor boxes (TMA cost is per box row)
                const int kc = it / taps, tap = it - kc * taps;
                const float* src = wtiled + ((size_t)(blockIdx.y * p.kchunks + kc) * taps + tap) * (size_t)(2 * BNT * 32);
                bulk_load_1d(b_hi(s), src, 2 * B_BYTES, &bar_bfull[s]);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {   // ---- MMA issuer
            constexpr uint32_t idesc = make_idesc_tf32(128, BNT, 0, 0);
            for (int it = 0; it < iters; ++it) {
                const int s = it % SB, ph = (it / SB) & 1;
                const int t = it % kTmemSlots, pt = (it / kTmemSlots) & 1;
                mbar_wait(&bar_bfull[s], ph, 1);
                mbar_wait(&bar_tready[t], pt, 2);
                tcgen05_fence_after();
                const uint32_t bh = smem_u32(b_hi(s)), bl = smem_u32(b_lo(s));
                const uint32_t ahi = a_base + t * 64, alo = ahi + 32;
                const uint32_t d2 = tmem + (uint32_t)p.nacc * BNT, d1 = tmem + (uint32_t)(it % p.nacc) * BNT;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const uint64_t dbh = make_smem_desc_sw128(bh + ks * 32, 16, 1024), dbl = make_smem_desc_sw128(bl + ks * 32, 16, 1024);
                    const uint32_t acc = (it > 0 || ks > 0) ? 1u : 0u;
                    umma_tf32_ts(d2, alo + ks * 8, dbh, idesc, acc);
                    umma_tf32_ts(d2, ahi + ks * 8, dbl, idesc, 1u);
                    umma_tf32_ts(d1, ahi + ks * 8, dbh, idesc, (it >= p.nacc || ks > 0) ? 1u : 0u);
                }
                umma_commit(&bar_bfree[s]);
                umma_commit(&bar_tfree[t]);
            }
            umma_commit(&bar_accum);
        }
    } else {
        // ---- converters (warps 2-9): A tile row -> registers -> hi / lo -> TMEM. Two warps share each TMEM lane quarter
        // (a warp may only touch lanes [32 (warp % 4), +32)) and take 16 of the 32 K-columns each: with one warp per
        // quarter the conversion (one warp per scheduler, nothing to hide its latencies) was slower than the MMAs it feeds.
        constexpr int NH = NCW / 4;            // warps per TMEM lane quarter (1 or 2)
        constexpr int CW = 32 / NH;            // K-columns converted by one thread
        const int q = warp & 3, row = q * 32 + lane, half = (warp - 2) >> 2;
        // this thread's pixel inside the halo box (tap (0,0) position); tap (a, b) adds (sy(a) * hw + sx(b)) pixels
        const int prow0 = ((row / (p.bw * p.bh)) * p.hh + (row / p.bw) % p.bh) * p.hw + row % p.bw;
        for (int it = 0; it < iters; ++it) {
            const int kc = it / taps, tap = it - kc * taps;
            const int s = kc % p.sa, ph = (kc / p.sa) & 1;
            const int t = it % kTmemSlots, pt = (it / kTmemSlots) & 1;
            if (tap == 0) mbar_wait(&bar_afull[s], ph, 3);
            const int ta = tap / p.kw, tb = tap - ta * p.kw;
            const int prow = prow0 + (p.off_sign > 0 ? ta : p.kh - 1 - ta) * p.hw + (p.off_sign > 0 ? tb : p.kw - 1 - tb);
            float hi[CW], lo[CW];
            const uint8_t* rp = a_halo(s) + prow * 128;
#pragma unroll
            for (int c = 0; c < CW / 4; ++c) {
                const float4 v = *reinterpret_cast<const float4*>(rp + ((((CW / 4) * half + c) ^ (prow & 7)) << 4));
                const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float h = __uint_as_float(__float_as_uint(vv[e]) & 0xFFFFE000u);
                    hi[4 * c + e] = h;
                    lo[4 * c + e] = vv[e] - h;
                }
            }
            if (tap == taps - 1) mbar_arrive(&bar_afree[s]);   // last tap is in registers: hand the stage back to the A producer
            mbar_wait(&bar_tfree[t], pt ^ 1, 5);
            tcgen05_fence_after();
            const uint32_t dst = a_base + ((uint32_t)(q * 32) << 16) + t * 64 + CW * half;
            if constexpr (CW == 32) { tmem_st_32x32(dst, hi); tmem_st_32x32(dst + 32, lo); }
            else { tmem_st_32x16(dst, hi); tmem_st_32x16(dst + 32, lo); }
            tmem_st_wait();
            tcgen05_fence_before();
            mbar_arrive(&bar_tready[t]);
        }
        // ---- epilogue
        mbar_wait(&bar_accum, 0, 4);
        tcgen05_fence_after();
        const int lw = row % p.bw, lh = (row / p.bw) % p.bh, ln = row / (p.bw * p.bh);
        const int ox = x0 + lw, oy = y0 + lh, on = n0 + ln;
        const bool valid = on < p.n && oy < p.oh && ox < p.ow;
#pragma unroll 1
        for (int c = half; c < BNT / 32; c += NH) {   // the warps of a lane quarter interleave the 32-column chunks
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
                float tv = v[j];
                if (p.bias) { const int bi = co0 + j; tv += __ldg(p.bias + (p.bias_mod > 0 ? bi % p.bias_mod : bi)); }
                v[j] = act_apply(tv, p.act, p.alpha);
            }
            float4* o4 = reinterpret_cast<float4*>(o);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float4 w4 = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                if (p.accumulate) { const float4 old = o4[j]; w4.x += old.x; w4.y += old.y; w4.z += old.z; w4.w += old.w; }
                o4[j] = w4;
            }
        }
        tcgen05_fence_before();
    }
    __syncthreads();
    if (warp == 1) { tcgen05_fence_after(); tmem_dealloc(tmem, TMEM_COLS); }
}

// ------------------------------------------------------------------------------------------------------------ wgrad
// dW[(tap,ci), co] = sum_pixels X_tap[pixel, ci] * dY[pixel, co]. M = (tap, ci) flattened in 32-row atoms, N = co, K = pixels.
// A (X^T): each of the 4 A-converter warps owns one atom: lane = channel, reads its 32 pixels from the TMA-landed
// [pixel][channel] tile (conflict-free through the swizzle) and writes hi / lo rows straight to TMEM.
// B (dY^T) has to be K-major in shared memory (kind::tf32 has no MN-major mode): 4 transposer warps turn the pixel-major
// tile around with scalar stores (also conflict-free).
// Pipeline depth per tile width: a stage is 16 KB of raw A + 3 x BNT x 128 B (raw B, B hi, B lo); TMEM holds 2 x BNT accumulator
// columns + 64 A columns per stage. Measured: the loop is bound by the converter / transposer warps, not by the TMA round trip --
// BNT <= 64 is faster as two co-resident 2-stage CTAs per SM (7.35 ms) than as one 4-stage CTA (9.02 ms, FAN conv1 wgrad);
// BNT = 128 only fits one CTA per SM and gains from the third stage (78 -> 109 TFLOP/s).
template <int BNT> struct WgCfg { static constexpr int STAGES = BNT == 128 ? 3 : 2; };
constexpr int kThreadsWg = 64 + 128 + 128;

struct WgradParams {
    int n, oh, ow;
    int bw, bh, bn;
    int tiles_w, tiles_h;
    int kw, pad_t, pad_l;
    int cin_chunks, atoms, mtot, cout;
    int steps_total, steps_per_split;
    float* dw;
    int dy_block2_f;               // > 0: dy is read through depth_to_space(2) (5-D map, see GemmParams::src_block2_f)
    int defer_st;                  // generation 3: X converters complete their tcgen05.st behind the next step's loads
};

__device__ __forceinline__ void transpose_split_chunk(uint32_t raw, uint32_t hi, uint32_t lo, int row0, int r) {
    const int p = r & 31, c4 = r >> 5;
    const float4 v = lds128(raw + (uint32_t)(p * 128 + ((c4 ^ (p & 7)) << 4)));
    const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int m = row0 + c4 * 4 + e;
        const uint32_t off = (uint32_t)((m >> 3) * 1024 + (m & 7) * 128 + (((p >> 2) ^ (m & 7)) << 4) + (p & 3) * 4);
        const float h = __uint_as_float(__float_as_uint(vv[e]) & 0xFFFFE000u);
        sts32(hi + off, h);
        sts32(lo + off, vv[e] - h);
    }
}

// TMEM map: D1 [0,BNT), D2 [BNT, 2 BNT), A slots 2 x 64 columns.
template <int BNT>
__global__ void __launch_bounds__(kThreadsWg, 1)
conv_tc2_wgrad_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY, const WgradParams p) {
    constexpr int B_BYTES = BNT * 128;
    constexpr int kWgStages = WgCfg<BNT>::STAGES;
    constexpr int STAGE_BYTES = kAraw + 3 * B_BYTES;   // raw A (4 atoms), raw B, B hi, B lo
    const uint32_t TMEM_COLS = pow2_cols(2 * BNT + kWgStages * 64);
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar_full[kWgStages], bar_aready[kWgStages], bar_bready[kWgStages], bar_free[kWgStages], bar_accum;
    __shared__ uint32_t tmem_slot;

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int atom0 = blockIdx.x * 4;
    const int valid_atoms = min(4, p.atoms - atom0);
    const int co0 = blockIdx.y * BNT;
    const int step0 = blockIdx.z * p.steps_per_split;
    const int iters = min(p.steps_per_split, p.steps_total - step0);

    if (threadIdx.x == 0) {
        for (int s = 0; s < kWgStages; ++s) {
            mbar_init(&bar_full[s], 1); mbar_init(&bar_aready[s], 128); mbar_init(&bar_bready[s], 128); mbar_init(&bar_free[s], 1);
        }
        mbar_init(&bar_accum, 1);
        fence_barrier_init();
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmDY);
    }
    if (warp == 1) tmem_alloc(&tmem_slot, TMEM_COLS);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = __shfl_sync(0xffffffffu, tmem_slot, 0);
    const uint32_t a_base = tmem + 2 * BNT;

    auto raw_a = [&](int s) { return smem + s * STAGE_BYTES; };
    auto raw_b = [&](int s) { return smem + s * STAGE_BYTES + kAraw; };
    auto b_hi = [&](int s) { return smem + s * STAGE_BYTES + kAraw + B_BYTES; };
    auto b_lo = [&](int s) { return smem + s * STAGE_BYTES + kAraw + 2 * B_BYTES; };

    if (warp == 0) {
        if (lane == 0) {
            TCP_DECL
            for (int it = 0; it < iters; ++it) {
                const int s = it % kWgStages, ph = (it / kWgStages) & 1;
                TCP_START();
                mbar_wait(&bar_free[s], ph ^ 1, 0);
                TCP_ADD(32);
                mbar_expect_tx(&bar_full[s], valid_atoms * 4096 + B_BYTES);
                const int st = step0 + it;
                const int tw = st % p.tiles_w, th = (st / p.tiles_w) % p.tiles_h, tn = st / (p.tiles_w * p.tiles_h);
                const int x0 = tw * p.bw, y0 = th * p.bh, n0 = tn * p.bn;
                for (int j = 0; j < valid_atoms; ++j) {
                    const int atom = atom0 + j, tap = atom / p.cin_chunks, cc = atom - tap * p.cin_chunks;
                    const int a = tap / p.kw, b = tap - a * p.kw;
                    tma_load_4d(raw_a(s) + j * 4096, &tmX, &bar_full[s], cc * 32, x0 + b - p.pad_l, y0 + a - p.pad_t, n0);
                }
                for (int j = 0; j < BNT / 32; ++j) {
                    const int co = co0 + j * 32;
                    if (p.dy_block2_f > 0) {
                        const int blk = co / p.dy_block2_f, f0 = co - blk * p.dy_block2_f;
                        tma_load_5d(raw_b(s) + j * 4096, &tmDY, &bar_full[s], f0, blk & 1, x0, blk >> 1, n0 * p.oh + y0);
                    } else {
                        tma_load_4d(raw_b(s) + j * 4096, &tmDY, &bar_full[s], co, x0, y0, n0);
                    }
                }
                TCP_ADD(33);
            }
        }
    } else if (warp == 1) {
        {   // ---- MMA issuer: the whole warp walks the loop (uniform operands), one elected lane issues (see tc::elect_one)
            constexpr uint32_t idesc = make_idesc_tf32(128, BNT, 0, 0);
            TCP_DECL
            for (int it = 0; it < iters; ++it) {
                const int s = it % kWgStages, ph = (it / kWgStages) & 1;
                TCP_START();
                mbar_wait(&bar_aready[s], ph, 1);
                TCP_ADD(34);
                mbar_wait(&bar_bready[s], ph, 2);
                TCP_ADD(35);
                tcgen05_fence_after();
                const uint32_t bh = smem_u32(b_hi(s)), bl = smem_u32(b_lo(s));
                const uint32_t ahi = a_base + s * 64, alo = ahi + 32;
                const uint64_t dbh0 = make_smem_desc_sw128(bh, 16, 1024), dbl0 = make_smem_desc_sw128(bl, 16, 1024);
                if (elect_one()) {
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint64_t dbh = dbh0 + (uint64_t)(ks * 2), dbl = dbl0 + (uint64_t)(ks * 2);
                        const uint32_t acc = (it > 0 || ks > 0) ? 1u : 0u;
                        umma_tf32_ts(tmem + BNT, alo + ks * 8, dbh, idesc, acc);
                        umma_tf32_ts(tmem + BNT, ahi + ks * 8, dbl, idesc, 1u);
                        umma_tf32_ts(tmem, ahi + ks * 8, dbh, idesc, acc);
                    }
                    umma_commit(&bar_free[s]);
                }
                __syncwarp();
                TCP_ADD(36);
            }
            if (elect_one()) umma_commit(&bar_accum);
            __syncwarp();
        }
    } else if (warp < 6) {
        // ---- A converters: warp q <-> atom q (rows 32q .. 32q+31 of the M tile), lane = channel
        // a warp may only touch TMEM lanes [32 (warp % 4), +32): warp with quarter q converts atom q (M rows 32q .. 32q+31)
        const int q = warp & 3;
#ifdef NI_TC_PROFILE
        long long tcp_t = 0; const bool tcp_on = blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 64;
#endif
        for (int it = 0; it < iters; ++it) {
            const int s = it % kWgStages, ph = (it / kWgStages) & 1;
            TCP_START();
            mbar_wait(&bar_full[s], ph, 3);
            TCP_ADD(37);
            float hi[32], lo[32];
            if (q < valid_atoms) {
                const uint32_t ap = smem_u32(raw_a(s)) + (uint32_t)(q * 4096 + (lane & 3) * 4);
#pragma unroll
                for (int px = 0; px < 32; ++px) {
                    // element (pixel px, channel lane): 16-byte chunk (lane / 4) XOR (px % 8), word lane % 4
                    const float v = lds32(ap + (uint32_t)(px * 128 + (((lane >> 2) ^ (px & 7)) << 4)));
                    const float h = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
                    hi[px] = h; lo[px] = v - h;
                }
            } else {
#pragma unroll
                for (int px = 0; px < 32; ++px) { hi[px] = 0.f; lo[px] = 0.f; }
            }
            // slot s was last read by the MMAs of iteration it - kWgStages, whose completion released bar_free[s] to the
            // producer before this stage was refilled, so the slot is free once bar_full[s] has fired
            TCP_ADD(38);
            tcgen05_fence_after();
            const uint32_t dst = a_base + ((uint32_t)(q * 32) << 16) + s * 64;
            tmem_st_32x32(dst, hi);
            tmem_st_32x32(dst + 32, lo);
            tmem_st_wait();
            tcgen05_fence_before();
            mbar_arrive(&bar_aready[s]);
            TCP_ADD(39);
        }
        mbar_wait(&bar_accum, 0, 4);
        tcgen05_fence_after();
        const int row = q * 32 + lane;
        const int mm = blockIdx.x * 128 + row;
        const bool valid = mm < p.mtot && iters > 0;
#pragma unroll 1
        for (int c = 0; c < BNT / 32; ++c) {
            float v[32], v2[32];
            tmem_ld_32x32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), v);
            tmem_ld_32x32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(BNT + c * 32), v2);
            if (!valid) continue;
            float* o = p.dw + (long long)mm * p.cout + co0 + c * 32;     // 16-byte aligned: cout % 32 == 0
#pragma unroll
            for (int jj = 0; jj < 32; jj += 4)      // 128-bit reductions: a quarter of the L2 atomic operations of scalar atomicAdd
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + jj), "f"(v[jj] + v2[jj]), "f"(v[jj + 1] + v2[jj + 1]),
                             "f"(v[jj + 2] + v2[jj + 2]), "f"(v[jj + 3] + v2[jj + 3]) : "memory");
        }
        tcgen05_fence_before();
    } else {
        // ---- B transposers (128 threads)
        const int tid = threadIdx.x - 192;
#ifdef NI_TC_PROFILE
        long long tcp_t = 0; const bool tcp_on = blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 192;
#endif
        for (int it = 0; it < iters; ++it) {
            const int s = it % kWgStages, ph = (it / kWgStages) & 1;
            TCP_START();
            mbar_wait(&bar_full[s], ph, 6);
            TCP_ADD(40);
            for (int qq = tid; qq < (BNT / 32) * 256; qq += 128)
                transpose_split_chunk(smem_u32(raw_b(s)) + (uint32_t)((qq >> 8) * 4096), smem_u32(b_hi(s)), smem_u32(b_lo(s)), (qq >> 8) * 32, qq & 255);
            fence_proxy_async_smem();
            mbar_arrive(&bar_bready[s]);
            TCP_ADD(41);
        }
    }
    __syncthreads();
    if (warp == 1) { tcgen05_fence_after(); tmem_dealloc(tmem, TMEM_COLS); }
}

}  // namespace tcv2
