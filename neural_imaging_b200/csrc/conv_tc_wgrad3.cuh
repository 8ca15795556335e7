// tcgen05 weight-gradient kernel, generation 3: ONE TMA PRODUCER WARP PER BOX.
//
// dW[(tap,ci), co] = sum_pixels X_tap[pixel, ci] * dY[pixel, co]: M = (tap, ci) flattened in 32-row atoms (4 per CTA), N = co, K = pixels,
// 32 per step; split-K over pixel ranges with 128-bit L2 reductions (see conv_tc_v2.cuh for the operand paths: X^T reaches tensor
// memory through the register -> TMEM-lane mapping of tcgen05.st, dY^T is turned K-major in shared memory by transposer warps).
//
// What generation 2 was bound by (in-kernel clock64 spans, tools/tc_prof.py): its single producer thread spent 80 % of the CTA's life
// ISSUING the 5 - 8 tensor boxes of a step (~365 cycles each), the converters / transposers / MMA warp waited for data 60 % of theirs.
// tools/hw_probes.py then showed that the cost of a copy-engine operation (~735 - 870 cycles per cp.async.bulk, whatever its size up
// to 64 KB and however many are in flight) is paid PER ISSUING THREAD: four warps issuing concurrently move four times the bytes
// (9.4 -> 37.6 B/clk/SM for 8 KB copies). So every box of a step has its own producer warp (4 X atoms + N/32 dY boxes), one CTA per SM
// with a deeper ring instead of two co-resident CTAs with one producer each, eight X converter warps (two per TMEM lane quarter, 16
// pixels each, tcgen05.st completion deferred behind the next step's loads) and eight transposer warps so that the conversion keeps up.
#pragma once
#include "conv_desc.h"
#include "conv_tc_v2.cuh"
#include "ni_common.cuh"
#include "tc_common.cuh"

namespace tcw3 {
using namespace tc;
using tcv2::kAraw;
using tcv2::pow2_cols;
using tcv2::transpose_split_chunk;
using tcv2::WgradParams;

template <int BNT, int NP_> struct Cfg {
    static constexpr int STAGES = BNT == 128 ? 3 : 4;
    static constexpr int NT = 8;                                        // transposer warps
    static constexpr int NP = NP_;                                      // producer warps: the 4 + N/32 boxes of a step are dealt round-robin
    // warps: 0-7 X converters (+ epilogue), 8 MMA issuer / TMEM allocator, 9 .. 9+NP-1 producers, then NT transposers
    static constexpr int THREADS = (8 + 1 + NP + NT) * 32;
    static constexpr int B_BYTES = BNT * 128;
    static constexpr int STAGE_BYTES = kAraw + 3 * B_BYTES;             // raw X (4 atoms), raw dY, dY^T hi, dY^T lo
};

// TMEM map: D1 [0,BNT), D2 [BNT, 2 BNT), A slots STAGES x 64 columns (hi 32 | lo 32).
template <int BNT, int NP_>
__global__ void __launch_bounds__(Cfg<BNT, NP_>::THREADS, 1)
conv_tc3_wgrad_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY, const WgradParams p) {
    using C = Cfg<BNT, NP_>;
    constexpr int STAGES = C::STAGES, NT = C::NT, NP = C::NP, B_BYTES = C::B_BYTES, STAGE_BYTES = C::STAGE_BYTES;
    constexpr int NB = BNT / 32;                                        // dY boxes per step
    const uint32_t TMEM_COLS = pow2_cols(2 * BNT + STAGES * 64);
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar_full[STAGES], bar_aready[STAGES], bar_bready[STAGES], bar_free[STAGES], bar_accum;
    __shared__ uint32_t tmem_slot;

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int atom0 = blockIdx.x * 4;
    const int valid_atoms = min(4, p.atoms - atom0);
    const int co0 = blockIdx.y * BNT;
    const int step0 = blockIdx.z * p.steps_per_split;
    const int iters = min(p.steps_per_split, p.steps_total - step0);

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&bar_full[s], NP); mbar_init(&bar_aready[s], 8); mbar_init(&bar_bready[s], NT); mbar_init(&bar_free[s], 1);      // "ready" arrivals are one per WARP (warp_arrive)
        }
        mbar_init(&bar_accum, 1);
        fence_barrier_init();
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmDY);
    }
    if (warp == 8) tmem_alloc(&tmem_slot, TMEM_COLS);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = __shfl_sync(0xffffffffu, tmem_slot, 0);
    const uint32_t a_base = tmem + 2 * BNT;

    auto raw_a = [&](int s) { return smem + s * STAGE_BYTES; };
    auto raw_b = [&](int s) { return smem + s * STAGE_BYTES + kAraw; };
    auto b_hi = [&](int s) { return smem + s * STAGE_BYTES + kAraw + B_BYTES; };
    auto b_lo = [&](int s) { return smem + s * STAGE_BYTES + kAraw + 2 * B_BYTES; };

    if (warp >= 9 && warp < 9 + NP) {
        if (lane == 0) {   // ---- producer j: boxes j, j + NP, ... of every step (box b < 4: X atom b, else dY box b - 4)
            const int j = warp - 9;
            constexpr int MAXB = (4 + NB + NP - 1) / NP;          // boxes per producer (1 or 2)
            // per-box constants hoisted out of the step loop (integer divisions on the issue path delay every box of the step)
            int kind[MAXB], c0[MAXB], dx[MAXB], dy[MAXB], c1[MAXB], c3[MAXB];
            uint32_t dst_off[MAXB];
            uint32_t bytes = 0;
#pragma unroll
            for (int i = 0; i < MAXB; ++i) {
                const int b = j + i * NP;
                kind[i] = 0; c0[i] = dx[i] = dy[i] = c1[i] = c3[i] = 0; dst_off[i] = 0;
                if (b < 4) {
                    if (b < valid_atoms) {
                        const int atom = atom0 + b, tap = atom / p.cin_chunks, cc = atom - tap * p.cin_chunks;
                        const int ta = tap / p.kw, tb = tap - ta * p.kw;
                        kind[i] = 1; c0[i] = cc * 32; dx[i] = tb - p.pad_l; dy[i] = ta - p.pad_t; dst_off[i] = (uint32_t)(b * 4096);
                        bytes += 4096u;
                    }
                } else if (b < 4 + NB) {
                    const int co = co0 + (b - 4) * 32;
                    dst_off[i] = (uint32_t)(kAraw + (b - 4) * 4096);
                    bytes += 4096u;
                    if (p.dy_block2_f > 0) {
                        const int blk = co / p.dy_block2_f;
                        kind[i] = 3; c0[i] = co - blk * p.dy_block2_f; c1[i] = blk & 1; c3[i] = blk >> 1;
                    } else {
                        kind[i] = 2; c0[i] = co;
                    }
                }
            }
            int s = 0, ph = 0;
            int tw = step0 % p.tiles_w, th = (step0 / p.tiles_w) % p.tiles_h, tn = step0 / (p.tiles_w * p.tiles_h);
#ifdef NI_TC_PROFILE
            long long tcp_t = 0; const bool tcp_on = blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && j == 0;
#endif
            for (int it = 0; it < iters; ++it) {
                TCP_START();
                mbar_wait(&bar_free[s], ph ^ 1, 0);
                TCP_ADD(32);
                mbar_expect_tx(&bar_full[s], bytes);          // arrive (1 of NP) + this producer's byte count
                const int x0 = tw * p.bw, y0 = th * p.bh, n0 = tn * p.bn;
                uint8_t* stage = raw_a(s);
#pragma unroll
                for (int i = 0; i < MAXB; ++i) {
                    if (kind[i] == 1) tma_load_4d(stage + dst_off[i], &tmX, &bar_full[s], c0[i], x0 + dx[i], y0 + dy[i], n0);
                    else if (kind[i] == 2) tma_load_4d(stage + dst_off[i], &tmDY, &bar_full[s], c0[i], x0, y0, n0);
                    else if (kind[i] == 3) tma_load_5d(stage + dst_off[i], &tmDY, &bar_full[s], c0[i], c1[i], x0, c3[i], n0 * p.oh + y0);
                }
                TCP_ADD(33);
                if (++s == STAGES) { s = 0; ph ^= 1; }
                if (++tw == p.tiles_w) { tw = 0; if (++th == p.tiles_h) { th = 0; ++tn; } }
            }
        }
    } else if (warp == 8) {
        {   // ---- MMA issuer: the whole warp walks the loop (uniform operands), one elected lane issues (see tc::elect_one)
            constexpr uint32_t idesc = make_idesc_tf32(128, BNT, 0, 0);
            const uint64_t bdesc0 = make_smem_desc_sw128(smem_u32(b_hi(0)), 16, 1024);
            int s = 0, ph = 0;
            TCP_DECL
            for (int it = 0; it < iters; ++it) {
                TCP_START();
                mbar_wait(&bar_aready[s], ph, 1);
                TCP_ADD(34);
                mbar_wait(&bar_bready[s], ph, 2);
                TCP_ADD(35);
                tcgen05_fence_after();
                const uint32_t ahi = a_base + s * 64, alo = ahi + 32;
                const uint64_t dbh0 = bdesc0 + (uint64_t)((uint32_t)s * (uint32_t)(STAGE_BYTES >> 4)), dbl0 = dbh0 + (uint64_t)(B_BYTES >> 4);
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
                if (++s == STAGES) { s = 0; ph ^= 1; }
            }
            if (elect_one()) umma_commit(&bar_accum);
            __syncwarp();
        }
    } else if (warp < 8) {
        // ---- X converters: warps (quarter q = warp % 4, half = warp / 4): atom q, lane = channel, pixels [16 half, 16 half + 16)
        // (a warp may only touch TMEM lanes [32 (warp % 4), +32))
        const int q = warp & 3, half = warp >> 2;
#ifdef NI_TC_PROFILE
        long long tcp_t = 0; const bool tcp_on = blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0;
#endif
        int s = 0, ph = 0;
        int pending = -1;       // stage whose tcgen05.st is in flight: completion wait + "ready" arrive deferred behind the next step's loads
        for (int it = 0; it < iters; ++it) {
            TCP_START();
            if (!__all_sync(0xffffffffu, mbar_try_wait(&bar_full[s], ph))) {       // warp-uniform (one arrival per warp below)
                if (pending >= 0) {           // nothing to overlap with: do not keep the MMA warp waiting
                    tmem_st_wait();
                    tcgen05_fence_before();
                    warp_arrive(&bar_aready[pending], lane);
                    pending = -1;
                }
                mbar_wait(&bar_full[s], ph, 3);
            }
            TCP_ADD(37);
            float hi[16], lo[16];
            if (q < valid_atoms) {
                const uint32_t ap = smem_u32(raw_a(s)) + (uint32_t)(q * 4096 + (lane & 3) * 4);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int px = 16 * half + i;
                    // element (pixel px, channel lane): 16-byte chunk (lane / 4) XOR (px % 8), word lane % 4
                    const float v = lds32(ap + (uint32_t)(px * 128 + (((lane >> 2) ^ (px & 7)) << 4)));
                    const float h = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
                    hi[i] = h; lo[i] = v - h;
                }
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) { hi[i] = 0.f; lo[i] = 0.f; }
            }
            TCP_ADD(38);
            if (pending >= 0) {
                tmem_st_wait();
                tcgen05_fence_before();
                warp_arrive(&bar_aready[pending], lane);
            }
            // slot s was last read by the MMAs of iteration it - STAGES, whose completion released bar_free[s] to the producers before
            // this stage was refilled, so the slot is free once bar_full[s] has fired
            tcgen05_fence_after();
            const uint32_t dst = a_base + ((uint32_t)(q * 32) << 16) + s * 64 + 16 * half;
            tmem_st_32x16(dst, hi);
            tmem_st_32x16(dst + 32, lo);
            pending = s;
            if (!p.defer_st) {
                tmem_st_wait();
                tcgen05_fence_before();
                warp_arrive(&bar_aready[pending], lane);
                pending = -1;
            }
            TCP_ADD(39);
            if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        if (pending >= 0) {
            tmem_st_wait();
            tcgen05_fence_before();
            warp_arrive(&bar_aready[pending], lane);
        }
        mbar_wait(&bar_accum, 0, 4);
        tcgen05_fence_after();
        const int row = q * 32 + lane;
        const int mm = blockIdx.x * 128 + row;
        const bool valid = mm < p.mtot && iters > 0;
#pragma unroll 1
        for (int c = half; c < BNT / 32; c += 2) {        // the two warps of a lane quarter interleave the 32-column chunks
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
        // ---- dY transposers (NT warps)
        const int tid = threadIdx.x - (9 + NP) * 32;
#ifdef NI_TC_PROFILE
        long long tcp_t = 0; const bool tcp_on = blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && tid == 0;
#endif
        int s = 0, ph = 0;
        for (int it = 0; it < iters; ++it) {
            TCP_START();
            mbar_wait(&bar_full[s], ph, 6);
            TCP_ADD(40);
            for (int qq = tid; qq < NB * 256; qq += NT * 32)
                transpose_split_chunk(smem_u32(raw_b(s)) + (uint32_t)((qq >> 8) * 4096), smem_u32(b_hi(s)), smem_u32(b_lo(s)), (qq >> 8) * 32, qq & 255);
            fence_proxy_async_smem();
            warp_arrive(&bar_bready[s], lane);
            TCP_ADD(41);
            if (++s == STAGES) { s = 0; ph ^= 1; }
        }
    }
    __syncthreads();
    if (warp == 8) { tcgen05_fence_after(); tmem_dealloc(tmem, TMEM_COLS); }
}

}  // namespace tcw3
