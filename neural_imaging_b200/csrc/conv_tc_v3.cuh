// tcgen05 convolution gemm, generation 3: PERSISTENT, one CTA per SM, epilogue overlapped with the next tile's main loop.
//
// Measured on generation 2 (one CTA per output tile): per-tile time = T_fixed + iters * T_iter with T_fixed ~ 25 main-loop
// iterations (launch + TMEM allocation + first TMA round trip + accumulator drain + stores, none of it overlapped because a
// second CTA does not fit / blocks on the TMEM allocation), while the layers of this workload have only 9 - 100 iterations
// per tile (K = taps * cin / 32). Layers with 9 - 25 iterations ran at 40 - 100 TFLOP/s, the 100-iteration layer at 160.
// Here a CTA walks over tiles: barriers, TMEM and the tensor map are set up once; the producers run ahead into the next tile;
// dedicated epilogue warps drain one accumulator set while the MMA warp fills the other (N <= 64; N = 128 has TMEM room for
// one set only and just overlaps the stores).
//
// Warp roles (16 warps): 0 A producer (halo tiles, see conv_tc_v2.cuh), 1 MMA issuer, 2 B producer, 3 TMEM allocator + second MMA issuer,
// 4-11 converters (two per TMEM lane quarter, 16 K-columns each), 12-15 epilogue (one per lane quarter).
//
// Generation 3b (hardware probes: tools/hw_probes.py, profiles/r1_hw_probes.txt):
//  * a cp.async.bulk / TMA operation occupies the SM's copy engine for ~735 cycles whatever its size up to 32 KB (8, 16 and 32 KB
//    copies all complete one per ~735 cycles, at any depth in flight). With one weight copy per k-iteration the N = 32 / 64 main
//    loops (8 / 16 KB per copy) were bound by that, not by their MMAs (408 / 560 cycles per iteration): the weight stream now moves
//    in GROUPS of G = 4 / 2 / 1 consecutive k-iterations = one 32 KB copy (the prep kernel lays the tiles out [n tile][k-iteration]);
//  * the tensor pipe needs 20.5 + 0.42 N cycles per kind::tf32 MMA with A in tensor memory (34 / 47 / 74 cycles at N = 32 / 64 / 128)
//    and its queue is shallow: whatever the issuer spends between two iterations, the pipe idles. The issue loop is free of integer
//    divisions now (power-of-two rings, descriptors advanced by additions);
//  * when the whole weight slice of the launch fits in shared memory it is loaded once per CTA (resident) instead of once per tile;
//  * an A stage is handed back to the TMA producer only after the registers loaded from it have been consumed (the tcgen05.st that
//    reads them has been issued). The release used to follow the LDS *issue*; corrupted rows showed up in the second tile of a CTA
//    on cold launches (tools/tc_repro3.py).
#pragma once
#include "conv_desc.h"
#include "conv_tc_v2.cuh"
#include "ni_common.cuh"
#include "tc_common.cuh"

namespace tcv3 {
using namespace tc;
using tcv2::act_apply;
using tcv2::GemmParams;
using tcv2::pow2_cols;

constexpr int kThreads = 512;
constexpr int kNCW = 8;            // converter warps
constexpr int kMaxSA = 2;
constexpr int kMaxGroups = 32;     // weight-copy groups per tile in resident mode; ring stages (1, 2 or 4) when streaming
constexpr int kGroupBytes = 32768; // one weight copy = G k-iterations of [hi | lo] tiles

// N <= 64 is issue-bound with ONE issuer thread (~20 cycles of fixed cost per MMA against 34 / 47 of execution leave no slack for the
// waits), so those tiles get TWO issuer warps (different scheduler partitions) that take alternate k-iterations and own separate
// accumulators [D1_0, D2_0, D1_1, D2_1]; the epilogue adds them up (the sum is order-independent, only each accumulator's FIRST MMA
// has to overwrite, and that is a per-issuer property). Measured and dropped for N = 64: ONE issuer with two accumulator sets of
// (nacc + 1) accumulators and two A slots (epilogue overlapped with the next tile): 10 - 35 % slower than two issuers with one set.
template <int BNT, int SL = (BNT == 128 ? 2 : 4)> struct Cfg {
    static constexpr int ISSUERS = BNT == 128 ? 1 : 2;
    static constexpr int SLOTS = SL;                     // A (hi | lo) slots in tensor memory, 64 columns each (N = 128: 2, or 4 with one
                                                         // hi*hi accumulator, i.e. shallow contractions only)
    static constexpr int LOG_SLOTS = SL == 2 ? 1 : 2;
    static constexpr int NSETS = BNT == 32 ? 2 : 1;      // accumulator sets (512 TMEM columns: sets * set_cols + SLOTS * 64)
    static constexpr int B_BYTES = BNT * 128;            // one (BNT x 32) tf32 tile; a k-iteration reads hi + lo
    static constexpr int G = kGroupBytes / (2 * B_BYTES); // k-iterations per weight copy: 4 / 2 / 1
    static constexpr int LOG_G = BNT == 32 ? 2 : (BNT == 64 ? 1 : 0);
};

struct PersistParams {
    int mtiles, total_tiles;       // pixel tiles, pixel tiles * n tiles
    int sb, log_sb;                // weight ring: stages (1, 2 or 4 groups) and log2; resident: groups per tile (ring unused)
    int b_resident;                // 1: the CTA's whole weight slice is loaded ONCE and stays in shared memory (single n tile)
    int defer_st;                  // 1: converters complete their tcgen05.st one iteration later (behind the next loads + split)
    int experiment;                // development builds (-DNI_DEV): bit 0 converters skip LDS + split, bit 1 skip tcgen05.st, bit 2 issuers skip the MMAs, bit 3 skip the LDS only, bit 4 skip the split only, bit 5 epilogue skips its global stores
};

template <int BNT, int SL = (BNT == 128 ? 2 : 4)>
__global__ void __launch_bounds__(kThreads, 1)
conv_tc3_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const float* __restrict__ wtiled, const GemmParams p, const PersistParams q) {
    using C = Cfg<BNT, SL>;
    constexpr int SLOTS = C::SLOTS, LOG_SLOTS = C::LOG_SLOTS, NSETS = C::NSETS, B_BYTES = C::B_BYTES, ISSUERS = C::ISSUERS, G = C::G, LOG_G = C::LOG_G;
    const uint32_t set_cols = ISSUERS == 2 ? 4u * BNT : (uint32_t)(p.nacc + 1) * BNT;
    const uint32_t acc_cols = set_cols * NSETS;
    const uint32_t TMEM_COLS = pow2_cols(acc_cols + SLOTS * 64);
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar_afull[kMaxSA], bar_afree[kMaxSA], bar_bfull[kMaxGroups], bar_bfree[kMaxGroups], bar_tready[SLOTS], bar_tfree[SLOTS],
        bar_accfull[NSETS], bar_accfree[NSETS];
    __shared__ uint32_t tmem_slot;

    // lane-0 broadcasts: tell the compiler that the warp index and the TMEM base are warp-uniform (uniform registers, no waterfalls)
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int taps = p.kh * p.kw;
    const int iters = taps * p.kchunks;
    const int ngroups = (iters + G - 1) >> LOG_G;                   // weight copies per tile
    const int n_iss = (ISSUERS == 2 && iters >= 2) ? 2 : 1;          // active issuer warps
    const int nsum = ISSUERS == 2 ? 2 * n_iss : p.nacc + 1;         // accumulators the epilogue adds up

    if (threadIdx.x == 0) {
        for (int s = 0; s < kMaxSA; ++s) { mbar_init(&bar_afull[s], 1); mbar_init(&bar_afree[s], kNCW); }
        for (int s = 0; s < kMaxGroups; ++s) { mbar_init(&bar_bfull[s], 1); mbar_init(&bar_bfree[s], n_iss); }
        for (int t = 0; t < SLOTS; ++t) { mbar_init(&bar_tready[t], kNCW); mbar_init(&bar_tfree[t], 1); }
        for (int a = 0; a < NSETS; ++a) { mbar_init(&bar_accfull[a], n_iss); mbar_init(&bar_accfree[a], 4); }
        fence_barrier_init();
        tma_prefetch_desc(&tmA);
    }
    if (warp == 3) tmem_alloc(&tmem_slot, TMEM_COLS);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = __shfl_sync(0xffffffffu, tmem_slot, 0);
    const uint32_t a_base = tmem + acc_cols;

    auto a_halo = [&](int s) { return smem + s * p.a_stage; };
    uint8_t* const b_base = smem + p.sa * p.a_stage;                 // weight stages / resident groups, kGroupBytes apart
    // tile t -> (pixel tile m = t % mtiles, n tile = t / mtiles): neighbouring CTAs share the weight slice in L2
    auto tile_origin = [&](int tile, int& x0, int& y0, int& n0, int& nt) {
        const int m = tile % q.mtiles;
        nt = tile / q.mtiles;
        const int tw = m % p.tiles_w, th = (m / p.tiles_w) % p.tiles_h, tn = m / (p.tiles_w * p.tiles_h);
        x0 = tw * p.bw; y0 = th * p.bh; n0 = tn * p.bn;
    };

    if (warp == 0) {
        if (lane == 0) {   // ---- A producer: one halo box per (tile, 32-channel chunk)
            const uint32_t bytes = (uint32_t)(p.hw * p.hh * p.bn) * 128u;
            int s = 0, ph = 0;
            TCP_DECL
            for (int tile = blockIdx.x; tile < q.total_tiles; tile += gridDim.x) {
                int x0, y0, n0, nt;
                tile_origin(tile, x0, y0, n0, nt);
                const int bx = x0 + p.off_x0 + (p.off_sign < 0 ? -(p.kw - 1) : 0), by = y0 + p.off_y0 + (p.off_sign < 0 ? -(p.kh - 1) : 0);
                for (int kc = 0; kc < p.kchunks; ++kc) {
                    TCP_START();
                    mbar_wait(&bar_afree[s], ph ^ 1, 0);
                    TCP_ADD(12);
                    mbar_expect_tx(&bar_afull[s], bytes);
                    if (p.src_block2_f > 0) {          // 1x1 filter: halo == tile; channel chunk kc*32 lives in sub-pixel blk of the 2x2 block
                        const int co = kc * 32, blk = co / p.src_block2_f, f0 = co - blk * p.src_block2_f;
                        tma_load_5d(a_halo(s), &tmA, &bar_afull[s], f0, blk & 1, bx, blk >> 1, n0 * p.oh + by);
                    } else {
                        tma_load_4d(a_halo(s), &tmA, &bar_afull[s], kc * 32, bx, by, n0);
                    }
                    if (++s == p.sa) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 2 || (warp == 3 && ISSUERS == 1)) {
        // ---- B producer(s): pre-tiled [hi | lo] weight tiles of G consecutive k-iterations per bulk copy. A copy occupies its ISSUING
        // thread for ~735 - 870 cycles (tools/hw_probes.py), about one N = 128 iteration of MMAs, so where warp 3 is not an MMA issuer
        // (N = 128) it takes every other group.
        if (lane == 0) {
            const int nbp = ISSUERS == 1 ? 2 : 1, pb = warp == 2 ? 0 : 1;
            int gg = 0;
            TCP_DECL
            for (int tile = blockIdx.x; tile < q.total_tiles; tile += gridDim.x) {
                if (q.b_resident && tile != (int)blockIdx.x) break;      // resident weights: loaded while the first tile runs
                const int nt = tile / q.mtiles;
                const float* src = wtiled + (size_t)nt * iters * (size_t)(2 * BNT * 32);
                for (int g = 0; g < ngroups; ++g, ++gg) {
                    if (nbp == 2 && (gg & 1) != pb) continue;
                    const int s = q.b_resident ? g : (gg & (q.sb - 1)), ph = q.b_resident ? 0 : ((gg >> q.log_sb) & 1);
                    const int n_it = min(G, iters - (g << LOG_G));
                    TCP_START();
                    mbar_wait(&bar_bfree[s], ph ^ 1, 7);
                    TCP_ADD(13);
                    mbar_expect_tx(&bar_bfull[s], (uint32_t)n_it * 2u * B_BYTES);
                    bulk_load_1d(b_base + (size_t)s * kGroupBytes, src + (size_t)(g << LOG_G) * (size_t)(2 * BNT * 32), (uint32_t)n_it * 2u * B_BYTES,
                                 &bar_bfull[s]);
                }
            }
        }
    } else if (warp == 1 || (warp == 3 && ISSUERS == 2)) {
        const int iss = warp == 1 ? 0 : 1;
        if (iss < n_iss) {   // ---- MMA issuer(s): the whole warp walks the loop (uniform control flow), one elected lane issues
            constexpr uint32_t idesc = make_idesc_tf32(128, BNT, 0, 0);
            const uint64_t bdesc0 = make_smem_desc_sw128(smem_u32(b_base), 16, 1024);     // start-address field in 16-byte units
            const int tail_n = iters - ((ngroups - 1) << LOG_G);              // k-iterations in the last group of a tile
            int gbase = 0, ggbase = 0, tcount = 0;
            TCP_DECL
            for (int tile = blockIdx.x; tile < q.total_tiles; tile += gridDim.x, ++tcount, gbase += iters, ggbase += ngroups) {
                const int aset = NSETS == 2 ? (tcount & 1) : 0, use = NSETS == 2 ? (tcount >> 1) : tcount;
                TCP_START();
                mbar_wait(&bar_accfree[aset], (use & 1) ^ 1, 8);          // epilogue has drained this accumulator set
                TCP_ADD(3);
                tcgen05_fence_after();
                const uint32_t dbase = tmem + (uint32_t)aset * set_cols;
                int cur_g = -1, bs = 0, acc_i = 0;
                uint64_t gdesc = bdesc0;
                for (int it = iss; it < iters; it += n_iss) {
                    const int git = gbase + it;
                    const int t = git & (SLOTS - 1), pt = (git >> LOG_SLOTS) & 1;
                    const int g = it >> LOG_G;
                    TCP_START();
                    if (g != cur_g) {                                     // first k-iteration of this issuer in a new weight group
                        cur_g = g;
                        const int gg = ggbase + g;
                        bs = q.b_resident ? g : (gg & (q.sb - 1));
                        mbar_wait(&bar_bfull[bs], q.b_resident ? 0 : ((gg >> q.log_sb) & 1), 1);
                        gdesc = bdesc0 + (uint64_t)((uint32_t)bs * (uint32_t)(kGroupBytes >> 4));
                    }
                    TCP_ADD(1);
                    mbar_wait(&bar_tready[t], pt, 2);
                    TCP_ADD(2);
                    tcgen05_fence_after();
                    const uint32_t ahi = a_base + t * 64, alo = ahi + 32;
                    uint32_t d1, d2, first1, first2;
                    if (ISSUERS == 2) {
                        d1 = dbase + (uint32_t)(2 * iss) * BNT; d2 = d1 + BNT;
                        first1 = first2 = it == iss ? 1u : 0u;
                    } else {
                        d1 = dbase + (uint32_t)acc_i * BNT; d2 = dbase + (uint32_t)p.nacc * BNT;
                        first1 = it < p.nacc ? 1u : 0u; first2 = it == 0 ? 1u : 0u;
                        if (++acc_i == p.nacc) acc_i = 0;
                    }
                    const uint64_t dbh0 = gdesc + (uint64_t)((uint32_t)(it & (G - 1)) * (uint32_t)(2 * B_BYTES >> 4)), dbl0 = dbh0 + (uint64_t)(B_BYTES >> 4);
                    // last k-iteration of THIS issuer inside the weight group: its commit hands the stage back (bfree counts n_iss arrivals)
                    const bool group_done = !q.b_resident && (it + n_iss >= min((g + 1) << LOG_G, iters));
                    if (elect_one()) {
#ifdef NI_DEV
                        if (!(q.experiment & 4))
#endif
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            const uint64_t dbh = dbh0 + (uint64_t)(ks * 2), dbl = dbl0 + (uint64_t)(ks * 2);   // +32 bytes
                            umma_tf32_ts(d2, alo + ks * 8, dbh, idesc, (first2 && ks == 0) ? 0u : 1u);
                            umma_tf32_ts(d2, ahi + ks * 8, dbl, idesc, 1u);
                            umma_tf32_ts(d1, ahi + ks * 8, dbh, idesc, (first1 && ks == 0) ? 0u : 1u);
                        }
                        TCP_ADD(17);
                        umma_commit(&bar_tfree[t]);
                        if (group_done) umma_commit(&bar_bfree[bs]);
                        TCP_ADD(4);
                    }
                    __syncwarp();
                }
                // a last group holding a single k-iteration is read by one issuer only: the other one still owes the stage its arrival
                if (!q.b_resident && n_iss == 2 && tail_n == 1 && ((iters - 1) & 1) != iss) {
                    const int gg = ggbase + ngroups - 1;
                    if (elect_one()) mbar_arrive(&bar_bfree[gg & (q.sb - 1)]);
                    __syncwarp();
                }
                if (elect_one()) umma_commit(&bar_accfull[aset]);
                __syncwarp();
            }
        }
    } else if (warp >= 4 && warp < 4 + kNCW) {
        // ---- converters: A halo row -> registers -> hi / lo -> TMEM (two warps per lane quarter, 16 K-columns each)
        const int qd = warp & 3, row = qd * 32 + lane, half = (warp - 4) >> 2;
        const int prow0 = ((row / (p.bw * p.bh)) * p.hh + (row / p.bw) % p.bh) * p.hw + row % p.bw;
        int git = 0, s = 0, ph = 0;
        int pending = -1;            // slot whose tcgen05.st is in flight: its completion wait + "ready" arrive are deferred until
                                     // the next iteration's shared-memory loads and hi/lo split are done (hides ~200 cycles)
#ifdef NI_TC_PROFILE
        long long tcp_t = 0; const bool tcp_on = blockIdx.x == 0 && threadIdx.x == 128;
#endif
        for (int tile = blockIdx.x; tile < q.total_tiles; tile += gridDim.x) {
            for (int kc = 0; kc < p.kchunks; ++kc) {
                TCP_START();
                if (!__all_sync(0xffffffffu, mbar_try_wait(&bar_afull[s], ph))) {      // warp-uniform: the arrivals below are one per WARP
                    // the halo has not landed yet: do not keep the MMA warp waiting for the previous iteration's operand meanwhile
                    if (pending >= 0) {
                        tmem_st_wait();
                        tcgen05_fence_before();
                        warp_arrive(&bar_tready[pending], lane);
                        pending = -1;
                    }
                    mbar_wait(&bar_afull[s], ph, 3);
                }
                TCP_ADD(6);
                const uint32_t stage = smem_u32(a_halo(s));
                int ta = 0, tb = 0;
                for (int tap = 0; tap < taps; ++tap, ++git) {
                    const int t = git & (SLOTS - 1), pt = (git >> LOG_SLOTS) & 1;
                    const int prow = prow0 + (p.off_sign > 0 ? ta : p.kh - 1 - ta) * p.hw + (p.off_sign > 0 ? tb : p.kw - 1 - tb);
                    if (++tb == p.kw) { tb = 0; ++ta; }
                    float hi[16], lo[16];
                    const uint32_t rp = stage + (uint32_t)prow * 128u;
#ifdef NI_DEV
                    if (q.experiment & 1) {
#pragma unroll
                        for (int e = 0; e < 16; ++e) { hi[e] = 1.f; lo[e] = 0.f; }
                    } else
#endif
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
#ifdef NI_DEV
                        const float4 v = (q.experiment & 8) ? make_float4((float)prow, (float)c, (float)tap, 1.f)
                                                            : lds128(rp + (uint32_t)(((4 * half + c) ^ (prow & 7)) << 4));
#else
                        const float4 v = lds128(rp + (uint32_t)(((4 * half + c) ^ (prow & 7)) << 4));
#endif
                        const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
#ifdef NI_DEV
                            if (q.experiment & 16) { hi[4 * c + e] = vv[e]; lo[4 * c + e] = vv[e]; continue; }
#endif
                            const float h = __uint_as_float(__float_as_uint(vv[e]) & 0xFFFFE000u);
                            hi[4 * c + e] = h;
                            lo[4 * c + e] = vv[e] - h;
                        }
                    }
                    TCP_ADD(7);
                    if (pending >= 0) {
                        tmem_st_wait();
                        tcgen05_fence_before();
                        warp_arrive(&bar_tready[pending], lane);
                    }
                    TCP_ADD(9);
                    mbar_wait(&bar_tfree[t], pt ^ 1, 5);
                    TCP_ADD(8);
                    tcgen05_fence_after();
                    const uint32_t dst = a_base + ((uint32_t)(qd * 32) << 16) + t * 64 + 16 * half;
#ifdef NI_DEV
                    if (!(q.experiment & 2))
#endif
                    {
                        tmem_st_32x16(dst, hi);
                        tmem_st_32x16(dst + 32, lo);
                    }
                    pending = t;
                    if (!q.defer_st) {
                        tmem_st_wait();
                        tcgen05_fence_before();
                        warp_arrive(&bar_tready[pending], lane);
                        pending = -1;
                    }
                    // Stage back to the TMA producer only now: the tcgen05.st above consumed the registers of the last tap, so every
                    // shared-memory load of this thread from the stage has completed (an arrive right behind the LDS *issue* is not
                    // ordered after the loads' data phase)
                    if (tap == taps - 1) warp_arrive(&bar_afree[s], lane);
                }
                if (++s == p.sa) { s = 0; ph ^= 1; }
            }
        }
        if (pending >= 0) {
            tmem_st_wait();
            tcgen05_fence_before();
            warp_arrive(&bar_tready[pending], lane);
        }
    } else if (warp >= 12) {
        // ---- epilogue: accumulator set -> (+ bias, activation) -> global, while the MMA warp works on the other set
        const int qd = warp & 3, row = qd * 32 + lane;
        const int lw = row % p.bw, lh = (row / p.bw) % p.bh, ln = row / (p.bw * p.bh);
        int tcount = 0;
        // fused activation backward: per-lane partial bias gradients (lane = channel within a 32-channel chunk), kept in registers across
        // the tiles of this persistent CTA and flushed with one atomicAdd per channel when the n-tile changes / at the end
        const bool fuse_act = p.dact_y != nullptr || p.dbias != nullptr;
        float bs0 = 0.f, bs1 = 0.f, bs2 = 0.f, bs3 = 0.f;
        int bs_nt = -1;
        auto flush_bias = [&]() {
            if (p.dbias == nullptr || bs_nt < 0) return;
            const float bsv[4] = {bs0, bs1, bs2, bs3};
#pragma unroll
            for (int cc = 0; cc < BNT / 32; ++cc) {
                const int co = bs_nt * BNT + cc * 32 + lane;
                atomicAdd(p.dbias + (p.dbias_mod > 0 ? co % p.dbias_mod : co), bsv[cc]);
            }
            bs0 = bs1 = bs2 = bs3 = 0.f;
        };
#ifdef NI_TC_PROFILE
        long long tcp_t = 0; const bool tcp_on = blockIdx.x == 0 && threadIdx.x == 384;
#endif
        for (int tile = blockIdx.x; tile < q.total_tiles; tile += gridDim.x, ++tcount) {
            const int aset = NSETS == 2 ? (tcount & 1) : 0, use = NSETS == 2 ? (tcount >> 1) : tcount;
            int x0, y0, n0, nt;
            tile_origin(tile, x0, y0, n0, nt);
            const int ox = x0 + lw, oy = y0 + lh, on = n0 + ln;
            const bool valid = on < p.n && oy < p.oh && ox < p.ow;
            if (fuse_act && nt != bs_nt) { flush_bias(); bs_nt = nt; }
            // fused activation backward: the forward output of the layer below is fetched BEFORE waiting for the accumulators (chunk 0)
            // and one chunk ahead afterwards, so that its HBM latency hides behind the main loop / the previous chunk's work
            const bool want_y = fuse_act && p.dact_y != nullptr && valid && p.dact != NI_ACT_NONE;
            const float4* y4base = want_y ? reinterpret_cast<const float4*>(
                p.dact_y + (((long long)on * p.oh + oy) * p.ow + ox) * p.dact_pitch + p.dact_coff + nt * BNT) : nullptr;
            float4 ynext[8];
            if (want_y) {
#pragma unroll
                for (int j = 0; j < 8; ++j) ynext[j] = y4base[j];
            }
            TCP_START();
            mbar_wait(&bar_accfull[aset], use & 1, 4);
            TCP_ADD(10);
            tcgen05_fence_after();
            const uint32_t dbase = tmem + (uint32_t)aset * set_cols + ((uint32_t)(qd * 32) << 16);
#pragma unroll 1
            for (int c = 0; c < BNT / 32; ++c) {
                float v[32], v2[32];
                tmem_ld_32x32(dbase + (uint32_t)(c * 32), v);
                for (int a2 = 1; a2 < nsum; ++a2) {
                    tmem_ld_32x32(dbase + (uint32_t)(a2 * BNT + c * 32), v2);
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] += v2[j];
                }
                TCP_ADD(14);
                if (c == BNT / 32 - 1) {          // every TMEM read of this tile has completed: release the accumulator set
                    tcgen05_fence_before();
                    warp_arrive(&bar_accfree[aset], lane);
                }
                if (!valid && !fuse_act) continue;
                const int co0 = nt * BNT + c * 32;
                if (fuse_act) {
                    // v <- v * act'(y) with y = the forward output of the layer whose output gradient this is (same pixel, same channels)
                    if (want_y) {
                        float4 yv[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) yv[j] = ynext[j];
                        if (c + 1 < BNT / 32) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) ynext[j] = y4base[(c + 1) * 8 + j];
                        }
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float ye[4] = {yv[j].x, yv[j].y, yv[j].z, yv[j].w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float yy = ye[e];
                                float dd;
                                switch (p.dact) {
                                    case NI_ACT_LEAKY_RELU: dd = yy > 0.f ? 1.f : p.dact_alpha; break;
                                    case NI_ACT_RELU: dd = yy > 0.f ? 1.f : 0.f; break;
                                    case NI_ACT_TANH: dd = 1.f - yy * yy; break;
                                    case NI_ACT_SIGMOID: dd = yy * (1.f - yy); break;
                                    default: dd = 1.f; break;
                                }
                                v[4 * j + e] *= dd;
                            }
                        }
                    }
                    if (p.dbias != nullptr) {
                        // column sums over the 32 pixels of this warp: butterfly, lane l ends up with the sum of channel co0 + l
                        float r[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j) r[j] = valid ? v[j] : 0.f;
#pragma unroll
                        for (int off = 16; off >= 1; off >>= 1) {
                            const bool upper = (lane & off) != 0;
#pragma unroll
                            for (int j = 0; j < off; ++j) {
                                const float send = upper ? r[j] : r[j + off];
                                const float keep = upper ? r[j + off] : r[j];
                                r[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                            }
                        }
                        if (c == 0) bs0 += r[0]; else if (c == 1) bs1 += r[0]; else if (c == 2) bs2 += r[0]; else bs3 += r[0];
                    }
                    if (!valid) continue;
                }
                if (p.n_valid < p.ntot) {
                    // zero-padded n tile (U-Net output layer, 12 of 32 columns): bias + activation on the real channels only
                    const int F = p.n_valid >> 2;
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (j < p.n_valid) v[j] = act_apply(v[j] + (p.bias ? __ldg(p.bias + j) : 0.f), p.act, p.alpha);
                    if (p.out_mode != NI_MODE_PLAIN && F == 3 && p.out_pitch == 3 && p.out_coff == 0 && !p.accumulate) {
                        // depth_to_space(2) of 12 channels into an RGB image: the two sub-pixels of a row are 6 contiguous floats (8-byte aligned)
#pragma unroll
                        for (int r = 0; r < 2; ++r) {
                            float2* o2 = reinterpret_cast<float2*>(p.out + (((long long)on * 2 * p.oh + 2 * oy + r) * (2 * p.ow) + 2 * ox) * 3);
                            o2[0] = make_float2(v[6 * r], v[6 * r + 1]);
                            o2[1] = make_float2(v[6 * r + 2], v[6 * r + 3]);
                            o2[2] = make_float2(v[6 * r + 4], v[6 * r + 5]);
                        }
                        continue;
                    }
                    if (p.out_mode == NI_MODE_PLAIN && !(p.out_pitch & 3) && !(p.out_coff & 3) && !p.accumulate) {
                        float4* o4n = reinterpret_cast<float4*>(p.out + (((long long)on * p.oh + oy) * p.ow + ox) * p.out_pitch + p.out_coff);
#pragma unroll
                        for (int j4 = 0; j4 < 8; ++j4)
                            if (4 * j4 < p.n_valid) o4n[j4] = make_float4(v[4 * j4], v[4 * j4 + 1], v[4 * j4 + 2], v[4 * j4 + 3]);
                        continue;
                    }
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        if (j >= p.n_valid) continue;
                        float* oj;
                        if (p.out_mode == NI_MODE_PLAIN) oj = p.out + (((long long)on * p.oh + oy) * p.ow + ox) * p.out_pitch + p.out_coff + j;
                        else {
                            const int blk = j / F, f = j - blk * F;
                            oj = p.out + (((long long)on * 2 * p.oh + 2 * oy + (blk >> 1)) * (2 * p.ow) + 2 * ox + (blk & 1)) * p.out_pitch + p.out_coff + f;
                        }
                        *oj = p.accumulate ? *oj + v[j] : v[j];
                    }
                    continue;
                }
                float* o;
                if (p.out_mode == NI_MODE_PLAIN) {
                    o = p.out + (((long long)on * p.oh + oy) * p.ow + ox) * p.out_pitch + p.out_coff + co0;
                } else {
                    const int F = p.ntot >> 2, blk = co0 / F, f0 = co0 - blk * F;
                    o = p.out + (((long long)on * 2 * p.oh + 2 * oy + (blk >> 1)) * (2 * p.ow) + 2 * ox + (blk & 1)) * p.out_pitch + p.out_coff + f0;
                }
                // bias: 8 independent 128-bit loads issued together (a scalar __ldg per element inside the activation switch
                // serialised 32 L2 round trips per chunk: 12.5 k cycles of the 13 k-cycle epilogue, in-kernel clock64 spans)
                if (p.bias) {
                    const float4* b4 = reinterpret_cast<const float4*>(p.bias + (p.bias_mod > 0 ? co0 % p.bias_mod : co0));
                    float4 bv[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) bv[j] = __ldg(b4 + j);
#pragma unroll
                    for (int j = 0; j < 8; ++j) { v[4 * j] += bv[j].x; v[4 * j + 1] += bv[j].y; v[4 * j + 2] += bv[j].z; v[4 * j + 3] += bv[j].w; }
                }
                switch (p.act) {       // one branch per chunk, straight-line code inside
                    case NI_ACT_LEAKY_RELU:
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = v[j] > 0.f ? v[j] : p.alpha * v[j];
                        break;
                    case NI_ACT_RELU:
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
                        break;
                    case NI_ACT_NONE: break;
                    default:
#pragma unroll 4
                        for (int j = 0; j < 32; ++j) v[j] = act_apply(v[j], p.act, p.alpha);
                        break;
                }
                TCP_ADD(15);
                float4* o4 = reinterpret_cast<float4*>(o);
#ifdef NI_DEV
                if (q.experiment & 32) continue;
#endif
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float4 w4 = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    if (p.accumulate) { const float4 old = o4[j]; w4.x += old.x; w4.y += old.y; w4.z += old.z; w4.w += old.w; }
                    o4[j] = w4;
                }
                TCP_ADD(16);
            }
            TCP_ADD(11);
        }
        if (fuse_act) flush_bias();
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 3) { tcgen05_fence_after(); tmem_dealloc(tmem, TMEM_COLS); }
}

}  // namespace tcv3
