// tcgen05 implicit-GEMM convolutions for sm_100a (fprop, dgrad, wgrad), FP32-accurate through 3xTF32 — host side.
// Device kernels: conv_tc_v2.cuh (A operand through tensor memory).
//
// Why 3xTF32: BASELINE.json asks for 1e-5 relative parity with the reference's FP32 graphs; a single TF32 pass is ~3e-4.
// Each operand is split x = hi + lo (hi = x with the low 13 mantissa bits cleared = exactly what the tensor core's
// truncating FP32->TF32 conversion keeps; lo = x - hi, exact in FP32) and D = hi*hi + hi*lo + lo*hi is accumulated in
// FP32 in TMEM: the small correction terms in their own accumulator (D2), the hi*hi products in up to two rotated
// accumulators (the in-TMEM accumulation truncates addends at the accumulator's exponent; measured 1.2e-5 bias at
// K = 4608 with a single accumulator). Weights are split once per call by a prep kernel; activations are split tile by
// tile on their way from shared to tensor memory, so HBM / L2 traffic is that of a plain FP32 convolution.
//
// Implicit GEMM without im2col: for filter tap (a, b) the A tile (128 output pixels x 32 input channels) is ONE TMA box
// of the NHWC activation tensor shifted by (a - pad, b - pad); out-of-bounds box elements are zero-filled by the TMA
// unit, which IS TensorFlow's SAME zero padding. dgrad is the same kernel with the box shifted by (pad - a, pad - b) and
// the HWIO weights used as they are (K = cout is already contiguous). wgrad contracts over pixels: M = flattened
// (tap, ci) in 32-row atoms (fills the 128-row tile even for 32-channel layers), split-K over pixel ranges.
#include <cuda.h>

#include <stdlib.h>

#include <mutex>

#include "conv_desc.h"
#include "conv_tc_v2.cuh"
#include "conv_tc_v3.cuh"
#include "conv_tc_wgrad3.cuh"
#include "ni_common.cuh"
#include "tc_common.cuh"

// Tuning switches read from the environment exist in DEVELOPMENT builds only (NI_BUILD_TAG=dev -> -DNI_DEV, libni_b200_dev.so);
// the shipping library has one fixed configuration.
static inline const char* dev_env(const char* name) {
#ifdef NI_DEV
    return getenv(name);
#else
    (void)name;
    return nullptr;
#endif
}

namespace {

// w (taps, cin, cout) HWIO -> tiles for the gemm kernel: block ((nt * K/32 + kc) * taps + tap) = [hi | lo], each a
// (bnt rows x 32 k) K-major SWIZZLE_128B tile exactly as the MMA reads it. transpose: N = cout, K = cin (fprop); else N = cin, K = cout.
// kp = K rounded up to 32: layers with fewer than 32 channels on the contraction or the output side (U-Net 32 -> 12) run on padded tiles
// whose padding the launcher zero-fills; the activation side is padded by the tensor map's out-of-bounds fill.
__global__ void tc_prep_weights_kernel(const float* __restrict__ w, float* __restrict__ out, int taps, int cin, int cout, int transpose, int bnt, int kp) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    const long long total = (long long)taps * cin * cout;
    if (i >= total) return;
    const int N = transpose ? cout : cin, K = transpose ? cin : cout;
    const int k = (int)(i % K);
    const long long t = i / K;
    const int n = (int)(t % N), tap = (int)(t / N);
    const float v = transpose ? w[((long long)tap * cin + k) * cout + n] : w[((long long)tap * cin + n) * cout + k];
    const int kc = k >> 5, kl = k & 31, nt = n / bnt, nl = n - nt * bnt;
    const long long block = ((long long)nt * (kp >> 5) + kc) * taps + tap;    // [n tile][k-iteration = kc * taps + tap]: a tile's stream is contiguous
    const int off = (nl >> 3) * 256 + (nl & 7) * 32 + ((((kl >> 2) ^ (nl & 7)) << 2) + (kl & 3));
    const float h = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    float* o = out + block * (2 * bnt * 32);
    o[off] = h;
    o[bnt * 32 + off] = v - h;
}

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

}  // namespace

// cuTensorMapEncodeTiled is fetched through the runtime (cudaGetDriverEntryPoint) so that the library has no link-time
// dependency on libcuda.so.1: it must still load (and export its symbols) on a build box without a GPU driver.
// swizzle: 0 = CU_TENSOR_MAP_SWIZZLE_NONE (dense rows: the dJPEG pixel tiles), 1 = 128-byte swizzle (the convolution operands)
int ni_encode_tiled_sw(CUtensorMap* tm, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes, const cuuint32_t* box,
                       int swizzle) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        NI_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres));
        NI_REQUIRE(ptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled is not available from this driver");
        fn = reinterpret_cast<EncodeFn>(ptr);
    }
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, const_cast<void*>(base), dims, strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    swizzle ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        ni_set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
        return NI_ERR_CUDA;
    }
    return NI_OK;
}
int ni_encode_tiled(CUtensorMap* tm, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes, const cuuint32_t* box) {
    return ni_encode_tiled_sw(tm, base, rank, dims, strides_bytes, box, 1);
}

namespace {

int encode_map(CUtensorMap* tm, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes, const cuuint32_t* box) {
    return ni_encode_tiled(tm, base, rank, dims, strides_bytes, box);
}

// NHWC activation view (n, h, w, c) with channel pitch: 4-D map, box (32, bw, bh, bn)
int encode_act_map(CUtensorMap* tm, const float* base, int n, int h, int w, int c, int pitch, int bw, int bh, int bn) {
    cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
    cuuint64_t str[3] = {(cuuint64_t)pitch * 4, (cuuint64_t)w * pitch * 4, (cuuint64_t)h * w * pitch * 4};
    cuuint32_t box[4] = {32, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn};
    return encode_map(tm, base, 4, dims, str, box);
}

// Logical (n, h, w, 4F) view of a physical (n, 2h, 2w, F [pitch]) buffer through TensorFlow's block-major depth_to_space(2):
// logical channel blk * F + f of pixel (y, x) = physical channel f of pixel (2y + blk / 2, 2x + blk % 2). 5-D map
// (F, dx:2, w, dy:2, h*n) with box (32, 1, bw, 1, bh*bn); rows of consecutive images are contiguous in the merged last
// dimension, which is all a 1x1 filter needs (no halo, no zero fill).
int encode_block2_map(CUtensorMap* tm, const float* base, int n, int h, int w, int f, int pitch, int bw, int bh, int bn) {
    cuuint64_t dims[5] = {(cuuint64_t)f, 2, (cuuint64_t)w, 2, (cuuint64_t)h * n};
    cuuint64_t str[4] = {(cuuint64_t)pitch * 4, (cuuint64_t)2 * pitch * 4, (cuuint64_t)2 * w * pitch * 4, (cuuint64_t)4 * w * pitch * 4};
    cuuint32_t box[5] = {32, 1, (cuuint32_t)bw, 1, (cuuint32_t)(bh * bn)};
    return encode_map(tm, base, 5, dims, str, box);
}

// gemm tiles: 16 x 8 pixels where the image allows it (smallest halo), else the row-major rule of pick_tile
bool pick_tile(int h, int w, int pixels, int& bw, int& bh, int& bn);
bool pick_tile_gemm(int h, int w, int& bw, int& bh, int& bn) {
    if (w % 16 == 0 && h % 8 == 0) { bw = 16; bh = 8; bn = 1; return true; }
    return pick_tile(h, w, 128, bw, bh, bn);
}

// tile + halo geometry of the gemm kernel for a (th x tw) target and a kh x kw filter; false if it does not fit
bool gemm_geometry(int th, int tw, int kh, int kw, int& bw, int& bh, int& bn, int& hw, int& hh, int& a_stage) {
    if (!pick_tile_gemm(th, tw, bw, bh, bn)) return false;
    hw = bw + kw - 1; hh = bh + kh - 1;
    a_stage = (hw * hh * bn * 128 + 1023) / 1024 * 1024;
    return hw <= 256 && hh <= 256 && bn <= 256 && a_stage <= 48 * 1024;
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

int bnt_cap() {
    static int cap = 0;
    if (!cap) { const char* e = dev_env("NI_TC_BNT_MAX"); cap = e ? atoi(e) : 128; if (cap != 32 && cap != 64) cap = 128; }
    return cap;
}
int pick_bnt(int n) {
    const int cap = bnt_cap();
    if (n % 128 == 0 && cap >= 128) return 128;
    if (n % 64 == 0 && cap >= 64) return 64;
    return n % 32 == 0 ? 32 : 0;
}

// N-tile width of the persistent gemm chosen against WAVE QUANTISATION: one CTA per SM walks mtiles * N / bnt tiles; with few pixel tiles
// (deep U-Net layers, small per-GPU batches: 16 - 64 tiles of 128 pixels) a 128-wide tile leaves most SMs idle. Cost of a choice =
// waves * cycles per 32-deep 3xTF32 k-iteration of that width (12 MMAs of 20.5 + 0.42 N cycles, measured: profiles/r1_hw_probes.txt).
int pick_bnt_gemm(int n, int mtiles) {
    const int widest = pick_bnt(n);
    if (widest <= 32) return widest;
    const int sms = ni_num_sms();
    int best = widest;
    double best_cost = 1e30;
    for (int b = widest; b >= 32; b >>= 1) {
        if (n % b) continue;
        const long long tiles = (long long)mtiles * (n / b);
        const long long waves = (tiles + sms - 1) / sms;
        const double cost = (double)waves * (20.5 + 0.42 * b);
        if (cost < best_cost * 0.97) { best_cost = cost; best = b; }      // prefer the wider tile on near ties (less weight traffic per MMA)
    }
    return best;
}

template <typename K>
int set_dyn_smem(K kern, size_t bytes) {
    NI_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return NI_OK;
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// Activation backward of the layer BELOW fused into a dgrad (see GemmParams::dact_y)
struct FusedAct {
    const float* y = nullptr; int pitch = 0, coff = 0, act = NI_ACT_NONE; float alpha = 0.f; float* dbias = nullptr; int bias_mod = 0;
};

// Shared launcher for fprop (dgrad = false) and dgrad (dgrad = true).
int launch_gemm(const ni_conv_desc* d, bool dgrad, const float* src, const float* w, const float* bias, float* dst, cudaStream_t st,
                const FusedAct& fa = FusedAct()) {
    const int Kr = dgrad ? d->cout : d->cin, Nr = dgrad ? d->cin : d->cout;            // real channel counts
    const int K = (Kr + 31) & ~31, N = (Nr + 31) & ~31;                                // what the tiles are padded to (see tc_prep_weights_kernel)
    const int sh = dgrad ? d->oh : d->h, sw = dgrad ? d->ow : d->w;           // source dims
    const int th = dgrad ? d->h : d->oh, tw = dgrad ? d->w : d->ow;           // target dims
    const int spitch = dgrad ? d->out_pitch : d->in_pitch, scoff = dgrad ? d->out_coff : d->in_coff;
    const int dpitch = dgrad ? d->in_pitch : d->out_pitch, dcoff = dgrad ? d->in_coff : d->out_coff;
    const int dmode = dgrad ? d->in_mode : d->out_mode;
    const int smode = dgrad ? d->out_mode : d->in_mode;
    const int taps = d->kh * d->kw;
    tcv2::GemmParams p;
    if (!gemm_geometry(th, tw, d->kh, d->kw, p.bw, p.bh, p.bn, p.hw, p.hh, p.a_stage)) {
        ni_set_error("conv_tc: unsupported spatial tile");
        return NI_ERR_UNSUPPORTED;
    }
    p.sa = tcv2::kMaxSA;           // two halo stages: the next chunk / the next tile's halo is in flight while this one is converted
    const int bnt = pick_bnt_gemm(N, (tw / p.bw) * (th / p.bh) * ((d->n + p.bn - 1) / p.bn));
    float* scratch = nullptr;
    const size_t wbytes = (size_t)2 * taps * K * N * sizeof(float);
    int rc = get_scratch(wbytes, &scratch);
    if (rc) return rc;
    if (K != Kr || N != Nr) NI_CUDA(cudaMemsetAsync(scratch, 0, wbytes, st));          // padded rows / columns of the [hi | lo] tiles
    const long long total = (long long)taps * Kr * Nr;
    tc_prep_weights_kernel<<<ni_cdiv(total, 256), 256, 0, st>>>(w, scratch, taps, d->cin, d->cout, dgrad ? 0 : 1, bnt, K);
    NI_LAUNCH_CHECK();
    CUtensorMap tmA;
    p.src_block2_f = smode == NI_MODE_BLOCK2 ? K / 4 : 0;
    if (p.src_block2_f > 0) rc = encode_block2_map(&tmA, src + scoff, d->n, sh, sw, K / 4, spitch, p.hw, p.hh, p.bn);
    else rc = encode_act_map(&tmA, src + scoff, d->n, sh, sw, Kr, spitch, p.hw, p.hh, p.bn);     // channels >= Kr: out-of-bounds zero fill
    if (rc) return rc;
    p.n = d->n; p.oh = th; p.ow = tw;
    p.tiles_w = tw / p.bw; p.tiles_h = th / p.bh;
    const int tiles_n = (d->n + p.bn - 1) / p.bn;
    p.kh = d->kh; p.kw = d->kw;
    p.off_y0 = dgrad ? d->pad_t : -d->pad_t; p.off_x0 = dgrad ? d->pad_l : -d->pad_l; p.off_sign = dgrad ? -1 : 1;
    p.kchunks = K / 32; p.ntot = N; p.n_valid = Nr;
    p.out_pitch = dpitch; p.out_coff = dcoff; p.out_mode = dmode;
    p.bias_mod = dgrad ? 0 : d->bias_mod; p.act = dgrad ? NI_ACT_NONE : d->act; p.accumulate = d->accumulate; p.alpha = d->act_alpha;
    p.bias = dgrad ? nullptr : bias; p.out = dst;
    p.dact_y = fa.y; p.dact_pitch = fa.pitch; p.dact_coff = fa.coff; p.dact = fa.act; p.dact_alpha = fa.alpha;
    p.dbias = fa.dbias; p.dbias_mod = fa.bias_mod;
    const int ktot = taps * K;
    const int want_nacc = ktot > 2304 ? 3 : (ktot > 1024 ? 2 : 1);
    static const bool use_v2 = dev_env("NI_TC_GEMM_V2") != nullptr;
    const int mtiles = p.tiles_w * p.tiles_h * tiles_n;
    if (!use_v2) {
        // generation 3: persistent CTAs (one per SM), see conv_tc_v3.cuh
        tcv3::PersistParams q;
        q.mtiles = mtiles; q.total_tiles = mtiles * (N / bnt);
        // Weight stream: groups of G k-iterations = one 32 KB bulk copy (the copy engine completes ~1 operation per 735 cycles whatever its
        // size up to 32 KB). Resident: when the whole [hi | lo] weight slice of the (single) n-tile fits beside the A stages it is loaded once
        // per CTA instead of once per pixel tile.
        const int iters = taps * (K / 32);
        const int G = tcv3::kGroupBytes / (2 * bnt * 128);
        const int ngroups = (iters + G - 1) / G;
        const int grid_ctas = q.total_tiles < ni_num_sms() ? q.total_tiles : ni_num_sms();
        const int budget = 226 * 1024 - 1024;
        if (p.sa * p.a_stage + tcv3::kGroupBytes > budget) p.sa = 1;
        const int room = (budget - p.sa * p.a_stage) / tcv3::kGroupBytes;       // 32 KB groups that fit beside the A stages
        static const int resident_on = dev_env("NI_TC_B_RESIDENT") ? atoi(dev_env("NI_TC_B_RESIDENT")) : 1;
        q.b_resident = (resident_on && N == bnt && ngroups <= room && ngroups <= tcv3::kMaxGroups && q.total_tiles >= 2 * grid_ctas) ? 1 : 0;
        if (q.b_resident) { q.sb = ngroups; q.log_sb = 0; }
        else if (room >= 4) { q.sb = 4; q.log_sb = 2; }
        else if (room >= 2) { q.sb = 2; q.log_sb = 1; }
        else { q.sb = room >= 1 ? 1 : 0; q.log_sb = 0; }
        static const int defer_env = dev_env("NI_TC_DEFER") ? atoi(dev_env("NI_TC_DEFER")) : -1;
        q.defer_st = defer_env >= 0 ? (defer_env >> (bnt == 128 ? 2 : (bnt == 64 ? 1 : 0))) & 1 : 1;      // bit 0: N = 32, bit 1: N = 64, bit 2: N = 128
        q.experiment = dev_env("NI_TC_EXP") ? atoi(dev_env("NI_TC_EXP")) : 0;
        const int bstage = tcv3::kGroupBytes;
        // TMEM budget (tcv3::Cfg): N = 128 -> one set of (nacc + 1) accumulators + 2 A slots; N <= 64 -> fixed 4 accumulators per set
        int nacc = want_nacc;
        while (bnt == 128 && (nacc + 1) * bnt + 2 * 64 > 512) --nacc;
        if (nacc > iters) nacc = iters;
        p.nacc = nacc < 1 ? 1 : nacc;
        // (measured: four A slots instead of two for N = 128, possible with a single hi*hi accumulator, change nothing: 0.413 vs 0.417 ms)
        if (q.sb < 1) { ni_set_error("conv_tc: not enough shared memory for the weight ring"); return NI_ERR_UNSUPPORTED; }
        const size_t smem = (size_t)p.sa * p.a_stage + (size_t)q.sb * bstage + 1024;
        const int grid = grid_ctas;
#define NI_TC_GEMM3(B, SL)                                                                                     \
    {                                                                                                          \
        rc = set_dyn_smem(tcv3::conv_tc3_gemm_kernel<B, SL>, smem);                                            \
        if (rc) return rc;                                                                                     \
        if (dev_env("NI_TC_DEBUG")) {                                                                           \
            cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, tcv3::conv_tc3_gemm_kernel<B, SL>);              \
            fprintf(stderr, "tc gemm3<%d,%d>: grid %d, tiles %d, iters %d, smem %zu, regs %d, local %zu, sb %d%s, sa %d, nacc %d\n", B, SL, grid,  \
                    q.total_tiles, taps * (K / 32), smem, fa.numRegs, fa.localSizeBytes, q.sb, q.b_resident ? " (resident)" : "", p.sa, p.nacc);  \
        }                                                                                                      \
        tcv3::conv_tc3_gemm_kernel<B, SL><<<grid, tcv3::kThreads, smem, st>>>(tmA, scratch, p, q);             \
    }
        if (bnt == 128) NI_TC_GEMM3(128, 2) else if (bnt == 64) NI_TC_GEMM3(64, 4) else NI_TC_GEMM3(32, 4)
#undef NI_TC_GEMM3
        NI_LAUNCH_CHECK();
        NI_COUNT_LAUNCH(2);
        return NI_OK;
    }
#ifdef NI_DEV
    {
        // generation 2 (one CTA per tile): accumulator rotation bounded by the 512 TMEM columns
        int nacc = want_nacc;
        while ((nacc + 1) * bnt + tcv2::kTmemSlots * 64 > 512) --nacc;
        if (nacc > taps * (K / 32)) nacc = taps * (K / 32);
        p.nacc = nacc < 1 ? 1 : nacc;
    }
    dim3 grid((unsigned)mtiles, (unsigned)(N / bnt));
#define NI_TC_GEMM(B)                                                                                          \
    {                                                                                                          \
        const size_t smem = (size_t)p.sa * p.a_stage + tcv2::Rings<B>::SMEM_B + 1024;                          \
        rc = set_dyn_smem(tcv2::conv_tc2_gemm_kernel<B>, smem);                                                \
        if (rc) return rc;                                                                                     \
        tcv2::conv_tc2_gemm_kernel<B><<<grid, tcv2::Roles<B>::THREADS, smem, st>>>(tmA, scratch, p);           \
    }
    if (bnt == 128) NI_TC_GEMM(128) else if (bnt == 64) NI_TC_GEMM(64) else NI_TC_GEMM(32)
#undef NI_TC_GEMM
    NI_LAUNCH_CHECK();
    NI_COUNT_LAUNCH(2);
    return NI_OK;
#else
    ni_set_error("conv_tc: the generation-2 gemm exists in development builds only");
    return NI_ERR_UNSUPPORTED;
#endif
}

}  // namespace

#ifdef NI_DEV
// In-kernel timing counters of the persistent gemm [0,32) and of wgrad [32,64) (all zero unless the library was built with -DNI_TC_PROFILE)
extern "C" int ni_tc_prof_read(long long* out64, int reset) {
    NI_REQUIRE(out64, "ni_tc_prof_read: null pointer");
#ifdef NI_TC_PROFILE
    NI_CUDA(cudaDeviceSynchronize());
    NI_CUDA(cudaMemcpyFromSymbol(out64, tc::g_tc_prof, sizeof(long long) * 64));
    if (reset) { long long z[64] = {0}; NI_CUDA(cudaMemcpyToSymbol(tc::g_tc_prof, z, sizeof(z))); }
#else
    for (int i = 0; i < 64; ++i) out64[i] = 0;
#endif
    return NI_OK;
}

#endif

// Second internal scratch (transposed / flipped weights of the SIMT and direct dgrad paths), same grow-on-demand policy.
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
    // fewer than 32 channels on ONE side of a 32-multiple layer (fprop: outputs, dgrad: contraction) run on zero-padded tiles
    const bool narrow_out = op != 2 && d->cin % 32 == 0 && d->cout < 32 && d->cout % 4 == 0 && d->cout >= 8;
    const bool narrow_in = op != 2 && d->cout % 32 == 0 && d->cin < 32 && d->cin % 4 == 0 && d->cin >= 8;      // e.g. 12 -> 64: the image-end DCN layer in space_to_depth form
    if (!narrow_out && !narrow_in && (d->cin % 32 || d->cout % 32)) return 0;
    if (narrow_out && (d->in_mode != NI_MODE_PLAIN || (op == 1 && d->out_mode != NI_MODE_PLAIN) || d->bias_mod != 0)) return 0;
    if (narrow_in && (d->in_mode != NI_MODE_PLAIN || d->out_mode != NI_MODE_PLAIN)) return 0;
    if (d->kh * d->kw > 64) return 0;
    if ((d->in_pitch % 4) || (d->in_coff % 4)) return 0;
    if (((d->out_pitch % 4) || (d->out_coff % 4)) && !(narrow_out && op == 0)) return 0;      // the narrow epilogue stores scalars / float2
    int bw, bh, bn, hw, hh, ast;
    if (op == 0) {
        if (d->in_mode != NI_MODE_PLAIN) return 0;
        if (d->out_mode == NI_MODE_BLOCK2 && ((d->cout / 4) % 32) && !narrow_out) return 0;
        return gemm_geometry(d->oh, d->ow, d->kh, d->kw, bw, bh, bn, hw, hh, ast) ? 1 : 0;
    }
    if (op == 1) {
        // dy is the TMA source: depth_to_space addressing only for 1x1 filters (the transposed convolutions) on whole 32-channel chunks
        if (d->out_mode != NI_MODE_PLAIN && !(d->kh == 1 && d->kw == 1 && ((d->cout / 4) % 32) == 0 && dev_env("NI_TC_NO_BLOCK2") == nullptr)) return 0;
        if (d->in_mode == NI_MODE_BLOCK2 && ((d->cin / 4) % 32)) return 0;
        return gemm_geometry(d->h, d->w, d->kh, d->kw, bw, bh, bn, hw, hh, ast) ? 1 : 0;
    }
    if (d->in_mode != NI_MODE_PLAIN) return 0;
    if (d->out_mode != NI_MODE_PLAIN && !(d->kh == 1 && d->kw == 1 && ((d->cout / 4) % 32) == 0 && dev_env("NI_TC_NO_BLOCK2") == nullptr)) return 0;
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

// dgrad whose epilogue also applies the activation derivative of the layer that produced this layer's input and accumulates that layer's
// bias gradient: dx <- (dgrad result) * act'(y_prev), dbias_prev += column sums. y_prev is addressed like dx (plain NHWC, own pitch / offset).
extern "C" int ni_conv2d_dgrad_act_supported(const ni_conv_desc* d, int y_pitch, int y_coff) {
    if (!ni_conv2d_tc_supported(d, 1)) return 0;
    if (d->in_mode != NI_MODE_PLAIN) return 0;                          // dx (and y_prev) plain NHWC
    if ((y_pitch & 3) || (y_coff & 3)) return 0;
    static const bool use_v2 = dev_env("NI_TC_GEMM_V2") != nullptr;      // the fused epilogue exists in the generation-3 kernel only
    return use_v2 ? 0 : 1;
}
extern "C" int ni_conv2d_dgrad_act_tc(const ni_conv_desc* d, const float* dy, const float* w, float* dx, const float* y_prev, int y_pitch,
                                      int y_coff, int act_prev, float alpha_prev, float* dbias_prev, int bias_mod_prev, cudaStream_t st) {
    NI_REQUIRE(ni_conv2d_dgrad_act_supported(d, y_pitch, y_coff), "ni_conv2d_dgrad_act_tc: problem not supported by the fused tcgen05 path");
    NI_REQUIRE(dy && w && dx && aligned16(dy) && aligned16(dx) && (y_prev == nullptr || aligned16(y_prev)), "ni_conv2d_dgrad_act_tc: null or unaligned pointer");
    NI_REQUIRE(act_prev == NI_ACT_NONE || act_prev == NI_ACT_LEAKY_RELU || act_prev == NI_ACT_RELU || act_prev == NI_ACT_TANH ||
               act_prev == NI_ACT_SIGMOID, "ni_conv2d_dgrad_act_tc: activation %d has no derivative in terms of its output", act_prev);
    NI_REQUIRE(y_prev != nullptr || act_prev == NI_ACT_NONE, "ni_conv2d_dgrad_act_tc: activation backward needs the forward output");
    FusedAct fa;
    fa.y = y_prev; fa.pitch = y_pitch; fa.coff = y_coff; fa.act = act_prev; fa.alpha = alpha_prev; fa.dbias = dbias_prev; fa.bias_mod = bias_mod_prev;
    if (fa.y == nullptr && fa.dbias == nullptr) return launch_gemm(d, true, dy, w, nullptr, dx, st);
    if (fa.dbias) NI_CUDA(cudaMemsetAsync(fa.dbias, 0, sizeof(float) * (size_t)(bias_mod_prev > 0 ? bias_mod_prev : d->cin), st));   // the epilogue accumulates
    return launch_gemm(d, true, dy, w, nullptr, dx, st, fa);
}

extern "C" int ni_conv2d_wgrad_tc(const ni_conv_desc* d, const float* x, const float* dy, float* dw, cudaStream_t st) {
    NI_REQUIRE(ni_conv2d_tc_supported(d, 2), "ni_conv2d_wgrad_tc: problem not supported by the tcgen05 path");
    NI_REQUIRE(x && dy && dw && aligned16(x) && aligned16(dy), "ni_conv2d_wgrad_tc: null or unaligned pointer");
    const int taps = d->kh * d->kw;
    if (!d->accumulate) NI_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)taps * d->cin * d->cout, st));
    tcv2::WgradParams p;
    pick_tile(d->oh, d->ow, 32, p.bw, p.bh, p.bn);
    CUtensorMap tmX, tmDY;
    int rc = encode_act_map(&tmX, x + d->in_coff, d->n, d->h, d->w, d->cin, d->in_pitch, p.bw, p.bh, p.bn);
    if (rc) return rc;
    p.dy_block2_f = d->out_mode == NI_MODE_BLOCK2 ? d->cout / 4 : 0;
    if (p.dy_block2_f > 0) rc = encode_block2_map(&tmDY, dy + d->out_coff, d->n, d->oh, d->ow, d->cout / 4, d->out_pitch, p.bw, p.bh, p.bn);
    else rc = encode_act_map(&tmDY, dy + d->out_coff, d->n, d->oh, d->ow, d->cout, d->out_pitch, p.bw, p.bh, p.bn);
    if (rc) return rc;
    p.n = d->n; p.oh = d->oh; p.ow = d->ow;
    p.tiles_w = d->ow / p.bw; p.tiles_h = d->oh / p.bh;
    p.kw = d->kw; p.pad_t = d->pad_t; p.pad_l = d->pad_l;
    p.cin_chunks = d->cin / 32; p.atoms = taps * p.cin_chunks; p.mtot = taps * d->cin; p.cout = d->cout;
    p.steps_total = p.tiles_w * p.tiles_h * (d->n / p.bn);
    p.dw = dw;
    p.defer_st = 0;
    const int bnt = pick_bnt(d->cout);
    const int mtiles = (p.atoms + 3) / 4, ntiles = d->cout / bnt;
    // split-K over pixel ranges. Cost model (us): waves x steps per CTA x t_iter + per-wave fixed cost + the cross-CTA reduction
    // (every split adds mtot x cout floats through L2 reductions). t_iter = 1.5 us: the loop is bound by L2 -> SM operand traffic
    // (16 KB of A + BNT x 128 B of B per 32 pixels ~ 3.5 TB/s over 148 SMs), measured 1.4 - 1.8 us for every tile width. The old rule (always
    // ~4 waves) made the 1x1 transposed-conv layers reduction-bound: 592 CTAs x 16 K atomics for 28 iterations of work each.
    // Chains are capped at 2048 steps (65 K pixels) to bound the truncation bias of the in-TMEM accumulation.
    static const bool wgrad_v2 = dev_env("NI_TC_WGRAD_V2") != nullptr;      // generation 2 kernel (one producer thread, two CTAs per SM for N <= 64)
    const int tiles = mtiles * ntiles, sms = ni_num_sms() * ((bnt == 128 || !wgrad_v2) ? 1 : 2);   // resident CTAs
    const double t_iter = wgrad_v2 ? 1.5 : 0.7;                             // us per 32-pixel step of one CTA
    const int min_splits = (p.steps_total + 2047) / 2048;
    int max_splits = (p.steps_total + 15) / 16;
    if (max_splits > 1024) max_splits = 1024;
    if (max_splits < min_splits) max_splits = min_splits;
    int splits = min_splits;
    double best = 1e30;
    for (int sp = min_splits; sp <= max_splits; ++sp) {
        const int per = (p.steps_total + sp - 1) / sp;
        const int eff = (p.steps_total + per - 1) / per;
        const int waves = (tiles * eff + sms - 1) / sms;
        const double cost = waves * (per * t_iter + 6.0) + (double)eff * p.mtot * d->cout / 50e3;
        if (cost < best) { best = cost; splits = eff; }
    }
    p.steps_per_split = (p.steps_total + splits - 1) / splits;
    splits = (p.steps_total + p.steps_per_split - 1) / p.steps_per_split;
    dim3 grid((unsigned)mtiles, (unsigned)ntiles, (unsigned)splits);
    if (dev_env("NI_TC_DEBUG"))
        fprintf(stderr, "tc wgrad<%d>%s: grid (%d, %d, %d), steps %d total, %d per split, atoms %d\n", bnt, wgrad_v2 ? " v2" : "", mtiles, ntiles, splits,
                p.steps_total, p.steps_per_split, p.atoms);
    if (!wgrad_v2) {
#define NI_TC_WGRAD3(B, NPW)                                                                                   \
    {                                                                                                          \
        using C = tcw3::Cfg<B, NPW>;                                                                           \
        const size_t smem = (size_t)C::STAGES * C::STAGE_BYTES + 1024;                                         \
        rc = set_dyn_smem(tcw3::conv_tc3_wgrad_kernel<B, NPW>, smem);                                          \
        if (rc) return rc;                                                                                     \
        tcw3::conv_tc3_wgrad_kernel<B, NPW><<<grid, C::THREADS, smem, st>>>(tmX, tmDY, p);                     \
    }
        // producer warps per tile width (measured, tools/profile_conv.py): the copy engine's row rate (one 128-byte box row per ~4.5
        // cycles) bounds a step, more issuing warps than needed to reach it only add polling
        static const int np_env = dev_env("NI_TC_WG_NP") ? atoi(dev_env("NI_TC_WG_NP")) : 0;
        const bool many = np_env > 4;        // measured: 4 producer warps are enough for every tile width (more only add polling)
        static const int defer_env = dev_env("NI_TC_WG_DEFER") ? atoi(dev_env("NI_TC_WG_DEFER")) : -1;
        p.defer_st = defer_env > 0 ? 1 : 0;   // measured: deferring the converters' tcgen05.st completion delays the MMA warp more than it hides (-10 .. -25 %)
        if (bnt == 128) { if (many) NI_TC_WGRAD3(128, 8) else NI_TC_WGRAD3(128, 4) }
        else if (bnt == 64) { if (many) NI_TC_WGRAD3(64, 6) else NI_TC_WGRAD3(64, 4) }
        else { if (many) NI_TC_WGRAD3(32, 5) else NI_TC_WGRAD3(32, 4) }
#undef NI_TC_WGRAD3
        NI_LAUNCH_CHECK();
        NI_COUNT_LAUNCH(1);
        return NI_OK;
    }
#ifdef NI_DEV
#define NI_TC_WGRAD(B)                                                                                         \
    {                                                                                                          \
        const size_t smem = (size_t)tcv2::WgCfg<B>::STAGES * (tcv2::kAraw + 3 * B * 128) + 1024;                      \
        rc = set_dyn_smem(tcv2::conv_tc2_wgrad_kernel<B>, smem);                                               \
        if (rc) return rc;                                                                                     \
        tcv2::conv_tc2_wgrad_kernel<B><<<grid, tcv2::kThreadsWg, smem, st>>>(tmX, tmDY, p);                    \
    }
    if (bnt == 128) NI_TC_WGRAD(128) else if (bnt == 64) NI_TC_WGRAD(64) else NI_TC_WGRAD(32)
#undef NI_TC_WGRAD
    NI_LAUNCH_CHECK();
    NI_COUNT_LAUNCH(1);
    return NI_OK;
#else
    ni_set_error("conv_tc: the generation-2 wgrad exists in development builds only");
    return NI_ERR_UNSUPPORTED;
#endif
}
