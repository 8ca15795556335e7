// Measurement probe (not on the product path): how fast can ONE thread per CTA stream 16 KB activation boxes
// (32 channels x 128 pixels of an NHWC tensor) through TMA into shared memory, as a function of the channel pitch and of
// the number of boxes kept in flight? Used to size the tcgen05 convolution's TMA rings (profiles/README.md).
#include <cuda.h>

#include "ni_common.cuh"
#include "tc_common.cuh"

int ni_encode_tiled(CUtensorMap* tm, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes, const cuuint32_t* box);

namespace {
using namespace tc;

template <int STAGES>
__global__ void __launch_bounds__(32, 1) tma_probe_kernel(const __grid_constant__ CUtensorMap tm, int tiles_w, int tiles_h, int bw, int bh,
                                                          int kchunks, int boxes_per_cta, long long* cycles_out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar[STAGES];
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(&bar[s], 1);
        fence_barrier_init();
    }
    __syncwarp();
    if (threadIdx.x != 0) return;
    const long long t0 = clock64();
    const int total_tiles = tiles_w * tiles_h;
    for (int i = 0; i < boxes_per_cta + STAGES; ++i) {
        if (i >= STAGES) mbar_wait(&bar[(i - STAGES) % STAGES], ((i - STAGES) / STAGES) & 1, 0);
        if (i < boxes_per_cta) {
            const int box = blockIdx.x * boxes_per_cta + i;
            const int kc = box % kchunks, tile = (box / kchunks) % total_tiles, n = box / (kchunks * total_tiles);
            const int s = i % STAGES;
            mbar_expect_tx(&bar[s], 16384);
            tma_load_4d(smem + s * 16384, &tm, &bar[s], kc * 32, (tile % tiles_w) * bw, (tile / tiles_w) * bh, n);
        }
    }
    cycles_out[blockIdx.x] = clock64() - t0;
}
}  // namespace

// x: (n, h, w, c) float32 device tensor (c % 32 == 0, w in {8..128}). Every CTA streams `boxes_per_cta` boxes with `stages`
// (2, 4 or 8) in flight; cycles_out[grid] receives the per-CTA cycle counts. Returns the grid size used, or a negative error.
extern "C" int ni_tma_probe(const float* x, int n, int h, int w, int c, int stages, int boxes_per_cta, long long* cycles_out, int max_grid,
                            cudaStream_t st) {
    NI_REQUIRE(x && cycles_out && c % 32 == 0 && w <= 128 && 128 % w == 0, "ni_tma_probe: invalid arguments");
    const int bw = w, bh = 128 / w;
    NI_REQUIRE(h % bh == 0, "ni_tma_probe: h must be a multiple of 128 / w");
    CUtensorMap tm;
    cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
    cuuint64_t str[3] = {(cuuint64_t)c * 4, (cuuint64_t)w * c * 4, (cuuint64_t)h * w * c * 4};
    cuuint32_t box[4] = {32, (cuuint32_t)bw, (cuuint32_t)bh, 1};
    int rc = ni_encode_tiled(&tm, x, 4, dims, str, box);
    if (rc) return rc;
    const long long total_boxes = (long long)n * (h / bh) * (w / bw) * (c / 32);
    int grid = (int)(total_boxes / boxes_per_cta);
    if (grid > max_grid) grid = max_grid;
    NI_REQUIRE(grid >= 1, "ni_tma_probe: tensor too small");
    const size_t smem = (size_t)stages * 16384 + 1024;
#define NI_PROBE(S)                                                                                                        \
    {                                                                                                                      \
        NI_CUDA(cudaFuncSetAttribute(tma_probe_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));        \
        tma_probe_kernel<S><<<grid, 32, smem, st>>>(tm, w / bw, h / bh, bw, bh, c / 32, boxes_per_cta, cycles_out);        \
    }
    if (stages == 2) NI_PROBE(2) else if (stages == 4) NI_PROBE(4) else if (stages == 8) NI_PROBE(8) else {
        ni_set_error("ni_tma_probe: stages must be 2, 4 or 8");
        return NI_ERR_ARG;
    }
#undef NI_PROBE
    NI_LAUNCH_CHECK();
    return grid;
}

// ------------------------------------------------------------------------------------------------------------------
// Probe 2: tcgen05.mma rate. One CTA per SM; one elected lane issues `rounds` x 12 back-to-back kind::tf32 MMAs (M = 128, N = n,
// K = 8) on garbage operands, A from tensor memory (ts = 1) or shared memory (ts = 0), then commits and waits. Reports the cycles
// per MMA of every CTA: the hardware floor the convolution main loops are measured against.
namespace {
using namespace tc;

__global__ void __launch_bounds__(128, 1) mma_probe_kernel(int n, int ts, int rounds, int nacc, long long* cycles_out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    for (int i = threadIdx.x; i < (16384 + 32768 + 8192) / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 1.0f;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    fence_proxy_async_smem();
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = __shfl_sync(0xffffffffu, tmem_slot, 0);
    if (warp == 0) {
        const uint32_t idesc = make_idesc_tf32(128, n, 0, 0);
        const uint64_t da0 = make_smem_desc_sw128(smem_u32(smem), 16, 1024), db0 = make_smem_desc_sw128(smem_u32(smem + 16384), 16, 1024);
        const uint32_t a_t = tmem + 448;       // 64 columns of A at the top of tensor memory
        const long long t0 = clock64();
        for (int r = 0; r < rounds; ++r) {
            const uint32_t d = tmem + (uint32_t)((r % nacc) * n);
            if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        if (ts) umma_tf32_ts(d, a_t + ks * 8 + (j == 0 ? 32 : 0), db0 + (uint64_t)(ks * 2 + (j == 1 ? 256 : 0)), idesc, 1u);
                        else umma_tf32(d, da0 + (uint64_t)(ks * 2), db0 + (uint64_t)(ks * 2 + (j == 1 ? 256 : 0)), idesc, 1u);
                    }
                }
            }
            __syncwarp();
        }
        const long long t1 = clock64();
        if (elect_one()) umma_commit(&bar);
        __syncwarp();
        mbar_wait(&bar, 0, 9);
        const long long t2 = clock64();
        if ((threadIdx.x & 31) == 0) { cycles_out[2 * blockIdx.x] = t1 - t0; cycles_out[2 * blockIdx.x + 1] = t2 - t0; }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) { tcgen05_fence_after(); tmem_dealloc(tmem, 512); }
}

// Probe 3: 1-D bulk copies (cp.async.bulk) global -> shared, `bytes` each, `depth` in flight per issuing warp, `copies` per warp.
// same = 1: every CTA reads the SAME source sequence (the weight stream of the convolution kernels), same = 0: its own region.
// `warps` issuing warps per CTA (lane 0 of each, own barriers and buffers): is the ~735-cycle cost per copy a property of the
// issuing thread or of the SM's copy engine?
__global__ void __launch_bounds__(256, 1) bulk_probe_kernel(const uint8_t* __restrict__ src, long long src_bytes, int bytes, int depth, int copies,
                                                            int same, int warps, long long* cycles_out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar[8][8];
    __shared__ long long t_end[8];
    if (threadIdx.x == 0) {
        for (int w = 0; w < 8; ++w) for (int s = 0; s < 8; ++s) mbar_init(&bar[w][s], 1);
        fence_barrier_init();
    }
    __syncthreads();
    const int w = threadIdx.x >> 5;
    const long long t0 = clock64();
    if ((threadIdx.x & 31) == 0 && w < warps) {
        const long long region = (same ? 0 : (long long)blockIdx.x * bytes * copies * warps) + (long long)w * bytes * copies;
        uint8_t* buf = smem + (size_t)w * depth * bytes;
        for (int i = 0; i < copies + depth; ++i) {
            if (i >= depth) mbar_wait(&bar[w][(i - depth) % depth], ((i - depth) / depth) & 1, 0);
            if (i < copies) {
                const int s = i % depth;
                const long long off = (region + (long long)i * bytes) % (src_bytes - bytes);
                mbar_expect_tx(&bar[w][s], bytes);
                bulk_load_1d(buf + (size_t)s * bytes, src + (off & ~15LL), bytes, &bar[w][s]);
            }
        }
        t_end[w] = clock64() - t0;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        long long m = 0;
        for (int i = 0; i < warps; ++i) m = t_end[i] > m ? t_end[i] : m;
        cycles_out[blockIdx.x] = m;
    }
}
}  // namespace

extern "C" int ni_mma_probe(int n, int ts, int rounds, int nacc, long long* cycles_out, int grid, cudaStream_t st) {
    NI_REQUIRE(cycles_out && n >= 16 && n <= 256 && n % 16 == 0 && rounds > 0 && nacc >= 1 && nacc * n <= 448 && grid > 0, "ni_mma_probe: invalid arguments");
    const size_t smem = 16384 + 32768 + 8192 + 1024;
    NI_CUDA(cudaFuncSetAttribute(mma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mma_probe_kernel<<<grid, 128, smem, st>>>(n, ts, rounds, nacc, cycles_out);
    NI_LAUNCH_CHECK();
    return NI_OK;
}

extern "C" int ni_bulk_probe(const void* src, long long src_bytes, int bytes, int depth, int copies, int same, int warps, long long* cycles_out,
                             int grid, cudaStream_t st) {
    NI_REQUIRE(src && cycles_out && bytes >= 1024 && bytes % 16 == 0 && depth >= 1 && depth <= 8 && warps >= 1 && warps <= 8 &&
                   (size_t)warps * depth * bytes <= 200 * 1024 && copies > 0 && src_bytes > 2LL * bytes && grid > 0, "ni_bulk_probe: invalid arguments");
    const size_t smem = (size_t)warps * depth * bytes + 1024;
    NI_CUDA(cudaFuncSetAttribute(bulk_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    bulk_probe_kernel<<<grid, 256, smem, st>>>(static_cast<const uint8_t*>(src), src_bytes, bytes, depth, copies, same, warps, cycles_out);
    NI_LAUNCH_CHECK();
    return NI_OK;
}
