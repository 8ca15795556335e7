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
