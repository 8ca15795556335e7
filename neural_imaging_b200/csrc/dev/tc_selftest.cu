// Self-test of the UMMA shared-memory layouts used by conv_tc.cu: one CTA fills A (128 x 32) and B (N x 32) in shared
// memory with plain stores in the canonical SWIZZLE_128B layout (K-major or MN-major), issues 4 x tcgen05.mma
// (kind::tf32, K = 8), and writes the 128 x N accumulator to global memory. The host compares with a CPU matmul.
// It pins the descriptor semantics (start address advance, LBO / SBO roles, major bits) independently of TMA.
#include "ni_common.cuh"
#include "tc_common.cuh"

namespace {
using namespace tc;

constexpr int kN = 64;

// byte offset of element (row, col) in a K-major SW128 tile (row = M/N index, col = k in [0, 32))
__device__ __forceinline__ int off_kmajor(int row, int k) {
    const int atom = row >> 3, r = row & 7;
    return atom * 1024 + r * 128 + (((k >> 2) ^ r) << 4) + (k & 3) * 4;
}
// byte offset of element (mn, k) in an MN-major SW128 tile: 32-wide MN atoms `lbo` apart, 8-deep k groups `sbo` apart
__device__ __forceinline__ int off_mnmajor(int mn, int k, int lbo, int sbo) {
    const int ma = mn >> 5, e = mn & 31, kg = k >> 3, r = k & 7;
    return ma * lbo + kg * sbo + r * 128 + (((e >> 2) ^ r) << 4) + (e & 3) * 4;
}

__global__ void __launch_bounds__(128, 1) tc_selftest_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                             float* __restrict__ D, int mn_major) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sa = smem;                 // 128 x 32 floats = 16 KB
    uint8_t* sb = smem + 16384;         // kN x 32 floats
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // MN-major: 4 KB per 32-wide MN atom (32 k rows x 128 B), k groups of 8 rows = 1 KB
    for (int i = threadIdx.x; i < 128 * 32; i += 128) {
        const int m = i / 32, k = i % 32;
        const int o = mn_major == 1 ? off_mnmajor(m, k, 4096, 1024) : off_kmajor(m, k);
        *reinterpret_cast<float*>(sa + o) = A[m * 32 + k];
    }
    for (int i = threadIdx.x; i < kN * 32; i += 128) {
        const int n = i / 32, k = i % 32;
        const int o = mn_major == 1 ? off_mnmajor(n, k, 4096, 1024) : off_kmajor(n, k);
        *reinterpret_cast<float*>(sb + o) = B[n * 32 + k];
    }
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(&tmem_slot, 128);
    fence_proxy_async_smem();
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = tmem_slot;
    if (mn_major == 2) {
        // A operand through tensor memory: thread = row, 32 K-columns at TMEM columns [64, 96)
        float v[32];
        for (int k = 0; k < 32; ++k) v[k] = A[(warp * 32 + lane) * 32 + k];
        tmem_st_32x32(tmem + ((uint32_t)(warp * 32) << 16) + 64u, v);
        tmem_st_wait();
        tcgen05_fence_before();
        __syncthreads();
        tcgen05_fence_after();
    }
    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc_tf32(128, kN, mn_major == 1, mn_major == 1);
        for (int ks = 0; ks < 4; ++ks) {
            uint64_t da, db;
            if (mn_major == 2) {
                db = make_smem_desc_sw128(smem_u32(sb) + ks * 32, 16, 1024);
                umma_tf32_ts(tmem, tmem + 64u + ks * 8, db, idesc, ks > 0 ? 1u : 0u);
                continue;
            }
            if (mn_major == 1) {
                da = make_smem_desc_sw128(smem_u32(sa) + ks * 1024, 4096, 1024);
                db = make_smem_desc_sw128(smem_u32(sb) + ks * 1024, 4096, 1024);
            } else {
                da = make_smem_desc_sw128(smem_u32(sa) + ks * 32, 16, 1024);
                db = make_smem_desc_sw128(smem_u32(sb) + ks * 32, 16, 1024);
            }
            umma_tf32(tmem, da, db, idesc, ks > 0 ? 1u : 0u);
        }
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0, 9);
    tcgen05_fence_after();
    const int row = warp * 32 + lane;
    for (int c = 0; c < kN / 32; ++c) {
        float v[32];
        tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(c * 32), v);
        for (int j = 0; j < 32; ++j) D[row * kN + c * 32 + j] = v[j];
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) { tcgen05_fence_after(); tmem_dealloc(tmem, 128); }
}
}  // namespace

// A: (128, 32), B: (64, 32), D: (128, 64) device pointers. D = A * B^T computed by tcgen05 with the given operand majorness.
extern "C" int ni_tc_selftest(const float* a, const float* b, float* d, int mn_major, cudaStream_t st) {
    NI_REQUIRE(a && b && d, "ni_tc_selftest: null pointer");
    const size_t smem = 16384 + kN * 128 + 1024;
    NI_CUDA(cudaFuncSetAttribute(tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tc_selftest_kernel<<<1, 128, smem, st>>>(a, b, d, mn_major);
    NI_LAUNCH_CHECK();
    return NI_OK;
}
