// Common helpers for the ni_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define NI_OK 0
#define NI_ERR_ARG (-1)
#define NI_ERR_CUDA (-2)
#define NI_ERR_UNSUPPORTED (-3)

// Thread-local last-error message, returned by ni_last_error().
void ni_set_error(const char* fmt, ...);

#define NI_REQUIRE(cond, ...)                   \
    do {                                        \
        if (!(cond)) {                          \
            ni_set_error(__VA_ARGS__);          \
            return NI_ERR_ARG;                  \
        }                                       \
    } while (0)

#define NI_CUDA(call)                                                           \
    do {                                                                        \
        cudaError_t e_ = (call);                                                \
        if (e_ != cudaSuccess) {                                                \
            ni_set_error("%s:%d %s: %s", __FILE__, __LINE__, #call,             \
                         cudaGetErrorString(e_));                               \
            return NI_ERR_CUDA;                                                 \
        }                                                                       \
    } while (0)

#define NI_LAUNCH_CHECK()                                                       \
    do {                                                                        \
        cudaError_t e_ = cudaGetLastError();                                    \
        if (e_ != cudaSuccess) {                                                \
            ni_set_error("%s:%d kernel launch: %s", __FILE__, __LINE__,         \
                         cudaGetErrorString(e_));                               \
            return NI_ERR_CUDA;                                                 \
        }                                                                       \
    } while (0)

static inline int ni_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// Number of SMs on the current device (cached).
int ni_num_sms();

// Launch counter: every kernel launched by this library bumps it (bench.py reports it as gpu_launches).
extern unsigned long long g_ni_launches;
#define NI_COUNT_LAUNCH(n) (g_ni_launches += (n))

__device__ __forceinline__ float ni_clamp01(float v) { return fminf(fmaxf(v, 0.f), 1.f); }

// Round-half-to-even (== tf.round / rintf) for |v| < 2^22 using the magic-number trick (two full-rate FADDs).
__device__ __forceinline__ float ni_round_he(float v) {
    const float magic = 12582912.f;  // 1.5 * 2^23
    return __fsub_rn(__fadd_rn(v, magic), magic);
}

__device__ __forceinline__ float4 ni_ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
