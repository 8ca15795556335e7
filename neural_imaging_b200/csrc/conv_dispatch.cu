// Dispatch between the tcgen05 implicit-GEMM path (dense contractions) and the FP32 SIMT path (everything else).
#include "conv_desc.h"
#include "ni_common.cuh"

extern "C" int ni_conv2d_fprop_simt(const ni_conv_desc*, const float*, const float*, const float*, float*, cudaStream_t);
extern "C" int ni_conv2d_dgrad_simt(const ni_conv_desc*, const float*, const float*, float*, cudaStream_t);
extern "C" int ni_conv2d_wgrad_simt(const ni_conv_desc*, const float*, const float*, float*, cudaStream_t);

extern "C" int ni_conv2d_fprop(const ni_conv_desc* d, const float* x, const float* w, const float* bias, float* y, cudaStream_t st) {
    return ni_conv2d_fprop_simt(d, x, w, bias, y, st);
}
extern "C" int ni_conv2d_dgrad(const ni_conv_desc* d, const float* dy, const float* wt, float* dx, cudaStream_t st) {
    return ni_conv2d_dgrad_simt(d, dy, wt, dx, st);
}
extern "C" int ni_conv2d_wgrad(const ni_conv_desc* d, const float* x, const float* dy, float* dw, cudaStream_t st) {
    return ni_conv2d_wgrad_simt(d, x, dy, dw, st);
}
