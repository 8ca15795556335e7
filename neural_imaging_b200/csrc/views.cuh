// Logical-tensor addressing shared by the convolution and layer kernels: channel pitch/offset (concat-free skip
// connections) and the block-major space_to_depth / depth_to_space mapping of TensorFlow.
#pragma once
#include "conv_desc.h"

struct TensorView {
    int H, W, C;       // logical dims
    int pitch, coff, mode;
};

__device__ __forceinline__ long long view_addr(const TensorView& v, int n, int y, int x, int c) {
    if (v.mode == NI_MODE_PLAIN) return (((long long)n * v.H + y) * v.W + x) * v.pitch + v.coff + c;
    const int F = v.C >> 2, blk = c / F, f = c - blk * F;
    return (((long long)n * 2 * v.H + 2 * y + (blk >> 1)) * (2 * v.W) + 2 * x + (blk & 1)) * v.pitch + v.coff + f;
}
