// Shared-memory staging of an interleaved-RGB (C = 3) image tile with tf.pad semantics at the image border.
// Rows y0 - HALO .. y0 + TS - 1 + HALO, columns x0 - XOFF .. x0 - XOFF + COLS - 1 (XOFF and COLS multiples of 4, so that every 128-bit slot
// of a row segment is 16-byte aligned in global memory when W % 4 == 0). Each 128-bit slot that lies inside the image row is one
// vector load; only the slots that stick out of the image are filled element by element through the index map.
#pragma once
#include "ni_common.cuh"

enum { TILE_ZERO = 0, TILE_SYMMETRIC = 1, TILE_REFLECT = 2 };

template <int MODE>
__device__ __forceinline__ int tile_map(int u, int n) {
    if (MODE == TILE_SYMMETRIC) { if (u < 0) u = -u - 1; if (u >= n) u = 2 * n - 1 - u; }
    if (MODE == TILE_REFLECT) { if (u < 0) u = -u; if (u >= n) u = 2 * (n - 1) - u; }
    return u < 0 ? 0 : (u >= n ? n - 1 : u);       // positions further out than any halo in use are never read
}

template <int TS, int HALO, int XOFF, int COLS, int MODE, int THREADS>
__device__ __forceinline__ void load_tile3(float* tile, const float* __restrict__ img, int H, int W, int y0, int x0) {
    constexpr int ROWS = TS + 2 * HALO, RS = COLS * 3, V = RS / 4;
    static_assert(XOFF % 4 == 0 && COLS % 4 == 0, "128-bit slots");
    const int f_lo = (x0 - XOFF) * 3, row_floats = W * 3;
    const bool vec_ok = (W & 3) == 0;
    for (int t = threadIdx.x; t < ROWS * V; t += THREADS) {
        const int r = t / V, v = t - r * V;
        const int gy = y0 - HALO + r;
        const bool row_in = gy >= 0 && gy < H;
        float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
        const int f0 = f_lo + 4 * v;                                   // first float of this slot within the image row
        if (MODE != TILE_ZERO || row_in) {
            const int sy = MODE == TILE_ZERO ? gy : tile_map<MODE>(gy, H);
            const float* rowp = img + (size_t)sy * row_floats;
            if (vec_ok && f0 >= 0 && f0 + 4 <= row_floats) {
                val = ni_ldg4(rowp + f0);
            } else {
                float e[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int f = f0 + k;
                    // floor division by 3 for possibly negative f (f >= -3 * XOFF)
                    const int col = (f + 3 * XOFF) / 3 - XOFF, ch = f - col * 3;
                    if (MODE == TILE_ZERO) e[k] = (col >= 0 && col < W) ? __ldg(rowp + col * 3 + ch) : 0.f;
                    else e[k] = __ldg(rowp + tile_map<MODE>(col, W) * 3 + ch);
                }
                val = make_float4(e[0], e[1], e[2], e[3]);
            }
        }
        *reinterpret_cast<float4*>(tile + r * RS + 4 * v) = val;
    }
}
