// The manipulation stack of the joint training step with the 2x2 average pooling folded into the store.
//
// Reference: workflows/manipulation_classification.py:199-208 (run_manipulations: [x] + [op(x, s) ...] concatenated class-major) followed
// by :231-245 (run_downsampling 'pool:2' = tf.nn.avg_pool 2x2 stride 2), with the operators of helpers/tf_helpers.py:68-76 (resample,
// factor 50), :113-125 (gaussian, REFLECT pad), :156-184 (sharpen, SYMMETRIC pad, HSV, S-channel tap [2,2]).
//
// Unfused, the full-resolution stack m = (5B, H, W, 3) is written by four kernels, read back by the pooling kernel, and in the backward
// pass the equally large dm is written by the pooling backward and re-read by every manipulation backward (1 GB each way at B = 256).
// Here ONE forward kernel reads a shared-memory tile of the RGB image once (128-bit loads, halo 2) and writes the POOLED native / sharpen /
// resample / gaussian slots directly, and ONE backward kernel reads the pooled gradient slots and accumulates into dY. m and dm never
// exist. The only thing saved for the backward pass is the Gaussian clip mask (1 byte per pixel).
// The arithmetic of every operator follows the stand-alone kernels of manip.cu operation for operation (same tap order, same
// interpolation formulas), so the fused result equals avg_pool(manipulation(x)) to float32 rounding.
#include "ni_common.cuh"
#include "tile3.cuh"

namespace {

constexpr int kTS = 32;                    // full-resolution tile edge (16 x 16 pooled pixels, one per thread)
constexpr int kThreads = 256;

struct StackDesc {
    int s_native, s_sharpen, s_resample, s_gauss;   // class slot of each operator in the stack, -1 = absent
    int gk;                                         // gaussian kernel size: 3 or 5
    float sharp[9];
    float gauss[25];
};

__device__ __forceinline__ int reflect_i(int u, int n) {
    if (u < 0) u = -u;
    if (u >= n) u = 2 * (n - 1) - u;
    return u;
}
__device__ __forceinline__ int symm_i(int u, int n) {
    if (u < 0) u = -u - 1;
    if (u >= n) u = 2 * n - 1 - u;
    return u;
}
__device__ __forceinline__ int clamp_i(int u, int n) { return u < 0 ? 0 : (u >= n ? n - 1 : u); }

__device__ __forceinline__ void rgb2hsv(float r, float g, float b, float& h, float& s, float& v) {
    v = fmaxf(r, fmaxf(g, b));
    const float range = v - fminf(r, fminf(g, b));
    s = v > 0.f ? range / v : 0.f;
    const float norm = (1.f / range) * (1.f / 6.f);
    float hh;
    if (r == v) hh = norm * (g - b);
    else if (g == v) hh = norm * (b - r) + 2.f / 6.f;
    else hh = norm * (r - g) + 4.f / 6.f;
    if (!(range > 0.f)) hh = 0.f;
    if (hh < 0.f) hh += 1.f;
    h = hh;
}
__device__ __forceinline__ void hsv2rgb(float h, float s, float v, float& r, float& g, float& b) {
    const float dh = h * 6.f;
    const float dr = fminf(fmaxf(fabsf(dh - 3.f) - 1.f, 0.f), 1.f);
    const float dg = fminf(fmaxf(-fabsf(dh - 2.f) + 2.f, 0.f), 1.f);
    const float db = fminf(fmaxf(-fabsf(dh - 4.f) + 2.f, 0.f), 1.f);
    const float one_s = -s + 1.f;
    r = (one_s + s * dr) * v;
    g = (one_s + s * dg) * v;
    b = (one_s + s * db) * v;
}

// ------------------------------------------------------------------------------------------------------------ forward
constexpr int kFH = 2, kFX = 4, kFC = 40;                     // forward tile: halo 2, 40 columns
constexpr int kFRS = kFC * 3;
constexpr int kHsvN = kTS + 2;                                // HSV region: halo 1
constexpr int kSmN = kTS / 2 + 2;                             // half-resolution native image: halo 1

__global__ void __launch_bounds__(kThreads)
manip_stack_pool2_fwd_kernel(const float* __restrict__ Y, float* __restrict__ c, unsigned char* __restrict__ mask, int B, int H, int W,
                             StackDesc d) {
    __shared__ __align__(16) float tile[(kTS + 2 * kFH) * kFRS];
    __shared__ float hsv[kHsvN * kHsvN * 3];
    __shared__ float small[kSmN * kSmN * 3];
    const int n = blockIdx.z, y0 = blockIdx.y * kTS, x0 = blockIdx.x * kTS;
    const float* img = Y + (size_t)n * H * W * 3;
    load_tile3<kTS, kFH, kFX, kFC, TILE_REFLECT, kThreads>(tile, img, H, W, y0, x0);      // REFLECT halo (gaussian); sharpen / resample map their own
    __syncthreads();
    const int H2 = H >> 1, W2 = W >> 1;
    // HSV of the tile + halo 1 under SYMMETRIC padding (the source pixel of a padded position lies inside the image, hence inside the tile)
    if (d.s_sharpen >= 0) {
        for (int t = threadIdx.x; t < kHsvN * kHsvN; t += kThreads) {
            const int i = t / kHsvN, j = t - i * kHsvN;
            const int sy = symm_i(y0 - 1 + i, H), sx = symm_i(x0 - 1 + j, W);
            const int ly = min(max(sy - (y0 - kFH), 0), kTS + 2 * kFH - 1), lx = min(max(sx - (x0 - kFX), 0), kFC - 1);
            const float* p = tile + ly * kFRS + lx * 3;
            float h, s, v;
            rgb2hsv(p[0], p[1], p[2], h, s, v);
            hsv[t * 3] = h; hsv[t * 3 + 1] = s; hsv[t * 3 + 2] = v;
        }
    }
    // half-resolution image (the down-sampling half of resample@50 = what bilinear resize by 1/2 computes), edge-clamped halo 1
    if (d.s_resample >= 0) {
        for (int t = threadIdx.x; t < kSmN * kSmN; t += kThreads) {
            const int i = t / kSmN, j = t - i * kSmN;
            const int jy = clamp_i((y0 >> 1) - 1 + i, H2), jx = clamp_i((x0 >> 1) - 1 + j, W2);
            const int ly = min(max(2 * jy - (y0 - kFH), 0), kTS + 2 * kFH - 2), lx = min(max(2 * jx - (x0 - kFX), 0), kFC - 2);
            const float* p = tile + ly * kFRS + lx * 3;
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {      // resize_fwd_kernel with lerp 1/2 in both directions
                const float tl = p[ch], tr = p[3 + ch], bl = p[kFRS + ch], br = p[kFRS + 3 + ch];
                const float top = tl + (tr - tl) * 0.5f;
                const float bot = bl + (br - bl) * 0.5f;
                small[t * 3 + ch] = top + (bot - top) * 0.5f;
            }
        }
    }
    __syncthreads();
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    const int oy = (y0 >> 1) + ty, ox = (x0 >> 1) + tx;          // pooled pixel of this thread
    if (oy >= H2 || ox >= W2) return;
    const size_t slot = (size_t)B * H2 * W2 * 3;
    float* out = c + (((size_t)n * H2 + oy) * W2 + ox) * 3;
    const int ly0 = 2 * ty + kFH, lx0 = 2 * tx + kFX;            // top-left pixel of the quad in tile coordinates

    if (d.s_native >= 0) {                                       // avgpool_fwd_kernel: sum in (a, b) order, times 1/4
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                const float* p = tile + (ly0 + a) * kFRS + (lx0 + b) * 3;
                a0 += p[0]; a1 += p[1]; a2 += p[2];
            }
        float* o = out + d.s_native * slot;
        o[0] = a0 * 0.25f; o[1] = a1 * 0.25f; o[2] = a2 * 0.25f;
    }
    if (d.s_gauss >= 0) {
        // the four outputs of the quad share a (k + 1) x (k + 1) window: it is streamed row by row through registers (64-bit shared
        // loads) and every output accumulates its taps in the stand-alone kernel's order (a ascending, b ascending)
        const int k = d.gk;
        float acc[2][2][3];
#pragma unroll
        for (int qa = 0; qa < 2; ++qa)
#pragma unroll
            for (int qb = 0; qb < 2; ++qb) acc[qa][qb][0] = acc[qa][qb][1] = acc[qa][qb][2] = 0.f;
        if (k == 5) {
#pragma unroll
            for (int r = 0; r < 6; ++r) {
                float row[18];
                const float2* p = reinterpret_cast<const float2*>(tile + (ly0 + r - 2) * kFRS + (lx0 - 2) * 3);      // even float offset
#pragma unroll
                for (int v = 0; v < 9; ++v) { const float2 t2 = p[v]; row[2 * v] = t2.x; row[2 * v + 1] = t2.y; }
#pragma unroll
                for (int qa = 0; qa < 2; ++qa) {
                    const int a = r - qa;
                    if (a < 0 || a > 4) continue;
#pragma unroll
                    for (int qb = 0; qb < 2; ++qb)
#pragma unroll
                        for (int b = 0; b < 5; ++b) {
                            const float w = d.gauss[a * 5 + b];
                            acc[qa][qb][0] = fmaf(w, row[(qb + b) * 3], acc[qa][qb][0]);
                            acc[qa][qb][1] = fmaf(w, row[(qb + b) * 3 + 1], acc[qa][qb][1]);
                            acc[qa][qb][2] = fmaf(w, row[(qb + b) * 3 + 2], acc[qa][qb][2]);
                        }
                }
            }
        } else {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                float row[12];
                const float* p = tile + (ly0 + r - 1) * kFRS + (lx0 - 1) * 3;
#pragma unroll
                for (int v = 0; v < 12; ++v) row[v] = p[v];
#pragma unroll
                for (int qa = 0; qa < 2; ++qa) {
                    const int a = r - qa;
                    if (a < 0 || a > 2) continue;
#pragma unroll
                    for (int qb = 0; qb < 2; ++qb)
#pragma unroll
                        for (int b = 0; b < 3; ++b) {
                            const float w = d.gauss[a * 3 + b];
                            acc[qa][qb][0] = fmaf(w, row[(qb + b) * 3], acc[qa][qb][0]);
                            acc[qa][qb][1] = fmaf(w, row[(qb + b) * 3 + 1], acc[qa][qb][1]);
                            acc[qa][qb][2] = fmaf(w, row[(qb + b) * 3 + 2], acc[qa][qb][2]);
                        }
                }
            }
        }
        float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int qa = 0; qa < 2; ++qa)
#pragma unroll
            for (int qb = 0; qb < 2; ++qb) {
                const float a0 = acc[qa][qb][0], a1 = acc[qa][qb][1], a2 = acc[qa][qb][2];
                if (mask)       // which channels pass the clip: the backward pass reads this instead of recomputing the blur
                    mask[((size_t)n * H + 2 * oy + qa) * W + 2 * ox + qb] =
                        (unsigned char)(((a0 >= 0.f && a0 <= 1.f) ? 1 : 0) | ((a1 >= 0.f && a1 <= 1.f) ? 2 : 0) | ((a2 >= 0.f && a2 <= 1.f) ? 4 : 0));
                s0 += ni_clamp01(a0); s1 += ni_clamp01(a1); s2 += ni_clamp01(a2);
            }
        float* o = out + d.s_gauss * slot;
        o[0] = s0 * 0.25f; o[1] = s1 * 0.25f; o[2] = s2 * 0.25f;
    }
    if (d.s_sharpen >= 0) {
        // 4 x 4 window of (H, V) shared by the quad; S comes from tap [2,2] of each output
        float hw[4][4], vw[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float* p = hsv + ((2 * ty + r) * kHsvN + (2 * tx + q)) * 3;
                hw[r][q] = p[0]; vw[r][q] = p[2];
            }
        float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int qa = 0; qa < 2; ++qa)
#pragma unroll
            for (int qb = 0; qb < 2; ++qb) {
                float ah = 0.f, av = 0.f;
#pragma unroll
                for (int a = 0; a < 3; ++a)
#pragma unroll
                    for (int b = 0; b < 3; ++b) {
                        ah = fmaf(d.sharp[a * 3 + b], hw[qa + a][qb + b], ah);
                        av = fmaf(d.sharp[a * 3 + b], vw[qa + a][qb + b], av);
                    }
                const float as = hsv[((2 * ty + qa + 2) * kHsvN + (2 * tx + qb + 2)) * 3 + 1];
                float r, g, bb;
                hsv2rgb(ah, as, av, r, g, bb);
                s0 += ni_clamp01(r); s1 += ni_clamp01(g); s2 += ni_clamp01(bb);
            }
        float* o = out + d.s_sharpen * slot;
        o[0] = s0 * 0.25f; o[1] = s1 * 0.25f; o[2] = s2 * 0.25f;
    }
    if (d.s_resample >= 0) {
        // up-sampling half of resample@50 (resize_fwd_kernel, scale 1/2): even output 2j reads small[j-1], small[j] with lerp 3/4,
        // odd output 2j+1 reads small[j], small[j+1] with lerp 1/4 (neighbours edge-clamped), then the 2x2 mean
        const int sy = ty + 1, sx = tx + 1;
        float s[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int qa = 0; qa < 2; ++qa)
#pragma unroll
            for (int qb = 0; qb < 2; ++qb) {
                const int ya = qa ? sy : sy - 1, yb = qa ? sy + 1 : sy;
                const int xa = qb ? sx : sx - 1, xb = qb ? sx + 1 : sx;
                const float ly = qa ? 0.25f : 0.75f, lx = qb ? 0.25f : 0.75f;
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    const float tl = small[(ya * kSmN + xa) * 3 + ch], tr = small[(ya * kSmN + xb) * 3 + ch];
                    const float bl = small[(yb * kSmN + xa) * 3 + ch], br = small[(yb * kSmN + xb) * 3 + ch];
                    const float top = tl + (tr - tl) * lx;
                    const float bot = bl + (br - bl) * lx;
                    s[ch] += top + (bot - top) * ly;
                }
            }
        float* o = out + d.s_resample * slot;
        o[0] = s[0] * 0.25f; o[1] = s[1] * 0.25f; o[2] = s[2] * 0.25f;
    }
}

// ------------------------------------------------------------------------------------------------------------ backward
constexpr int kGN = kTS + 4;                                  // masked upstream gradient of the gaussian slot: halo 2
constexpr int kGRS = kGN * 3 + 2;                             // row stride (even: 64-bit row loads stay aligned)

// transposed weights of the (up-sample by 2, then 2x2 mean) chain of resample@50 along one axis:
// d small[j] = T(j, j-1) g[j-1] + T(j, j) g[j] + T(j, j+1) g[j+1], g = pooled upstream gradient
__device__ __forceinline__ float resample_T(int j, int k, int n) {
    if (k < 0 || k >= n) return 0.f;
    if (k == j) return (j == 0 || j == n - 1) ? 1.75f : 1.5f;
    return 0.25f;
}

__global__ void __launch_bounds__(kThreads)
manip_stack_pool2_bwd_kernel(const unsigned char* __restrict__ mask, const float* __restrict__ dc, float* __restrict__ dY, int B, int H, int W,
                             StackDesc d, int accumulate) {
    __shared__ __align__(16) float gm[kGN * kGRS];
    const int n = blockIdx.z, y0 = blockIdx.y * kTS, x0 = blockIdx.x * kTS;
    const int H2 = H >> 1, W2 = W >> 1;
    const size_t slot = (size_t)B * H2 * W2 * 3;
    const float* dcn = dc + (size_t)n * H2 * W2 * 3;
    const int k = d.gk, pad = k >> 1;
    if (d.s_gauss >= 0) {
        // masked upstream gradient G[q] = [0 <= gauss(Y)[q] <= 1] * dc_gauss[q >> 1] / 4 for q in the tile + halo 2 (0 outside the image);
        // the clip mask was written by the forward kernel (1 byte per pixel)
        const float* dg = dcn + d.s_gauss * slot;
        const unsigned char* mk = mask + (size_t)n * H * W;
        for (int t = threadIdx.x; t < kGN * kGN; t += kThreads) {
            const int i = t / kGN, j = t - i * kGN;
            const int qy = y0 - 2 + i, qx = x0 - 2 + j;
            float g0 = 0.f, g1 = 0.f, g2 = 0.f;
            if (qy >= 0 && qy < H && qx >= 0 && qx < W) {
                const unsigned m = mk[(size_t)qy * W + qx];
                const float* u = dg + ((size_t)(qy >> 1) * W2 + (qx >> 1)) * 3;
                g0 = (m & 1) ? 0.25f * __ldg(u) : 0.f;
                g1 = (m & 2) ? 0.25f * __ldg(u + 1) : 0.f;
                g2 = (m & 4) ? 0.25f * __ldg(u + 2) : 0.f;
            }
            float* o = gm + i * kGRS + j * 3;
            o[0] = g0; o[1] = g1; o[2] = g2;
        }
        __syncthreads();
    }
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    const int oy = (y0 >> 1) + ty, ox = (x0 >> 1) + tx;
    if (oy >= H2 || ox >= W2) return;
    // contributions shared by the four pixels of the quad: native slot and resample@50
    float q0 = 0.f, q1 = 0.f, q2 = 0.f;
    if (d.s_native >= 0) {
        const float* u = dcn + d.s_native * slot + ((size_t)oy * W2 + ox) * 3;
        q0 = 0.25f * __ldg(u); q1 = 0.25f * __ldg(u + 1); q2 = 0.25f * __ldg(u + 2);
    }
    if (d.s_resample >= 0) {
        const float* dr = dcn + d.s_resample * slot;
        float r0 = 0.f, r1 = 0.f, r2 = 0.f;
#pragma unroll
        for (int a = -1; a <= 1; ++a) {
            const float wy = resample_T(oy, oy + a, H2);
            if (wy == 0.f) continue;
#pragma unroll
            for (int b = -1; b <= 1; ++b) {
                const float w = wy * resample_T(ox, ox + b, W2);
                if (w == 0.f) continue;
                const float* u = dr + ((size_t)(oy + a) * W2 + (ox + b)) * 3;
                r0 = fmaf(w, __ldg(u), r0); r1 = fmaf(w, __ldg(u + 1), r1); r2 = fmaf(w, __ldg(u + 2), r2);
            }
        }
        // 1/4 (pooling) on the way up and 1/4 (box down-sampling) on the way down
        q0 = fmaf(0.0625f, r0, q0); q1 = fmaf(0.0625f, r1, q1); q2 = fmaf(0.0625f, r2, q2);
    }
    float acc[2][2][3];
#pragma unroll
    for (int qa = 0; qa < 2; ++qa)
#pragma unroll
        for (int qb = 0; qb < 2; ++qb) { acc[qa][qb][0] = q0; acc[qa][qb][1] = q1; acc[qa][qb][2] = q2; }
    if (d.s_gauss >= 0) {
        const bool border = (y0 < pad) || (x0 < pad) || (y0 + kTS + pad > H) || (x0 + kTS + pad > W);
        if (!border && k == 5) {
            // transpose of pad + correlation away from the image edge: dx[p] = sum_{a,b} w[a,b] G[p + 2 - (a,b)]; the quad shares a 6 x 6
            // window of G, streamed row by row (window row r holds G row py0 - 2 + r: output qa uses it with a = 4 + qa - r)
#pragma unroll
            for (int r = 0; r < 6; ++r) {
                float row[18];
                const float2* p = reinterpret_cast<const float2*>(gm + (2 * ty + r) * kGRS + (2 * tx) * 3);
#pragma unroll
                for (int v = 0; v < 9; ++v) { const float2 t2 = p[v]; row[2 * v] = t2.x; row[2 * v + 1] = t2.y; }
#pragma unroll
                for (int qa = 0; qa < 2; ++qa) {
                    const int a = 4 + qa - r;
                    if (a < 0 || a > 4) continue;
#pragma unroll
                    for (int qb = 0; qb < 2; ++qb)
#pragma unroll
                        for (int b = 0; b < 5; ++b) {
                            const float w = d.gauss[a * 5 + b];
                            const int col = 4 + qb - b;                 // window column of G[px + 2 - b]
                            acc[qa][qb][0] = fmaf(w, row[col * 3], acc[qa][qb][0]);
                            acc[qa][qb][1] = fmaf(w, row[col * 3 + 1], acc[qa][qb][1]);
                            acc[qa][qb][2] = fmaf(w, row[col * 3 + 2], acc[qa][qb][2]);
                        }
                }
            }
        } else {
#pragma unroll
            for (int qa = 0; qa < 2; ++qa)
#pragma unroll
                for (int qb = 0; qb < 2; ++qb) {
                    const int py = 2 * oy + qa, px = 2 * ox + qb;
                    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
                    // all padded coordinates u that REFLECT maps onto p (gaussian_bwd_kernel of manip.cu)
                    int uy[3], ux[3], ny = 0, nx = 0;
                    uy[ny++] = py;
                    if (py >= 1 && py <= pad) uy[ny++] = -py;
                    if (py <= H - 2 && 2 * (H - 1) - py <= H - 1 + pad) uy[ny++] = 2 * (H - 1) - py;
                    ux[nx++] = px;
                    if (px >= 1 && px <= pad) ux[nx++] = -px;
                    if (px <= W - 2 && 2 * (W - 1) - px <= W - 1 + pad) ux[nx++] = 2 * (W - 1) - px;
                    for (int iy = 0; iy < ny; ++iy)
                        for (int a = 0; a < k; ++a) {
                            const int qy = uy[iy] - (a - pad);
                            if (qy < 0 || qy >= H) continue;
                            for (int ix = 0; ix < nx; ++ix)
                                for (int b = 0; b < k; ++b) {
                                    const int qx = ux[ix] - (b - pad);
                                    if (qx < 0 || qx >= W) continue;
                                    const float* g = gm + (qy - y0 + 2) * kGRS + (qx - x0 + 2) * 3;
                                    const float w = d.gauss[a * k + b];
                                    a0 = fmaf(w, g[0], a0); a1 = fmaf(w, g[1], a1); a2 = fmaf(w, g[2], a2);
                                }
                        }
                    acc[qa][qb][0] += a0; acc[qa][qb][1] += a1; acc[qa][qb][2] += a2;
                }
        }
    }
#pragma unroll
    for (int qa = 0; qa < 2; ++qa) {
        // the two pixels of a quad row are 6 consecutive floats, 8-byte aligned (even x): three 64-bit accesses
        float2* o = reinterpret_cast<float2*>(dY + (((size_t)n * H + 2 * oy + qa) * W + 2 * ox) * 3);
        float v[6] = {acc[qa][0][0], acc[qa][0][1], acc[qa][0][2], acc[qa][1][0], acc[qa][1][1], acc[qa][1][2]};
        if (accumulate) {
            const float2 v0 = o[0], v1 = o[1], v2 = o[2];
            v[0] += v0.x; v[1] += v0.y; v[2] += v1.x; v[3] += v1.y; v[4] += v2.x; v[5] += v2.y;
        }
        o[0] = make_float2(v[0], v[1]); o[1] = make_float2(v[2], v[3]); o[2] = make_float2(v[4], v[5]);
    }
}

int check_desc(const char* who, int n_classes, const int* slots, int gk, const float* sharp9, const float* gauss, StackDesc& d) {
    NI_REQUIRE(slots && n_classes >= 1, "%s: null slot table", who);
    d.s_native = slots[0]; d.s_sharpen = slots[1]; d.s_resample = slots[2]; d.s_gauss = slots[3];
    for (int i = 0; i < 4; ++i) NI_REQUIRE(slots[i] >= -1 && slots[i] < n_classes, "%s: slot %d out of range", who, slots[i]);
    d.gk = 1;
    if (d.s_gauss >= 0) {
        NI_REQUIRE(gauss && (gk == 3 || gk == 5), "%s: fused gaussian needs a 3x3 or 5x5 filter, got %d", who, gk);
        d.gk = gk;
        for (int i = 0; i < gk * gk; ++i) d.gauss[i] = gauss[i];
    }
    if (d.s_sharpen >= 0) {
        NI_REQUIRE(sharp9, "%s: sharpen slot without a filter", who);
        for (int i = 0; i < 9; ++i) d.sharp[i] = sharp9[i];
    }
    return NI_OK;
}

}  // namespace

// slots[4]: class slot of {native, sharpen, resample@50, gaussian} in the pooled stack c = (n_classes * b, h/2, w/2, 3), -1 = not handled
// here (the caller fills other slots — jpeg, awgn, gamma, median, other resampling factors — through the stand-alone kernels + ni_avgpool).
extern "C" int ni_manip_stack_pool2_fwd(const float* y, float* c, unsigned char* mask, int b, int h, int w, int n_classes, const int* slots,
                                        const float* sharp9, const float* gauss, int gk, cudaStream_t st) {
    NI_REQUIRE(y && c, "ni_manip_stack_pool2_fwd: null pointer");
    NI_REQUIRE(b >= 0 && h >= 8 && w >= 8 && (h % 2) == 0 && (w % 2) == 0, "ni_manip_stack_pool2_fwd: H, W must be even and >= 8, got %d x %d", h, w);
    StackDesc d;
    int rc = check_desc("ni_manip_stack_pool2_fwd", n_classes, slots, gk, sharp9, gauss, d);
    if (rc) return rc;
    if (b == 0) return NI_OK;
    dim3 grid(ni_cdiv(w, kTS), ni_cdiv(h, kTS), b);
    manip_stack_pool2_fwd_kernel<<<grid, kThreads, 0, st>>>(y, c, mask, b, h, w, d);
    NI_LAUNCH_CHECK(); NI_COUNT_LAUNCH(1);
    return NI_OK;
}

// dY (+)= sum over the handled slots of d(pooled slot)/dY applied to dc. sharpen carries no gradient (TF 2.1: RGBToHSV / HSVToRGB are
// NotDifferentiable), so its slot is ignored here.
extern "C" int ni_manip_stack_pool2_bwd(const unsigned char* mask, const float* dc, float* dy, int b, int h, int w, int n_classes,
                                        const int* slots, const float* gauss, int gk, int accumulate, cudaStream_t st) {
    NI_REQUIRE(dc && dy, "ni_manip_stack_pool2_bwd: null pointer");
    NI_REQUIRE(b >= 0 && h >= 8 && w >= 8 && (h % 2) == 0 && (w % 2) == 0, "ni_manip_stack_pool2_bwd: H, W must be even and >= 8, got %d x %d", h, w);
    NI_REQUIRE(slots, "ni_manip_stack_pool2_bwd: null slot table");
    const int bslots[4] = {slots[0], -1, slots[2], slots[3]};          // no gradient through the sharpen slot
    StackDesc d;
    int rc = check_desc("ni_manip_stack_pool2_bwd", n_classes, bslots, gk, nullptr, gauss, d);
    if (rc) return rc;
    NI_REQUIRE(d.s_gauss < 0 || mask, "ni_manip_stack_pool2_bwd: the gaussian slot needs the clip mask written by the forward pass");
    if (b == 0) return NI_OK;
    dim3 grid(ni_cdiv(w, kTS), ni_cdiv(h, kTS), b);
    manip_stack_pool2_bwd_kernel<<<grid, kThreads, 0, st>>>(mask, dc, dy, b, h, w, d, accumulate);
    NI_LAUNCH_CHECK(); NI_COUNT_LAUNCH(1);
    return NI_OK;
}
