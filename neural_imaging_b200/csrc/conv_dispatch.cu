// Dispatch between the tcgen05 implicit-GEMM path (dense contractions) and the FP32 SIMT path (everything else:
// 3/4-channel layers, strided / mirrored-pad / space-to-depth addressing). Both run on the GPU; there is no CPU path.
// ni_conv2d_set_force_simt(1) forces the SIMT path (tests compare the two on the device); development builds (-DNI_DEV) also read
// NI_CONV_FORCE_SIMT=1 from the environment.
#include <stdlib.h>

#include "conv_desc.h"
#include "ni_common.cuh"

extern "C" int ni_conv2d_fprop_simt(const ni_conv_desc*, const float*, const float*, const float*, float*, cudaStream_t);
extern "C" int ni_conv2d_dgrad_simt(const ni_conv_desc*, const float*, const float*, float*, cudaStream_t);
extern "C" int ni_conv2d_wgrad_simt(const ni_conv_desc*, const float*, const float*, float*, cudaStream_t);
extern "C" int ni_conv2d_tc_supported(const ni_conv_desc*, int);
extern "C" int ni_conv2d_fprop_tc(const ni_conv_desc*, const float*, const float*, const float*, float*, cudaStream_t);
extern "C" int ni_conv2d_dgrad_tc(const ni_conv_desc*, const float*, const float*, float*, cudaStream_t);
extern "C" int ni_conv2d_wgrad_tc(const ni_conv_desc*, const float*, const float*, float*, cudaStream_t);
extern "C" int ni_weight_transpose_io(const float*, float*, int, int, int, cudaStream_t);
extern "C" int ni_conv2d_small_supported(const ni_conv_desc*, int);
extern "C" int ni_conv2d_fprop_small(const ni_conv_desc*, const float*, const float*, const float*, float*, cudaStream_t);
extern "C" int ni_conv2d_dgrad_small(const ni_conv_desc*, const float*, const float*, float*, cudaStream_t);
extern "C" int ni_conv2d_wgrad_small(const ni_conv_desc*, const float*, const float*, float*, cudaStream_t);
extern "C" int ni_conv2d_direct_supported(const ni_conv_desc*, int);
extern "C" int ni_conv2d_fprop_direct(const ni_conv_desc*, const float*, const float*, const float*, float*, cudaStream_t);
extern "C" int ni_conv2d_dgrad_direct(const ni_conv_desc*, const float*, const float*, float*, cudaStream_t);
extern "C" int ni_conv2d_wgrad_direct(const ni_conv_desc*, const float*, const float*, float*, cudaStream_t);
int ni_get_scratch2(size_t bytes, float** out);

static bool force_simt() {
    static int v = -1;
    if (v < 0) {
#ifdef NI_DEV
        const char* e = getenv("NI_CONV_FORCE_SIMT");
        v = (e && e[0] == '1') ? 1 : 0;
#else
        v = 0;
#endif
    }
    return v == 1;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

static int g_force_simt_override = -1;   // -1: follow the environment, 0 / 1: explicit
extern "C" void ni_conv2d_set_force_simt(int on) { g_force_simt_override = on; }
static bool use_simt() { return g_force_simt_override >= 0 ? g_force_simt_override == 1 : force_simt(); }

// Layers with 8 .. 28 channels on one side of a 32-multiple layer (U-Net 32 -> 12, DCN 64 -> 12): zero-padded tensor-core tiles beat the
// register-blocked FP32 kernels (fprop 1.54 -> 0.7 ms, dgrad 1.17 -> 0.7 ms for the U-Net output layer at 256 x 128 x 128)
static bool narrow_tc(const ni_conv_desc* d, int op) {
    if (!d) return false;
    const bool out_side = d->cout >= 8 && d->cout < 32 && d->cin % 32 == 0, in_side = d->cin >= 8 && d->cin < 32 && d->cout % 32 == 0;
    return (out_side || in_side) && ni_conv2d_tc_supported(d, op);
}

extern "C" int ni_conv2d_fprop(const ni_conv_desc* d, const float* x, const float* w, const float* bias, float* y, cudaStream_t st) {
    if (!use_simt() && narrow_tc(d, 0) && aligned16(x) && aligned16(y)) return ni_conv2d_fprop_tc(d, x, w, bias, y, st);
    if (!use_simt() && ni_conv2d_direct_supported(d, 0)) return ni_conv2d_fprop_direct(d, x, w, bias, y, st);
    if (!use_simt() && ni_conv2d_small_supported(d, 0)) return ni_conv2d_fprop_small(d, x, w, bias, y, st);
    if (!use_simt() && ni_conv2d_tc_supported(d, 0) && aligned16(x) && aligned16(y)) return ni_conv2d_fprop_tc(d, x, w, bias, y, st);
    return ni_conv2d_fprop_simt(d, x, w, bias, y, st);
}

// w: the layer's HWIO weights (kh, kw, cin, cout).
extern "C" int ni_conv2d_dgrad(const ni_conv_desc* d, const float* dy, const float* w, float* dx, cudaStream_t st) {
    NI_REQUIRE(d && w, "ni_conv2d_dgrad: null pointer");
    if (!use_simt() && narrow_tc(d, 1) && aligned16(dy) && aligned16(dx)) return ni_conv2d_dgrad_tc(d, dy, w, dx, st);
    if (!use_simt() && ni_conv2d_direct_supported(d, 1)) return ni_conv2d_dgrad_direct(d, dy, w, dx, st);
    if (!use_simt() && ni_conv2d_small_supported(d, 1)) return ni_conv2d_dgrad_small(d, dy, w, dx, st);
    if (!use_simt() && ni_conv2d_tc_supported(d, 1) && aligned16(dy) && aligned16(dx)) return ni_conv2d_dgrad_tc(d, dy, w, dx, st);
    float* wt = nullptr;
    const int taps = d->kh * d->kw;
    int rc = ni_get_scratch2(sizeof(float) * (size_t)taps * d->cin * d->cout, &wt);
    if (rc) return rc;
    rc = ni_weight_transpose_io(w, wt, taps, d->cin, d->cout, st);
    if (rc) return rc;
    return ni_conv2d_dgrad_simt(d, dy, wt, dx, st);
}

extern "C" int ni_conv2d_wgrad(const ni_conv_desc* d, const float* x, const float* dy, float* dw, cudaStream_t st) {
    if (!use_simt() && ni_conv2d_direct_supported(d, 2)) return ni_conv2d_wgrad_direct(d, x, dy, dw, st);
    if (!use_simt() && ni_conv2d_small_supported(d, 2)) return ni_conv2d_wgrad_small(d, x, dy, dw, st);
    if (!use_simt() && ni_conv2d_tc_supported(d, 2) && aligned16(x) && aligned16(dy)) return ni_conv2d_wgrad_tc(d, x, dy, dw, st);
    return ni_conv2d_wgrad_simt(d, x, dy, dw, st);
}
