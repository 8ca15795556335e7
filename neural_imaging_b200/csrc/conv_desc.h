// Convolution problem descriptor shared by the SIMT and tcgen05 implicit-GEMM kernels and mirrored (field for
// field) by ni_conv_desc in include/ni_b200.h and by the ctypes Structure in neural_imaging_b200/_lib.py.
#pragma once

#ifdef __cplusplus
extern "C" {
#endif

// Activations fused into the fprop epilogue (reference helpers/tf_helpers.py:22-28 activation_mapping).
enum { NI_ACT_NONE = 0, NI_ACT_LEAKY_RELU = 1, NI_ACT_RELU = 2, NI_ACT_TANH = 3, NI_ACT_SIGMOID = 4, NI_ACT_CLIP01 = 5 };

// Tensor addressing modes. D2S/S2D follow tf.nn.depth_to_space / space_to_depth block-major channel order
// ((di*2+dj)*C + c), so that Conv2DTranspose(2x2, stride 2) == 1x1 conv (Cin -> 4*Cout) + depth_to_space(2).
enum { NI_MODE_PLAIN = 0, NI_MODE_BLOCK2 = 1 };

// tf.pad modes folded into the convolution's input addressing.
enum { NI_PAD_ZERO = 0, NI_PAD_SYMMETRIC = 1, NI_PAD_REFLECT = 2 };

typedef struct ni_conv_desc {
    int n, h, w;             // logical input: batch, height, width
    int cin, cout;           // logical channels
    int kh, kw;              // filter taps
    int stride;              // 1 or 2 (both dims)
    int pad_t, pad_l;        // zero padding before (TF SAME: computed by the caller; VALID: 0)
    int oh, ow;              // logical output height / width
    int in_pitch, in_coff;   // physical channel pitch / first channel of the input buffer
    int out_pitch, out_coff; // same for the output buffer
    int in_mode;             // NI_MODE_BLOCK2: logical (n,h,w,cin) is read through space_to_depth(2) of a (n,2h,2w,cin/4) buffer
    int out_mode;            // NI_MODE_BLOCK2: logical (n,oh,ow,cout) is written through depth_to_space(2) into (n,2oh,2ow,cout/4)
    int act;                 // NI_ACT_* applied after the bias (fprop only)
    float act_alpha;         // leaky-relu slope
    int accumulate;          // fprop/dgrad: out += result (residual connections); wgrad: dw += result
    int bias_mod;            // > 0: bias index = co % bias_mod (transposed conv: one bias per real feature)
    int pad_mode;            // NI_PAD_ZERO, or NI_PAD_SYMMETRIC / NI_PAD_REFLECT index mapping (fprop and wgrad only; dgrad of a
                             // mirrored pad = zero-pad dgrad on the padded domain + ni_pad_fold)
} ni_conv_desc;

#ifdef __cplusplus
}
#endif
