/*
 * ni_b200 — C-ABI of the B200-native neural-imaging hot path (libni_b200.so).
 *
 * The reference (pkorus/neural-imaging) has NO FFI on this path: every operator below is a chain of TensorFlow ops
 * issued from Python (models/ (all modules), helpers/tf_helpers.py, workflows/manipulation_classification.py). Each entry point
 * therefore cites the reference Python code it replaces; INTEGRATION.md shows the ctypes stub a maintainer adds.
 *
 * Conventions
 *   - every function returns 0 on success, a negative NI_ERR_* code otherwise; ni_last_error() has the message;
 *   - all tensor pointers are DEVICE pointers to float32 NHWC data owned by the caller; no hidden allocation;
 *   - `stream` is a cudaStream_t (passed as void* so that this header needs no CUDA include);
 *   - functions are asynchronous w.r.t. the host (ordered on `stream`), except where stated;
 *   - build target: sm_100a only. There is no CPU fallback.
 */
#ifndef NI_B200_H
#define NI_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef void* ni_stream_t;

#define NI_OK 0
#define NI_ERR_ARG (-1)
#define NI_ERR_CUDA (-2)
#define NI_ERR_UNSUPPORTED (-3)

const char* ni_last_error(void);
int ni_version(void);
int ni_device_arch(void);                    /* compute capability major*10+minor of the current device */
unsigned long long ni_launch_count(void);    /* kernels launched by this library since the last reset */
void ni_reset_launch_count(void);

/* ------------------------------------------------------------------------------------------------ differentiable JPEG
 * Replaces DifferentiableJPEG.call (models/jpeg.py:91-159) incl. Quantization (models/layers.py:118-136) in ONE kernel.
 * x, y: (n,h,w,3) in [0,1]; h, w multiples of 8. q_luma/q_chroma: HOST pointers to 64 floats (jpeg_qtable,
 * compression/jpeg_helpers.py:264-305), row-major [k][l]. mode: 0 'soft', 1 'sin', 2 'harmonic'.
 * x_deq (optional, may be NULL): de-quantised coefficients (3*n*(h/8)*(w/8), 8, 8) in the reference block order
 * ((n*3+c)*nb + by*(w/8)+bx)  — the second return value of DifferentiableJPEG.call (models/jpeg.py:159).
 * Algorithmic HBM bytes: 24 B/pixel forward, 36 B/pixel backward (X/Q recomputed from x, nothing saved). */
int ni_djpeg_fwd(const float* x, float* y, float* x_deq, int n, int h, int w, const float* q_luma, const float* q_chroma,
                 int mode, ni_stream_t stream);
int ni_djpeg_bwd(const float* x, const float* dy, float* dx, int n, int h, int w, const float* q_luma,
                 const float* q_chroma, int mode, ni_stream_t stream);
/* gradient w.r.t. the two 8x8 quantisation tables of DifferentiableJPEG(trainable=True) (models/jpeg.py:58-62): dq = [luma 64 | chroma 64] */
int ni_djpeg_bwd_tables(const float* x, const float* dy, float* dq, int n, int h, int w, const float* q_luma, const float* q_chroma,
                        int mode, ni_stream_t stream);

/* ------------------------------------------------------------------------------------------------ training-data feed
 * Replaces the host-side batch assembly of Dataset.next_training_batch (helpers/dataset.py:89-131): crop, astype(float) / 65535
 * (uint16 RGGB stacks) or / 255 (uint8 RGB), float32 batch. Results are bit-identical to the reference's float32 batches.
 * ni_feed_convert: dst[i] = float(src[i]) / denom for n contiguous elements (src_bytes 1 = uint8, 2 = uint16; 16-byte aligned).
 * ni_feed_gather : images (n_images,h,w,c) integers RESIDENT on the device; coords (batch,3) int32 DEVICE triples (image, y, x);
 *                  out (batch,ph,pw,c) float32 = images[image, y:y+ph, x:x+pw, :] / denom. */
int ni_feed_convert(const void* src, int src_bytes, float* dst, long long n, float denom, ni_stream_t stream);
int ni_feed_gather(const void* images, int src_bytes, int n_images, int h, int w, int c, const int* coords, int batch, int ph, int pw,
                   float denom, float* out, ni_stream_t stream);

/* ------------------------------------------------------------------------------------------------ metrics
 * Structural similarity, one value per image: mean over the VALID region and the channels of the SSIM map built from a separable
 * k-tap window (HOST pointer, k <= 15). Replaces tf.image.ssim(a, b, 1.0) (models/compression.py:89, helpers/tf_helpers.py:39-40):
 * Gaussian 11 taps sigma 1.5, cov_norm 1, c1 = 1e-4, c2 = 9e-4; and helpers/metrics.py:9-26 (skimage structural_similarity,
 * multichannel, data_range 1): uniform 7 taps, cov_norm 49/48, same constants. a, b: (n,h,w,c); out_n: n floats. */
int ni_ssim(const float* a, const float* b, float* out_n, int n, int h, int w, int c, const float* win_host, int k, float cov_norm, float c1,
            float c2, ni_stream_t stream);
/* SSIM / MS-SSIM as image LOSSES of the ISPs (NIPModel.construct_loss, models/pipelines.py:53-63 -> helpers/tf_helpers.py:39-44:
 * mean(255 (1 - tf.image.ssim(a, b, 1))), mean(255 (1 - tf.image.ssim_multiscale(a, b, 1)))), forward and backward w.r.t. a.
 * ni_ssim_stats    : stats_nc2[(img*c+ch)*2 + {0,1}] = mean over the VALID region of {luminance*cs, cs} (TF's _ssim_per_channel).
 * ni_msssim_combine: stats / coef are [levels][n*c][2]; levels = 1 -> SSIM loss, levels = 5 with the TF power factors (HOST pointer)
 *                    -> MS-SSIM (relu of cs at scales 0..3 and of ssim at the last scale, weighted geometric mean, mean over
 *                    channels and images). Adds loss * loss_scale to *loss_acc; coef = grad_scale * d loss / d stats.
 * ni_ssim_bwd      : da (+)= d/da of sum_{img,ch} coef[.][0] * mean(lum*cs) + coef[.][1] * mean(cs), fused in shared memory. */
int ni_ssim_stats(const float* a, const float* b, float* stats_nc2, int n, int h, int w, int c, const float* win_host, int k, float cov_norm,
                  float c1, float c2, ni_stream_t stream);
int ni_ssim_bwd(const float* a, const float* b, const float* coef_nc2, float* da, int accumulate, int n, int h, int w, int c,
                const float* win_host, int k, float cov_norm, float c1, float c2, ni_stream_t stream);
int ni_msssim_combine(const float* stats, float* coef, float* loss_acc, int n, int c, int levels, const float* weights_host, float loss_scale,
                      float grad_scale, ni_stream_t stream);

/* ------------------------------------------------------------------------------------------------ l3ic bit-stream codec
 * Batched, bit-exact replacements of the reference's only native component: pyfse (pyfse/pyfse.pyx:24-72 over the vendored FSE
 * library, fse_compress.c:648-714 / fse_decompress.c:262-302) and of the host loop around it (compression/codec.py:87-265).
 * All pointers are DEVICE pointers; one warp codes one stream, so throughput comes from the number of streams per launch.
 *
 * ni_fse_compress_batch  : pyfse.compress of n byte strings; string i = src + i*src_stride, src_len[i] bytes; the result goes to
 *                          dst + i*dst_stride (dst_stride >= src_len[i] is enough). dst_len[i] = coded size (> 1), 0 = not
 *                          compressible (FSENotCompressibleError), 1 = one repeated byte (FSESymbolRepetitionError), < 0 = FSEException.
 * ni_fse_decompress_batch: pyfse.decompress(src_i, max_length = dst_cap) into rows of dst_stride >= dst_cap bytes; dst_len[i] = decoded
 *                          size or < 0 (FSEException).
 * ni_l3ic_encode         : codec.compress for a batch. latent (n,h,w,c) float32; codebook n_codes <= 256 floats; scratch: indices
 *                          n*c*h*w bytes, layer_bytes n*c*layer_slot bytes (layer_slot >= h*w), layer_len n*c ints. Output: stream i at
 *                          streams + i*stream_stride (stream_stride >= 5 + 2c + c*h*w), stream_len[i] bytes; status[i] != 0 flags the
 *                          conditions under which the reference raises (L3ICError / uncaught pyfse errors).
 * ni_l3ic_decode         : codec.decompress up to the latent: streams as above -> latent (n,h,w,c) = codebook[decoded indices];
 *                          scratch layer_off / layer_len n*c ints, indices n*c*index_slot bytes (index_slot >= h*w + 4);
 *                          status[i] != 0: corrupt / truncated stream, shape mismatch, symbol outside the code book. */
int ni_fse_compress_batch(const unsigned char* src, long long src_stride, const int* src_len, unsigned char* dst, long long dst_stride,
                          int* dst_len, int n, ni_stream_t stream);
int ni_fse_decompress_batch(const unsigned char* src, long long src_stride, const int* src_len, unsigned char* dst, long long dst_stride,
                            int dst_cap, int* dst_len, int n, ni_stream_t stream);
int ni_l3ic_encode(const float* latent, int n, int h, int w, int c, const float* codebook, int n_codes, unsigned char* indices,
                   unsigned char* layer_bytes, int layer_slot, int* layer_len, unsigned char* streams, long long stream_stride,
                   int* stream_len, int* status, ni_stream_t stream);
int ni_l3ic_decode(const unsigned char* streams, long long stream_stride, const int* stream_len, int n, int h, int w, int c,
                   const float* codebook, int n_codes, int* layer_off, int* layer_len, unsigned char* indices, int index_slot, float* latent,
                   int* status, ni_stream_t stream);

/* ------------------------------------------------------------------------------------------------ manipulations
 * helpers/tf_helpers.py:68-184. All on (n,h,w,3). */
/* manipulation_sharpen (tf_helpers.py:156-184): filt9 = HOST 3x3 filter applied to H and V; S takes tap [2,2]. */
int ni_manip_sharpen_fwd(const float* x, float* y, int n, int h, int w, const float* filt9, ni_stream_t stream);
/* manipulation_gaussian (tf_helpers.py:113-125): REFLECT pad k/2, depth-wise k x k filter (HOST, k*k floats), clip.
 * mask (optional, n*h*w bytes): bit c set where channel c of the un-clipped output lies in [0,1] (clip_by_value grad). */
int ni_manip_gaussian_fwd(const float* x, float* y, unsigned char* mask, int n, int h, int w, const float* filt, int k,
                          int clip, ni_stream_t stream);
int ni_manip_gaussian_bwd(const float* dy, const unsigned char* mask, float* dx, int n, int h, int w, const float* filt, int k,
                          float scale, int accumulate, ni_stream_t stream);
/* tf.image.resize(method='bilinear') (TF2 half-pixel centres, no antialias) used by manipulation_resample
 * (tf_helpers.py:68-76) and run_downsampling 'bilinear' (workflows/manipulation_classification.py:238). */
int ni_resize_bilinear_fwd(const float* x, float* y, int n, int ih, int iw, int oh, int ow, int clip, ni_stream_t stream);
int ni_resize_bilinear_bwd(const float* dy, float* dx_zeroed, int n, int ih, int iw, int oh, int ow, float scale, ni_stream_t stream);
/* manipulation_awgn (tf_helpers.py:79-82). noise: optional injected N(0,1) tensor; NULL = on-device Philox4x32-10(seed). */
int ni_manip_awgn_fwd(const float* x, const float* noise, float* y, int n, int h, int w, float strength,
                      unsigned long long seed, ni_stream_t stream);
int ni_manip_awgn_bwd(const float* x, const float* noise, const float* dy, float* dx, int n, int h, int w, float strength,
                      unsigned long long seed, int accumulate, ni_stream_t stream);
/* manipulation_gamma (tf_helpers.py:85-88) */
int ni_manip_gamma_fwd(const float* x, float* y, int n, int h, int w, float strength, ni_stream_t stream);
int ni_manip_gamma_bwd(const float* x, const float* dy, float* dx, int n, int h, int w, float strength, int accumulate,
                       ni_stream_t stream);
/* manipulation_median (tf_helpers.py:91-110), k odd */
int ni_manip_median_fwd(const float* x, float* y, int n, int h, int w, int k, ni_stream_t stream);
int ni_manip_median_bwd(const float* x, const float* dy, float* dx_accum, int n, int h, int w, int k, ni_stream_t stream);
/* run_manipulations + run_downsampling 'pool:2' fused (workflows/manipulation_classification.py:199-208, :231-245): the pooled
 * class-major stack c = (n_classes * b, h/2, w/2, 3) is written straight from y = (b, h, w, 3) for the slots listed in
 * slots[4] = {native, sharpen, resample at factor 50, gaussian 3x3 / 5x5} (class index, -1 = absent / filled by the caller through the
 * stand-alone kernels + ni_avgpool_fwd); the full-resolution stack is never materialised. The backward accumulates
 * d(loss)/dy from the pooled gradient slots (the sharpen slot carries none: helpers/tf_helpers.py:177,182 are not differentiable in TF 2.1). */
int ni_manip_stack_pool2_fwd(const float* y, float* c, unsigned char* gauss_clip_mask_or_null, int b, int h, int w, int n_classes,
                             const int* slots, const float* sharp9, const float* gauss, int gk, ni_stream_t stream);
int ni_manip_stack_pool2_bwd(const unsigned char* gauss_clip_mask, const float* dc, float* dy, int b, int h, int w, int n_classes,
                             const int* slots, const float* gauss, int gk, int accumulate, ni_stream_t stream);
/* tf.nn.avg_pool k x k stride k SAME (workflows/manipulation_classification.py:235) */
int ni_avgpool_fwd(const float* x, float* y, int n, int h, int w, int k, ni_stream_t stream);
int ni_avgpool_bwd(const float* dy, float* dx, int n, int h, int w, int k, ni_stream_t stream);
int ni_axpy(float* y, const float* x, float a, long long n, ni_stream_t stream);   /* y += a*x */

/* ------------------------------------------------------------------------------------------------ convolutions
 * Keras Conv2D / Conv2DTranspose / Dense of the reference models (models/pipelines.py:190-218,
 * models/forensics.py:65-90, models/compression.py:219-266). NHWC activations, HWIO weights (kh,kw,cin,cout). */
enum { NI_ACT_NONE = 0, NI_ACT_LEAKY_RELU = 1, NI_ACT_RELU = 2, NI_ACT_TANH = 3, NI_ACT_SIGMOID = 4, NI_ACT_CLIP01 = 5 };
enum { NI_MODE_PLAIN = 0, NI_MODE_BLOCK2 = 1 };
enum { NI_PAD_ZERO = 0, NI_PAD_SYMMETRIC = 1, NI_PAD_REFLECT = 2 };

typedef struct ni_conv_desc {
    int n, h, w;             /* logical input: batch, height, width */
    int cin, cout;
    int kh, kw;
    int stride;              /* 1 or 2 */
    int pad_t, pad_l;        /* zero padding before (TF SAME rule computed by the caller) */
    int oh, ow;              /* logical output size */
    int in_pitch, in_coff;   /* physical channel pitch and first channel of the input buffer (concat-free skips) */
    int out_pitch, out_coff;
    int in_mode;             /* NI_MODE_BLOCK2: input read through tf.nn.space_to_depth(2) of a (n,2h,2w,cin/4) buffer */
    int out_mode;            /* NI_MODE_BLOCK2: output written through tf.nn.depth_to_space(2) into (n,2oh,2ow,cout/4) */
    int act;                 /* fused after bias (fprop) */
    float act_alpha;
    int accumulate;          /* out += result */
    int bias_mod;            /* > 0: bias index = co % bias_mod (Conv2DTranspose as 1x1 conv + depth_to_space) */
    int pad_mode;            /* NI_PAD_* index mapping of out-of-range taps (fprop, wgrad) */
} ni_conv_desc;

/* Dispatchers: tcgen05 implicit GEMM where the layer is a dense contraction, FP32 SIMT otherwise. */
int ni_conv2d_fprop(const ni_conv_desc* d, const float* x, const float* w, const float* bias, float* y, ni_stream_t stream);
int ni_conv2d_dgrad(const ni_conv_desc* d, const float* dy, const float* w /* HWIO, same tensor as fprop */, float* dx, ni_stream_t stream);
/* dgrad fused with the activation backward + bias gradient of the layer BELOW (the tape's LeakyReluGrad / BiasAddGrad of the producing
 * Conv2D, models/pipelines.py:190-214): dx <- dgrad(dy) * act'(y_prev) (+= when d->accumulate), dbias_prev[c] += sum over pixels. y_prev is
 * that layer's forward output, addressed like dx with its own channel pitch / offset. tcgen05 path only: query _supported first. */
int ni_conv2d_dgrad_act_supported(const ni_conv_desc* d, int y_pitch, int y_coff);
int ni_conv2d_dgrad_act_tc(const ni_conv_desc* d, const float* dy, const float* w, float* dx, const float* y_prev, int y_pitch, int y_coff,
                           int act_prev, float alpha_prev, float* dbias_prev, int bias_mod_prev, ni_stream_t stream);
int ni_conv2d_wgrad(const ni_conv_desc* d, const float* x, const float* dy, float* dw, ni_stream_t stream);
/* The tcgen05 (3xTF32, FP32-accurate) implementations; ni_conv2d_tc_supported(d, op) op: 0 fprop, 1 dgrad, 2 wgrad. */
int ni_conv2d_tc_supported(const ni_conv_desc* d, int op);
int ni_conv2d_fprop_tc(const ni_conv_desc* d, const float* x, const float* w, const float* bias, float* y, ni_stream_t stream);
int ni_conv2d_dgrad_tc(const ni_conv_desc* d, const float* dy, const float* w, float* dx, ni_stream_t stream);
int ni_conv2d_wgrad_tc(const ni_conv_desc* d, const float* x, const float* dy, float* dw, ni_stream_t stream);
/* Direct FP32 stencils for the 3/4-input-channel and 3/12-output-channel layers (FAN front end, U-Net first / last conv):
 * register-blocked kernels for the hot shapes (FAN 5x5 3->32 and its input gradient, U-Net 4->32 and 32->12), generic
 * thread-per-pixel stencils ("small") for the other shapes of INet / DNet / ClassicISP / TwitterDCN. */
int ni_conv2d_direct_supported(const ni_conv_desc* d, int op);
int ni_conv2d_fprop_direct(const ni_conv_desc* d, const float* x, const float* w, const float* bias, float* y, ni_stream_t stream);
int ni_conv2d_dgrad_direct(const ni_conv_desc* d, const float* dy, const float* w, float* dx, ni_stream_t stream);
int ni_conv2d_wgrad_direct(const ni_conv_desc* d, const float* x, const float* dy, float* dw, ni_stream_t stream);
/* Conv2D (3 / 4 input channels) + bias + activation + MaxPool2D(2x2) in one kernel (first block of the FAN, models/forensics.py:66-69):
 * pooled (n, oh/2, ow/2, cout) + one code byte per pooled element (argmax position | (value > 0) << 2); the full-resolution activation
 * is never written. ni_maxpool2_code_bwd_bias is the matching backward of (activation + pooling): dx (n, oh, ow, c) from dy, dbias overwritten. */
int ni_conv2d_pool2_supported(const ni_conv_desc* d);
int ni_conv2d_pool2_fwd(const ni_conv_desc* d, const float* x, const float* w, const float* bias, float* pooled, unsigned char* code,
                        ni_stream_t stream);
int ni_maxpool2_code_bwd_bias(const unsigned char* code, const float* dy, float* dx, float* dbias, int n, int oh, int ow, int c, int act,
                              float alpha, ni_stream_t stream);
int ni_conv2d_small_supported(const ni_conv_desc* d, int op);
int ni_conv2d_fprop_small(const ni_conv_desc* d, const float* x, const float* w, const float* bias, float* y, ni_stream_t stream);
int ni_conv2d_dgrad_small(const ni_conv_desc* d, const float* dy, const float* w, float* dx, ni_stream_t stream);
int ni_conv2d_wgrad_small(const ni_conv_desc* d, const float* x, const float* dy, float* dw, ni_stream_t stream);
/* test hook: 1 = run every convolution on the FP32 SIMT kernels (tests compare the tcgen05 path against them on the device), 0 = dispatch normally */
void ni_conv2d_set_force_simt(int on);
/* The SIMT implementations, callable directly (tests compare the two paths on the device). */
int ni_conv2d_fprop_simt(const ni_conv_desc* d, const float* x, const float* w, const float* bias, float* y, ni_stream_t stream);
int ni_conv2d_dgrad_simt(const ni_conv_desc* d, const float* dy, const float* w_t, float* dx, ni_stream_t stream);
int ni_conv2d_wgrad_simt(const ni_conv_desc* d, const float* x, const float* dy, float* dw, ni_stream_t stream);
int ni_weight_transpose_io(const float* w, float* w_t, int taps, int cin, int cout, ni_stream_t stream);

/* ------------------------------------------------------------------------------------------------ other layers */
/* MaxPool2D 2x2 (SAME in UNet models/pipelines.py:197, VALID in FAN models/forensics.py:70) */
int ni_maxpool2_fwd(const float* x, float* y, int n, int h, int w, int c, int same, int x_pitch, int x_coff, int y_pitch,
                    int y_coff, ni_stream_t stream);
int ni_maxpool2_bwd(const float* x, const float* dy, const float* add_or_null, float* dx, int n, int h, int w, int c, int same,
                    int x_pitch, int x_coff, int dy_pitch, int dy_coff, int add_pitch, int add_coff, int dx_pitch, int dx_coff,
                    ni_stream_t stream);
/* ni_maxpool2_bwd + ni_act_bwd_bias of the pooled layer (x = that layer's activated output) in one pass; dbias may be NULL */
int ni_maxpool2_act_bwd_bias(const float* x, const float* dy, const float* add_or_null, float* dx, float* dbias_or_null, int n, int h, int w,
                             int c, int same, int x_pitch, int x_coff, int dy_pitch, int dy_coff, int add_pitch, int add_coff, int dx_pitch,
                             int dx_coff, int act, float alpha, ni_stream_t stream);
/* dy <- dy * act'(y) in place and dbias = column sums (Keras layer activation + bias gradients) */
int ni_act_bwd_bias(const float* y, float* dy, float* dbias, int n, int h, int w, int c, int y_pitch, int y_coff, int y_mode,
                    int dy_pitch, int dy_coff, int dy_mode, int act, float alpha, int bias_mod, ni_stream_t stream);
/* GlobalAveragePooling2D (models/forensics.py:79) */
int ni_gap_fwd(const float* x, float* y, int n, int hw, int c, ni_stream_t stream);
int ni_gap_bwd(const float* dy, float* dx, int n, int hw, int c, ni_stream_t stream);
/* softmax + SparseCategoricalCrossentropy on probabilities, Keras eager semantics (models/forensics.py:90,94) */
int ni_softmax_ce(const float* logits, const int* labels, float* probs, float* loss_sum, float* dlogits, int m, int c,
                  float gscale, ni_stream_t stream);
/* ConstrainedConv2D.call (models/layers.py:45-57) on the normalised filter nf (5, 5, 3, 3): SYMMETRIC pad 2 + VALID conv, no bias; its
   input gradient with the transpose of the mirrored pad folded in; its filter gradient (dnf is zeroed first). x, y, dy, dx: (n, h, w, 3) */
int ni_cconv5_fwd(const float* x, const float* nf, float* y, int n, int h, int w, ni_stream_t stream);
int ni_cconv5_bwd_data(const float* dy, const float* nf, float* dx, int n, int h, int w, int accumulate, ni_stream_t stream);
int ni_cconv5_bwd_filter(const float* x, const float* dy, float* dnf, int n, int h, int w, ni_stream_t stream);
/* A SAME stride-2 5x5 convolution on an even-sized input (the DCN encoder's down-sampling layers, models/compression.py:221-222,237) =
 * a SAME stride-1 3x3 convolution over space_to_depth(2) of the input with the taps scattered into a zero-padded 6x6 grid: that form runs
 * on the tcgen05 path. ni_space_to_depth2 moves activations (inverse = 1: plain (+)= deep, the input-gradient way back), ni_s2conv_weights
 * moves weights (inverse = 1: the 5x5 weight gradient gathered from the 3x3 x 4cin one). tf.nn.space_to_depth channel order. */
int ni_space_to_depth2(float* plain, float* deep, int n, int h2, int w2, int c, int inverse, int accumulate, ni_stream_t stream);
int ni_s2conv_weights(float* w5, float* w3, int cin, int cout, int inverse, ni_stream_t stream);
/* tf.keras.layers.Dropout(rate) in training mode (models/forensics.py:88; FAN.process(x, training=True)): inverted dropout, own generator */
int ni_dropout(const float* x, float* y, long long n, float rate, unsigned long long seed, ni_stream_t stream);
/* validate_fan's decisions and confusion matrix (training/validation.py:163-203; workflows/manipulation_classification.py:178-180):
   pred[i] = argmax probs[i, :], conf[labels[i] * c + pred[i]] += 1 (int32, caller zeroes); pred or conf may be null */
int ni_confusion_accumulate(const float* probs, const int* labels, int* conf, int* pred, int m, int c, ni_stream_t stream);
/* tf_helpers.mse / mae (helpers/tf_helpers.py:31-36): kind 0 = L2, 1 = L1 */
int ni_image_loss(const float* a, const float* b, float* acc, long long n, int kind, ni_stream_t stream);
int ni_image_loss_grad(const float* a, const float* b, float* da, long long n, int kind, float scale, int accumulate,
                       ni_stream_t stream);
/* ConstrainedConv2D filter normalisation (models/layers.py:45-53) and its backward */
int ni_constrained_filter_fwd(const float* k, float* nf, int ksize, int channels, float strength, ni_stream_t stream);
int ni_constrained_filter_bwd(const float* k, const float* dnf, float* dk, int ksize, int channels, float strength, ni_stream_t stream);
/* gradient of tf.pad(SYMMETRIC|REFLECT): fold the padded-domain gradient back (models/layers.py:56) */
int ni_pad_fold(const float* dpad, float* dx, int n, int h, int w, int c, int pad, int mode, int accumulate, ni_stream_t stream);
/* tf.keras.optimizers.Adam.apply_gradients on a flat buffer (workflows/manipulation_classification.py:156,279-283);
 * also replaces the host-side NaN scan (:281): *nonfinite_flag |= 1 if any gradient is NaN/Inf. */
int ni_adam_keras(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps,
                  long long step, float gscale, int* nonfinite_flag, ni_stream_t stream);
/* Same update with lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t) read from DEVICE memory, so that a captured CUDA graph of the
 * training step can be replayed with a new step count / learning rate (optimizer.lr.assign, :279). */
int ni_adam_keras_dev(float* p, const float* g, float* m, float* v, long long n, const float* lr_t_dev, float beta1, float beta2,
                      float eps, float gscale, int* nonfinite_flag, ni_stream_t stream);
/* DiscreteLatent + Quantization('soft-codebook') + tf_helpers.entropy histogram (models/layers.py:139-170,195-203,
 * helpers/tf_helpers.py:290-333), float64 inside like the reference. hist_acc: ncodes doubles zeroed by the caller (sum of the
 * normalised weights per bin, evaluated at the QUANTISED values like the reference); q: the forward's output;
 * gh: ncodes doubles = d loss / d hist_k / n; dscale_acc: one double zeroed by the caller. */
int ni_latent_softcodebook_fwd(const float* x, const float* scale, const float* codebook, float* out, double* hist_acc, long long n,
                               int ncodes, double nu, double gamma, ni_stream_t stream);
int ni_latent_softcodebook_bwd(const float* x, const float* scale, const float* codebook, const float* q, const float* g_out,
                               const double* gh, float* dx, double* dscale_acc, long long n, int ncodes, double nu, double gamma, ni_stream_t stream);
/* the same with the latent quantiser selectable (DCN's `rounding`, models/compression.py:66 -> models/layers.py:118-170):
 * rounding 0 'soft-codebook', 1 'sin', 2 'soft' (round forward, sine gradient), 3 'identity'; the entropy term is unchanged */
int ni_latent_quantise_fwd(const float* x, const float* scale, const float* codebook, float* out, double* hist_acc, long long n,
                           int ncodes, double nu, double gamma, int rounding, ni_stream_t stream);
int ni_latent_quantise_bwd(const float* x, const float* scale, const float* codebook, const float* q, const float* g_out,
                           const double* gh, float* dx, double* dscale_acc, long long n, int ncodes, double nu, double gamma, int rounding,
                           ni_stream_t stream);
/* Quantization layer, scalar modes (models/layers.py:118-136): 0 'round', 1 'sin', 2 'soft', 3 'harmonic', 4 'identity' (forward values) */
int ni_quantize_scalar(const float* x, float* y, long long n, int mode, int taylor_terms, ni_stream_t stream);
/* entropy estimate from the accumulated soft histogram (helpers/tf_helpers.py:326-331) and its gradient w.r.t. the histogram */
int ni_entropy_from_hist(const double* hist_acc, long long n, int ncodes, double upstream, float* h_out, double* gh_out, ni_stream_t stream);
/* tf.nn.leaky_relu on a residual-branch input (models/compression.py:224) */
int ni_leaky_relu_fwd(const float* x, float* y, long long n, float alpha, ni_stream_t stream);
int ni_leaky_relu_bwd(const float* x, const float* dy, float* dx, long long n, float alpha, int accumulate, ni_stream_t stream);
/* ClassicISP tails: straight-through clip + gamma (models/pipelines.py:441-445) and the residual demosaicing combination
 * y = x_bilinear - alpha * f with a trainable device scalar alpha (models/layers.py:249-256). dalpha is accumulated into. */
int ni_gamma_clip_fwd(const float* x, float* y, long long n, float lo, float hi, float exponent, ni_stream_t stream);
int ni_gamma_clip_bwd(const float* x, const float* dy, float* dx, long long n, float lo, float hi, float exponent, ni_stream_t stream);
int ni_residual_alpha_fwd(const float* xb, const float* f, const float* alpha, float* y, long long n, int clip, ni_stream_t stream);
int ni_residual_alpha_bwd(const float* dy, const float* f, const float* alpha, float* df, float* dalpha, long long n, ni_stream_t stream);
int ni_fill(float* p, float value, long long n, ni_stream_t stream);
int ni_affine(const float* x, float* y, float a, float b, int clip, long long n, ni_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* NI_B200_H */
