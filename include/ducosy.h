/* libducosy_sm100.so -- C ABI of the B200-native DuCoSy-GAN hot path.
 *
 * The reference (qqaazz0222/DuCoSy-GAN) is pure Python/PyTorch and has no FFI of its own; each entry point
 * below names the reference code (file:line under /root/reference) whose arithmetic it replaces.  The Python
 * host (ducosy_gan_b200/modules/model.py, ducosy_gan_b200/synthesis.py) binds these with ctypes and presents
 * the reference's modules/model.py API on top.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host; the caller owns all memory, the library
 *     never allocates or frees device memory and keeps no pointer after returning;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), makes no host sync and is
 *     CUDA-graph capturable;
 *   - return value: 0 on success, a negative DUCOSY_ERR_* code otherwise; ducosy_last_error() gives the
 *     thread-local message.  Nothing aborts or throws across the ABI;
 *   - activations are NHWC 16-bit (DUCOSY_F16 or DUCOSY_BF16), statistics / final output fp32;
 *   - only sm_100 (B200) devices are accepted: there is no fallback path.
 */
#ifndef DUCOSY_H_
#define DUCOSY_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DUCOSY_VERSION 100

/* Operand type of the tensor-core convolutions.  DUCOSY_F16 / DUCOSY_BF16: 16-bit activations and weights, fp32 accumulate.
 * DUCOSY_F16X2 ("split-operand" mode, the <= 1 HU precision arm of the generator forward): every activation and packed
 * weight is a pair of fp16 values (hi = rn(v), lo = rn(v - hi)); NHWC buffers hold 2*C 16-bit channels per pixel (hi plane,
 * then lo plane), convolutions accumulate A_hi*W_hi + A_lo*W_hi + A_hi*W_lo in fp32 (3 tensor-core products per tap, ~21
 * significant bits).  Accepted where a function's comment says so (ducosy_generator_* and the kernels it is made of);
 * the training entry points take the 16-bit types only. */
enum { DUCOSY_F16 = 0, DUCOSY_BF16 = 1, DUCOSY_F16X2 = 2 };
enum {
  DUCOSY_OK = 0,
  DUCOSY_ERR_SHAPE = -1,     /* unsupported shape / channel count */
  DUCOSY_ERR_ALIGN = -2,     /* pointer alignment */
  DUCOSY_ERR_WORKSPACE = -3, /* workspace too small */
  DUCOSY_ERR_CUDA = -4,      /* CUDA runtime / driver error (launch failure, no device) */
  DUCOSY_ERR_ARCH = -5,      /* device is not sm_100 */
  DUCOSY_ERR_ARG = -6        /* null pointer / bad enum */
};
enum { DUCOSY_PAD_ZERO = 0, DUCOSY_PAD_REFLECT = 1 };
enum { DUCOSY_ACT_NONE = 0, DUCOSY_ACT_RELU = 1, DUCOSY_ACT_LRELU02 = 2 };

typedef void* ducosy_stream_t; /* cudaStream_t */

int ducosy_version(void);
const char* ducosy_last_error(void);
/* 0 if the current device is sm_100, DUCOSY_ERR_ARCH / DUCOSY_ERR_CUDA otherwise. */
int ducosy_check_device(void);

/* ---------------------------------------------------------------- HU windowing / composite (bandwidth kernels) */

/* preprocess_dicom, modules/preprocess.py:72-84: hu = float(px)*slope + intercept; for each of the two windows
 * clip(hu, lo, hi) then 2*(x-lo)/(hi-lo)-1, every step rounded to fp32 exactly like numpy.  n stored values ->
 * out_soft[n], out_lung[n] (fp32).  Either output may be NULL. */
int ducosy_hu_window(const int16_t* px, float* out_soft, float* out_lung, long long n, float slope, float intercept,
                     float soft_lo, float soft_hi, float lung_lo, float lung_hi, ducosy_stream_t stream);

/* Training-side windowing with soft squeezing, apply_hu_transform + apply_soft_squeezing (modules/preprocess.py:6-55,
 * modules/dataset.py:118-120): n stored values -> out[n] fp32 in [-1, 1]; sigma = 50 in the reference.  Same float32 steps
 * as numpy; only exp() may differ by an ulp. */
int ducosy_hu_window_soft(const int16_t* px, float* out, long long n, float slope, float intercept, float hu_lo, float hu_hi,
                          float sigma, ducosy_stream_t stream);
/* apply_windowing (modules/preprocess.py:58-65): tanh-range tensor -> display intensity in [0,1] for a (center, width) window. */
int ducosy_apply_windowing(const float* y, float* out, long long n, float hu_lo, float hu_hi, float window_center,
                           float window_width, ducosy_stream_t stream);

/* HU threshold candidates of the anatomical mask generator (modules/mask_generator.py:14-20,179-183):
 * body = hu > -1000, lung = -1000 <= hu <= -300 & body, bone = hu >= 200 & body, uint8 {0,1}. NULL outputs are skipped. */
int ducosy_hu_thresholds(const int16_t* px, uint8_t* body, uint8_t* lung, uint8_t* bone, long long n, float slope,
                         float intercept, ducosy_stream_t stream);

/* postprocess_tensor (modules/preprocess.py:96-111) for both generators followed by the complementary composite
 * (generate.py:140-145,218-237) in ONE pass: y_soft / y_lung are the tanh outputs (fp32), raw_px the NCCT stored
 * values.  merged[i] = lung_px if lung range, else soft_px if soft range, else raw (lung wins at HU == lung_hi).
 * Optional outputs (may be NULL): soft_px / lung_px (the de-windowed stored values, truncation toward zero) and
 * masks (bit0 = soft range, bit1 = lung range).  Bit-exact with the numpy reference for the same y. */
int ducosy_dewindow_composite(const int16_t* raw_px, const float* y_soft, const float* y_lung, int16_t* merged,
                              int16_t* soft_px, int16_t* lung_px, uint8_t* masks, long long n, float slope,
                              float intercept, float soft_lo, float soft_hi, float lung_lo, float lung_hi,
                              ducosy_stream_t stream);

/* ---------------------------------------------------------------- layer kernels (exported for tests and reuse) */

/* Pack an OIHW fp32 conv weight into the K-major 16-bit GEMM operand [Cout][kh*kw*Cin] (k = (r*kw+s)*Cin + c). */
int ducosy_pack_conv_weight(const float* w_oihw, void* packed, int Cout, int Cin, int kh, int kw, int dtype,
                            ducosy_stream_t stream);
/* Pack the 3x3 weight of "Upsample(x2, nearest) + Conv2d(3x3, pad 1)" (modules/model.py:108) into four
 * phase-specific 2x2 kernels on the source grid: [4*Cout][4*Cin], phase = py*2+px, tap = a*2+b. */
int ducosy_pack_upconv_weight(const float* w_oihw, void* packed, int Cout, int Cin, int dtype, ducosy_stream_t stream);
/* Merged-phase packing of the same layer for Cout = 64: [4*Cout][9*Cin], row = phase*Cout + o, k = (dy*3+dx)*Cin + c
 * over the nine source offsets; zero where a phase does not use an offset (one N = 256 GEMM instead of four N = 64). */
int ducosy_pack_upconv_merged_weight(const float* w_oihw, void* packed, int Cout, int Cin, int dtype,
                                     ducosy_stream_t stream);
/* Pack the 7x7 stem weight [64][Cin][7][7] into [64][Kpad], Kpad = roundup(49*Cin, 64), k = c*49 + r*7 + s. */
int ducosy_pack_stem_weight(const float* w_oihw, void* packed, int Cin, int dtype, ducosy_stream_t stream);

/* Convolution as implicit GEMM on tcgen05 tensor cores.
 *   in      : NHWC 16-bit, already padded: [B][Hp][Wp][Cin], Cin % 64 == 0
 *   w       : packed by ducosy_pack_conv_weight
 *   out     : raw (pre-normalisation) output NHWC [B][Ho][Wo][Cout]
 *   partials: fp32 [B][Ho*Wo/128][3][Cout] per-tile per-channel (sum, sum of squares, max), or NULL
 *   kh,kw,stride: 3,3,1 (input padded by 1) | 3,3,2 (zero-padded by 1) | 4,4,2 (zero-padded by 1) | 1,1,1
 *   bias/act: NULL/0 for convs followed by InstanceNorm; bias + LeakyReLU(0.2) when act == DUCOSY_ACT_LRELU02. */
int ducosy_conv2d_nhwc(const void* in, const void* w, void* out, float* partials, const float* bias, int act, int B,
                       int Hp, int Wp, int Cin, int Cout, int kh, int kw, int stride, int dtype, ducosy_stream_t stream);
/* Upsample(x2)+Conv3x3 (modules/model.py:108-109) from the zero-padded source [B][Hs+2][Ws+2][Cin];
 * out [B][2Hs][2Ws][Cout]; partials [B][4*Hs*Ws/128][3][Cout]. */
int ducosy_upconv2x_nhwc(const void* in_pad, const void* w_packed4, void* out, float* partials, int B, int Hs, int Ws,
                         int Cin, int Cout, int dtype, ducosy_stream_t stream);

/* Same layer through the merged-phase packing (Cout = 64): partials [B][Hs*Ws/128][3][Cout]. */
int ducosy_upconv2x_merged_nhwc(const void* in_pad, const void* w_merged, void* out, float* partials, int B, int Hs,
                                int Ws, int Cin, int Cout, int dtype, ducosy_stream_t stream);

/* The three convolution entry points above with the InstanceNorm finalize (ducosy_in_finalize without the CBAM weights) fused
 * into the same launch: the CTA that completes the last tile of a sample reduces that sample's per-tile partials in fixed
 * order and writes scale = 1/sqrt(var + 1e-5), shift = -mean * scale [B][Cout] (and, when chmax != NULL, the per-channel max
 * of the normalised map).  `partials` is still required (it is the reduction's input).  `tickets`: int32 [B] in device memory,
 * ZERO before the first use and left zero by every launch; launches that may overlap (different streams) need different
 * ticket arrays.  Removes one dependent launch behind every convolution (modules/model.py:94-111: Conv -> InstanceNorm). */
int ducosy_conv2d_nhwc_in(const void* in, const void* w, void* out, float* partials, float* scale, float* shift, float* chmax,
                          int* tickets, int B, int Hp, int Wp, int Cin, int Cout, int kh, int kw, int stride, int dtype,
                          ducosy_stream_t stream);
int ducosy_upconv2x_nhwc_in(const void* in_pad, const void* w_packed4, void* out, float* partials, float* scale, float* shift,
                            int* tickets, int B, int Hs, int Ws, int Cin, int Cout, int dtype, ducosy_stream_t stream);
int ducosy_upconv2x_merged_nhwc_in(const void* in_pad, const void* w_merged, void* out, float* partials, float* scale, float* shift,
                                   int* tickets, int B, int Hs, int Ws, int Cin, int Cout, int dtype, ducosy_stream_t stream);

/* im2col for the 7x7 reflect-padded stem (modules/model.py:94): x fp32 NCHW [B][Cin][H][W] -> A [B*H*W][Kpad]. */
int ducosy_stem_im2col(const float* x_nchw, void* a_mat, int B, int Cin, int H, int W, int dtype, ducosy_stream_t stream);
/* Same, fused with the HU windowing of modules/preprocess.py:72-84 (Cin = 1): stored px int16 [B][H][W] in. */
int ducosy_stem_im2col_hu(const int16_t* px, void* a_mat, int B, int H, int W, float slope, float intercept, float lo,
                          float hi, int dtype, ducosy_stream_t stream);

/* Fused stem for input_channels == 1 (modules/preprocess.py:72-84 + modules/model.py:94).
 * ducosy_stem_prepare: network input -- either x_nchw (fp32 [B][1][H][W]) or px (stored int16 [B][H][W] through the
 *   slope/intercept/lo/hi window) -- -> 16-bit xw [B][H][W].
 * ducosy_stem_fused(xw, w_packed from ducosy_pack_stem_weight, ...):
 *   apply == 0: statistics pass -> partials [B][H][3][64] (then ducosy_in_finalize(partials, H, H*W, ...));
 *   apply == 1: conv again, (y*scale+shift), ReLU -> out_pad [B][H+2][W+2][64] 16-bit with a zero border. */
int ducosy_stem_prepare(const float* x_nchw, const int16_t* px, float slope, float intercept, float lo, float hi, void* xw,
                        int B, int H, int W, int dtype, ducosy_stream_t stream);
int ducosy_stem_fused(const void* xw, const void* w_packed, float* partials, const float* scale, const float* shift,
                      void* out_pad, int B, int H, int W, int apply, int dtype, ducosy_stream_t stream);

/* InstanceNorm statistics (modules/model.py InstanceNorm2d: biased var, eps 1e-5) from the conv partials:
 * scale[b][c] = rstd, shift[b][c] = -mean*rstd.  When fc0/fc2 are given (CBAM channel attention,
 * modules/model.py:13-24: fc0 [C/16][C], fc2 [C][C/16], fp32) the channel attention
 * s = sigmoid(fc(avgpool) + fc(maxpool)) of the NORMALISED map is folded in: scale *= s, shift *= s;
 * chmax_scratch [B][C] fp32 then receives the per-channel max of the normalised map (required with fc0/fc2). */
int ducosy_in_finalize(const float* partials, int tiles_per_sample, int npix_per_sample, float* scale, float* shift,
                       const float* fc0, const float* fc2, float* chmax_scratch, int B, int C, ducosy_stream_t stream);

/* out_pad[b][y+p][x+p][c] = act(y*scale + shift), borders filled by reflection or zeros. */
int ducosy_in_apply_pad(const void* y, const float* scale, const float* shift, void* out_pad, int B, int H, int W, int C,
                        int pad, int pad_mode, int act, int dtype, ducosy_stream_t stream);

/* CBAM spatial attention (modules/model.py:34-39), C == 256:
 *   pool : pooled[b][y][x] = (mean_c v, max_c v), v = y*scale + shift
 *   conv : sa = sigmoid(conv7x7_{2->1, zero pad 3}(pooled)), w_sa = spatial_attention.conv.weight [1][2][7][7] fp32
 *   apply: out_pad = res_pad(interior) + v * sa   (sa == NULL: plain ResidualBlock, modules/model.py:65) */
int ducosy_cbam_pool(const void* y, const float* scale, const float* shift, float* pooled, int B, int H, int W, int C,
                     int dtype, ducosy_stream_t stream);
int ducosy_cbam_spatial_conv(const float* pooled, const float* w_sa, float* sa, int B, int H, int W, ducosy_stream_t stream);
int ducosy_residual_apply_pad(const void* y, const float* scale, const float* shift, const float* sa,
                              const void* res_pad, int res_pad_width, void* out_pad, int B, int H, int W, int C, int pad,
                              int pad_mode, int dtype, ducosy_stream_t stream);
/* The same pass with the spatial attention of modules/model.py:34-39 evaluated inside: `pooled` [B][H][W][2] (ducosy_cbam_pool)
 * and the 7x7 conv weight `w_sa` [1][2][7][7] replace the precomputed attention map and the ducosy_cbam_spatial_conv launch. */
int ducosy_residual_cbam_apply_pad(const void* y, const float* scale, const float* shift, const float* pooled, const float* w_sa,
                                   const void* res_pad, int res_pad_width, void* out_pad, int B, int H, int W, int C, int pad,
                                   int pad_mode, int dtype, ducosy_stream_t stream);

/* Output conv (modules/model.py:112): reflect-padded input [B][H+6][W+6][64] 16-bit, weight [1][64][7][7] fp32,
 * bias[1] fp32 -> out fp32 [B][H][W] = tanh(conv + bias).  w_packed from ducosy_pack_out_weight. */
int ducosy_pack_out_weight(const float* w_oihw, void* packed, int dtype, ducosy_stream_t stream);
int ducosy_out_conv7x7_tanh(const void* in_pad, const void* w_packed, const float* bias, float* out, int B, int H, int W,
                            int dtype, ducosy_stream_t stream);
/* Same from the RAW previous conv output y_raw [B][H][W][64]: InstanceNorm apply (scale/shift [B][64]) + ReLU + the
 * reflection padding are folded into the loader (modules/model.py:110-112 in one kernel). */
int ducosy_out_conv7x7_tanh_fused(const void* y_raw, const float* scale, const float* shift, const void* w_packed,
                                  const float* bias, float* out, int B, int H, int W, int dtype, ducosy_stream_t stream);

/* Post-composite volume smoothing (generate.py:254-263 + modules/postprocess.py:47-60,99-109,114-160):
 * z Gaussian (float32) -> z Gaussian (float32) -> xy unsharp mask in float64 -> clip to the range of the first result ->
 * voxels >= hu_threshold keep the first result -> int16 (truncation).  Bit-exact with scipy's correlate1d arithmetic
 * (symmetric-kernel order, no FMA, mode 'reflect').  merged/out: device int16 [S][H][W]; scratch: device,
 * ducosy_postprocess_scratch_bytes; wz1/wz2/wxy: HOST arrays of 2r+1 normalised float64 weights (radius 1..4 for z,
 * 1..8 for xy), as scipy.ndimage computes them.
 * phases: 1 = the two z filters + (min, max) of the first result over slices [mm_z0, mm_z1) written to the two floats at
 * ducosy_postprocess_minmax_offset_bytes inside scratch; 2 = the in-plane part, reading that pair; 3 = both.  A
 * z-sharded volume runs phase 1 on its halo-extended slab, min/max-all-reduces the pair across ranks, then runs phase 2. */
size_t ducosy_postprocess_scratch_bytes(int S, int H, int W);
size_t ducosy_postprocess_minmax_offset_bytes(int S, int H, int W);
int ducosy_postprocess_volume(const int16_t* merged, int16_t* out, float* scratch, int S, int H, int W, const double* wz1, int rz1,
                              const double* wz2, int rz2, const double* wxy, int rxy, double sharpen_amount, float hu_threshold,
                              int phases, int mm_z0, int mm_z1, ducosy_stream_t stream);

/* Generator backward pieces (what autograd computes through modules/model.py:90-92 and :112-113 for
 * modules/trainer.py:497).  16-bit gradient maps carry the power-of-two scale gs[0]; fp32 results leave with the true scale.
 *
 * Output conv + tanh: dout, out fp32 [B][H][W]; in_pad = the reflect-padded activation the forward conv read
 * [B][H+6][W+6][64] 16-bit; w fp32 [1][64][7][7] -> da 16-bit [B][H][W][64] (scaled by gs[0]), dw fp32 [64*49], db fp32 [1]. */
size_t ducosy_out_conv_backward_scratch_bytes(int B, int H, int W);
int ducosy_out_conv_backward(const float* dout, const float* out, const void* in_pad, const float* w, void* da, float* dw,
                             float* db, float* scratch, const float* gs, int B, int H, int W, int dtype, ducosy_stream_t stream);
/* Stem: dcol [B][H][W][64] 16-bit = gradient of the first 64 im2col columns (channel 0, k = r*7+s) -> image gradient
 * dx fp32 [B][H][W] (true scale: multiplied by gs[1]) through the adjoint of ReflectionPad2d(3) + im2col. */
int ducosy_stem_col2im(const void* dcol, float* dx, const float* gs, int B, int H, int W, int dtype, ducosy_stream_t stream);
/* packed [64][Kpad] weight gradient of the stem GEMM -> OIHW [64][Cin][7][7], multiplied by gs[1]. */
int ducosy_unpack_stem_wgrad(const float* packed, float* g_oihw, int Cin, int Kpad, const float* gs, ducosy_stream_t stream);
/* CBAM in training mode (modules/model.py:12-53): the channel-attention MLP keeps ca [B][C], the hidden layer [B][C/16] and
 * writes the attention-folded affine (scale_v, shift_v) next to the un-folded InstanceNorm one (scale_n, shift_n);
 * chmax [B][C] is the normalised per-channel max ducosy_in_finalize emits. */
int ducosy_cbam_channel_train(const float* chmax, const float* fc0, const float* fc2, const float* scale_n, const float* shift_n,
                              float* scale_v, float* shift_v, float* ca, float* hidden, int B, int C, ducosy_stream_t stream);
/* CBAM backward for the 256-channel residual blocks: dout = gradient of (x + CBAM(n)) w.r.t. the CBAM branch
 * [B][H][W][C] 16-bit (scaled by gs[0]); yb = raw conv output; pooled [B][H][W][2], sa [B][H][W] from the forward.
 * Returns dn (gradient w.r.t. n = InstanceNorm(yb), 16-bit, scaled) and the fp32 true-scale gradients of fc.0.weight
 * [C/16][C], fc.2.weight [C][C/16] and spatial_attention.conv.weight [1][2][7][7]. */
size_t ducosy_cbam_backward_scratch_bytes(int B, int H, int W, int C);
int ducosy_cbam_backward(const void* dout, const void* yb, const float* scale_n, const float* shift_n, const float* scale_v,
                         const float* shift_v, const float* ca, const float* hidden, const float* chmax, const float* pooled,
                         const float* sa, const float* fc0, const float* fc2, const float* wsa, void* dn, float* dfc0, float* dfc2,
                         float* dwsa, float* scratch, const float* gs, int B, int H, int W, int C, int dtype,
                         ducosy_stream_t stream);
/* The same with the three parameter gradients ACCUMULATED (+=) into dfc0 / dfc2 / dwsa (the parameters' existing .grad). */
int ducosy_cbam_backward_acc(const void* dout, const void* yb, const float* scale_n, const float* shift_n, const float* scale_v,
                             const float* shift_v, const float* ca, const float* hidden, const float* chmax, const float* pooled,
                             const float* sa, const float* fc0, const float* fc2, const float* wsa, void* dn, float* dfc0, float* dfc2,
                             float* dwsa, float* scratch, const float* gs, int B, int H, int W, int C, int dtype,
                             ducosy_stream_t stream);
/* torch.optim.Adam step (no weight decay, no amsgrad) as constructed at modules/trainer.py:360-362: fp32 param / grad /
 * exp_avg / exp_avg_sq of n elements updated in one pass; step counts from 1. */
int ducosy_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, float lr, float beta1,
                     float beta2, float eps, int step, ducosy_stream_t stream);
/* The same update with {lr, step} read from DEVICE memory (state[0] = lr, state[1] = step as a float), for CUDA-graph capture
 * of a whole optimisation step: ducosy_adam_advance increments the step once per optimizer.step(), then one
 * ducosy_adam_step_dev per tensor. */
int ducosy_adam_advance(float* state, ducosy_stream_t stream);
int ducosy_adam_step_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, const float* state,
                         float beta1, float beta2, float eps, ducosy_stream_t stream);
/* One optimizer.step() of torch.optim.Adam (modules/trainer.py:360-362) over ALL tensors of a parameter group in three launches,
 * with a non-finite guard.  `tensors_dev` / `chunks_dev` are tables in DEVICE memory built by the caller: one
 * ducosy_adam_tensor per parameter, and one ducosy_adam_chunk per DUCOSY_ADAM_CHUNK elements of each tensor (start is a
 * multiple of DUCOSY_ADAM_CHUNK).  `state` is 8 floats in device memory:
 *   [0] lr   [1] step count (advanced by this call on a clean step)   [2] scratch flag (must start at 0)
 *   [3] number of skipped steps   [4] 1 if THIS call was skipped   [5..7] reserved.
 * With check_finite != 0 a step whose gradients contain Inf/NaN changes nothing but [3] and [4] (torch.amp.GradScaler
 * semantics): the 16-bit gradient maps of the training path can overflow where the reference's fp32 autograd cannot.
 * Nothing that changes between steps is a launch argument, so the call is CUDA-graph capturable. */
#define DUCOSY_ADAM_CHUNK 16384
typedef struct {
  float* param;
  const float* grad;
  float* exp_avg;
  float* exp_avg_sq;
  long long n;
} ducosy_adam_tensor;
typedef struct {
  int tensor;      /* index into the tensor table */
  int reserved;
  long long start; /* first element of the chunk */
} ducosy_adam_chunk;
int ducosy_adam_multi_step(const ducosy_adam_tensor* tensors_dev, const ducosy_adam_chunk* chunks_dev, int num_chunks, float* state,
                           float beta1, float beta2, float eps, int check_finite, ducosy_stream_t stream);
/* a += b on 16-bit maps of n elements (n % 8 == 0): the skip connection of modules/model.py:65,87 in the backward. */
int ducosy_add_inplace(void* a, const void* b, long long n, int dtype, ducosy_stream_t stream);

/* ---------------------------------------------------------------- stand-alone building blocks (NCHW fp32, the reference's layout)
 * modules/model.py:6-24  ChannelAttention.forward: out = x * sigmoid(fc(avgpool x) + fc(maxpool x)) with BOTH pooling branches;
 * fc0 [hidden][C] and fc2 [C][hidden] are the 1x1 conv weights (no bias).  scratch: ducosy_channel_attention_scratch_bytes. */
size_t ducosy_channel_attention_scratch_bytes(int B, int C);
int ducosy_channel_attention_nchw(const float* x, const float* fc0, const float* fc2, float* out, float* scratch, int B, int C,
                                  int hidden, int H, int W, ducosy_stream_t stream);
/* modules/model.py:27-39  SpatialAttention.forward: out = x * sigmoid(conv_kxk(cat[mean_C x, max_C x])), weight [1][2][k][k],
 * zero padding k/2, no bias, k odd.  scratch: ducosy_spatial_attention_scratch_bytes. */
size_t ducosy_spatial_attention_scratch_bytes(int B, int H, int W);
int ducosy_spatial_attention_nchw(const float* x, const float* w, float* out, float* scratch, int B, int C, int H, int W, int ksize,
                                  ducosy_stream_t stream);
/* Layout converters between the reference's NCHW fp32 tensors and the NHWC 16-bit maps of the tensor-core path (used by the
 * stand-alone ResidualBlock / ResidualBlockWithCBAM forward, modules/model.py:56-87): fp32 [B][C][H][W] -> 16-bit
 * [B][H+2p][W+2p][C] with reflect / zero padding, and 16-bit [B][H][W][C] -> fp32 [B][C][H][W]. */
int ducosy_nchw_to_nhwc_pad(const float* x, void* out, int B, int C, int H, int W, int pad, int pad_mode, int dtype,
                            ducosy_stream_t stream);
int ducosy_nhwc_to_nchw(const void* y, float* out, int B, int C, int H, int W, int dtype, ducosy_stream_t stream);

/* ---------------------------------------------------------------- anatomical masks for training batches (first half)
 * The scipy.ndimage pieces of modules/mask_generator.py for batches of 2-D slices [B][H][W], bit-exact with scipy.
 * All four need `scratch` of ducosy_masks_scratch_bytes(B, H, W) bytes (256-byte aligned); B*H*W < 2^31.
 *
 * ducosy_label4: scipy.ndimage.label with the default structure (4-connectivity) per slice -- labels int32 [B][H][W] numbered
 * in raster order of each component's first pixel (scipy's numbering), num_features int32 [B] (may be NULL).
 * mask_generator.py:32,46,64,85,109,190,232. */
size_t ducosy_masks_scratch_bytes(int B, int H, int W);
int ducosy_label4(const uint8_t* mask, int32_t* labels, int32_t* num_features, int B, int H, int W, void* scratch,
                  size_t scratch_bytes, ducosy_stream_t stream);
/* scipy.ndimage.binary_fill_holes, default structure: background components (4-connectivity) that do not reach the slice
 * border become foreground.  out uint8 {0,1}.  mask_generator.py:69,90,242,310. */
int ducosy_binary_fill_holes(const uint8_t* mask, uint8_t* out, int B, int H, int W, void* scratch, size_t scratch_bytes,
                             ducosy_stream_t stream);
/* detect_lung (mask_generator.py:11-52): (lung_lower <= hu <= lung_upper) & (hu > -1000), `border_margin` rows / columns
 * cleared on every side, components smaller than `min_size` pixels removed.  hu fp32 [B][H][W] -> uint8 {0,1}. */
int ducosy_detect_lung(const float* hu, uint8_t* lung_mask, int B, int H, int W, float lung_lower, float lung_upper, int min_size,
                       int border_margin, void* scratch, size_t scratch_bytes, ducosy_stream_t stream);
/* detect_lung_vessels (mask_generator.py:55-99): on slices with >= 2 lung components, body_area > 0 and
 * lung_area / body_area >= 0.1 (float64), (binary_fill_holes(lung) - lung) & (vessel_lower <= hu <= vessel_upper); zero on the
 * other slices. */
int ducosy_detect_lung_vessels(const float* hu, const uint8_t* lung_mask, uint8_t* vessel_mask, int B, int H, int W,
                               float vessel_lower, float vessel_upper, void* scratch, size_t scratch_bytes, ducosy_stream_t stream);
/* Second half of modules/mask_generator.py: the convex hull of a slice's lung pixels and the two masks built on it.
 * ducosy_lung_hull: hull_mask = Path(ConvexHull(argwhere(lung == 1)).vertices).contains_points(every pixel)
 * (mask_generator.py:115-127); the vertices are scipy's (strict corners, counter-clockwise, (row, col)), the rasterisation is
 * matplotlib's crossings test restated (PARITY UNPINNED: matplotlib absent).  Slices with fewer than 3 lung pixels or a
 * degenerate (collinear) set get hull_mask = lung_mask and nverts = 0, the reference's fallback.  verts (optional): int32
 * [B][2H+4][2], nverts (optional): int32 [B].
 * ducosy_detect_mediastinum (mask_generator.py:100-170): (hull - lung) & (lower <= hu <= upper) on slices with >= 2 lung
 * components, body area > 0 and lung / body area >= 0.1.
 * ducosy_detect_bone (mask_generator.py:173-311): (hu >= threshold) & (hu > -1000), minus hull & ~lung above row
 * spine_start_row (= int(H * (1 - spine_margin_ratio)), computed by the caller in float64), components of the candidates that
 * still touch the remainder restored whole, holes filled. */
int ducosy_lung_hull(const uint8_t* lung_mask, uint8_t* hull_mask, int32_t* verts, int32_t* nverts, int B, int H, int W, void* scratch,
                     size_t scratch_bytes, ducosy_stream_t stream);
int ducosy_detect_mediastinum(const float* hu, const uint8_t* lung_mask, uint8_t* mediastinum_mask, int B, int H, int W, float lower,
                              float upper, void* scratch, size_t scratch_bytes, ducosy_stream_t stream);
int ducosy_detect_bone(const float* hu, const uint8_t* lung_mask, uint8_t* bone_mask, int B, int H, int W, float bone_threshold,
                       int spine_start_row, void* scratch, size_t scratch_bytes, ducosy_stream_t stream);

/* ---------------------------------------------------------------- whole-generator entry points */

typedef struct {
  int input_channels;      /* Generator(input_channels=...) modules/model.py:92 */
  int num_residual_blocks; /* default 9 */
  int use_cbam;            /* default 1 */
  int dtype;               /* DUCOSY_F16 | DUCOSY_BF16 (16-bit operand type of the tensor-core convs) | DUCOSY_F16X2 (split-operand arm) */
} ducosy_gen_config;

/* Number of parameter tensors in state_dict order (modules/model.py:92-113) and packed-cache size in bytes. */
int ducosy_generator_num_params(const ducosy_gen_config* cfg);
size_t ducosy_generator_packed_bytes(const ducosy_gen_config* cfg);
size_t ducosy_generator_workspace_bytes(const ducosy_gen_config* cfg, int B, int H, int W);
/* params_host: host array of DEVICE pointers to the fp32 parameters in state_dict order. */
int ducosy_generator_pack(const ducosy_gen_config* cfg, const float* const* params_host, int num_params, void* packed,
                          ducosy_stream_t stream);
/* Generator.forward (modules/model.py:114): x fp32 NCHW [B][Cin][H][W] -> out fp32 [B][1][H][W].
 * Shape contract (DUCOSY_ERR_SHAPE otherwise; ducosy_generator_workspace_bytes returns 0): B >= 1, 1 <= Cin <= 16, H a
 * multiple of 32 and W a multiple of 128, both >= 128, with W/4 in {32, 64, 128 k} and H/4 a multiple of 128 / min(W/4, 128)
 * (512 x 512 is what generate.py and train.py use).  workspace: 1024-byte aligned, packed: 256-byte aligned.
 * The TRAINING path built on the per-layer entry points (ducosy_gan_b200/training.py) additionally needs W = 256 or a
 * multiple of 512 for the generator and H, W multiples of 256 for the discriminator (input_channels == 1 only). */
int ducosy_generator_forward(const ducosy_gen_config* cfg, const void* packed, const float* x_nchw, float* out, int B,
                             int H, int W, void* workspace, size_t workspace_bytes, ducosy_stream_t stream);
/* Same with the input taken straight from stored pixel values through the HU window (generate.py:91-96), Cin = 1. */
int ducosy_generator_forward_hu(const ducosy_gen_config* cfg, const void* packed, const int16_t* px, float slope,
                                float intercept, float hu_lo, float hu_hi, float* out, int B, int H, int W,
                                void* workspace, size_t workspace_bytes, ducosy_stream_t stream);
/* Number of kernels one forward launches (for bench accounting). */
int ducosy_generator_num_launches(const ducosy_gen_config* cfg);

/* ---------------------------------------------------------------- loss terms of the CycleGAN step, forward + backward
 * (modules/trainer.py:22-184,347-358,469-512).  Images are fp32 [B][1][H][W]; loss_out / gout are DEVICE scalars (no host
 * sync); reductions are fixed-order (deterministic).  scratch: ducosy_loss_scratch_bytes() bytes, 16-byte aligned.
 * Backward functions return the gradient w.r.t. the first argument (the generated image), scaled by gout[0]. */
size_t ducosy_loss_scratch_bytes(void);
/* nn.L1Loss (cycle / identity, trainer.py:348-349) and nn.MSELoss against a constant patch target (trainer.py:347,459-460) */
int ducosy_loss_l1_forward(const float* a, const float* b, long long n, float* loss_out, float* scratch, ducosy_stream_t stream);
int ducosy_loss_l1_backward(const float* a, const float* b, long long n, const float* gout, float* da, ducosy_stream_t stream);
int ducosy_loss_mse_const_forward(const float* a, float target, long long n, float* loss_out, float* scratch, ducosy_stream_t stream);
int ducosy_loss_mse_const_backward(const float* a, float target, long long n, const float* gout, float* da, ducosy_stream_t stream);
/* GradientLoss (trainer.py:22-40) */
int ducosy_loss_gradient_forward(const float* pred, const float* target, int B, int H, int W, float* loss_out, float* scratch,
                                 ducosy_stream_t stream);
int ducosy_loss_gradient_backward(const float* pred, const float* target, int B, int H, int W, const float* gout, float* dpred,
                                  ducosy_stream_t stream);
/* ContrastAttentionLoss (trainer.py:43-86), blur kernel 7; umap [B*H*W] (may be NULL in forward-only use) feeds the backward */
int ducosy_loss_contrast_attention_forward(const float* pred, const float* target, const float* source, int B, int H, int W,
                                           float sigma, float min_w, float max_w, float* loss_out, float* umap, float* scratch,
                                           ducosy_stream_t stream);
int ducosy_loss_contrast_attention_backward(const float* umap, int B, int H, int W, const float* gout, float* dpred,
                                            ducosy_stream_t stream);
/* ContrastRegionLoss (trainer.py:89-130); state: 4 floats */
int ducosy_loss_contrast_region_forward(const float* pred, const float* target, const float* source, int B, int H, int W,
                                        float threshold, float weight, float* loss_out, float* state, float* scratch,
                                        ducosy_stream_t stream);
int ducosy_loss_contrast_region_backward(const float* pred, const float* target, const float* source, int B, int H, int W,
                                         float threshold, float weight, const float* state, const float* gout, float* dpred,
                                         ducosy_stream_t stream);
/* ContrastEdgeLoss (trainer.py:133-184): Sobel magnitude mean/std + exact top-10 % mean by radix selection instead of
 * torch.topk; ep/et: edge maps [B*H*W]; state: 8 floats */
int ducosy_loss_contrast_edge_forward(const float* pred, const float* target, int B, int H, int W, float* loss_out, float* ep,
                                      float* et, float* state, float* scratch, ducosy_stream_t stream);
int ducosy_loss_contrast_edge_backward(const float* pred, const float* ep, int B, int H, int W, const float* state,
                                       const float* gout, float* dpred, ducosy_stream_t stream);
/* mean SSIM (pytorch_msssim.SSIM(data_range, size_average=True, channel=1) as called at trainer.py:351,485 -- PARITY
 * UNPINNED, the package is absent from the reference tree): tmp 5*B*H*(W-10) floats, dmaps 3*B*(H-10)*(W-10) floats */
int ducosy_loss_ssim_forward(const float* x, const float* y, int B, int H, int W, float data_range, float* ssim_out, float* tmp,
                             float* dmaps, float* scratch, ducosy_stream_t stream);
int ducosy_loss_ssim_backward(const float* x, const float* y, const float* dmaps, int B, int H, int W, const float* gout, float* tmp,
                              float* dx, ducosy_stream_t stream);

/* ---------------------------------------------------------------- backward building blocks (training-step rows, in progress)
 *
 * Weight gradient of the NHWC convolutions on the tensor cores: dW[o][(r*kw+s)*Cin + c] = sum_pixels dy[p][o] *
 * x_pad[p*stride + (r,s)][c] (what autograd computes for modules/trainer.py:513,519,524).  x_pad is the padded input
 * the forward conv read, dy the output gradient [B][Ho+2*dy_pad][Wo+2*dy_pad][Cout] (interior), both 16-bit NHWC; dw fp32 in the packed forward
 * layout.  Cin in {64,128,192,256}, Cout = 64 or Cout % 128 == 0.  Deterministic (fixed-order reduction of the K splits). */
size_t ducosy_conv2d_wgrad_workspace_bytes(int B, int Ho, int Wo, int Cin, int Cout, int kh, int kw);
int ducosy_conv2d_wgrad_nhwc(const void* x_pad, const void* dy, int dy_pad, float* dw, int B, int Hp, int Wp, int Cin,
                             int Cout, int kh, int kw, int stride, void* workspace, size_t workspace_bytes, int dtype,
                             ducosy_stream_t stream);
/* The same, reduced straight into the parameter's own layout: dw_oihw fp32 [Cout][Cin][kh][kw], multiplied by gs[1] when gs
 * (the pair written by ducosy_grad_scale) is non-NULL: ducosy_conv2d_wgrad_nhwc + ducosy_unpack_wgrad in one reduction pass. */
int ducosy_conv2d_wgrad_nhwc_oihw(const void* x_pad, const void* dy, int dy_pad, float* dw_oihw, const float* gs, int B, int Hp,
                                  int Wp, int Cin, int Cout, int kh, int kw, int stride, void* workspace, size_t workspace_bytes,
                                  int dtype, ducosy_stream_t stream);
/* The same, ACCUMULATED into dw_oihw (dw += gradient * gs[1]) -- for a destination that already holds a gradient (the
 * parameter's .grad inside a flat data-parallel bucket): no temporary, no separate add. */
int ducosy_conv2d_wgrad_nhwc_oihw_acc(const void* x_pad, const void* dy, int dy_pad, float* dw_oihw, const float* gs, int B, int Hp,
                                      int Wp, int Cin, int Cout, int kh, int kw, int stride, void* workspace, size_t workspace_bytes,
                                      int dtype, ducosy_stream_t stream);

/* Weight gradient of Upsample(x2 nearest) + Conv3x3(pad 1) (modules/model.py:108-109) on the SOURCE grid: the gradient of
 * the 16 pre-summed (phase, tap) blocks of ducosy_pack_upconv_weight (4/9 of the MACs of a wgrad over the up-sampled map,
 * which is never materialised).  src_pad [B][Hs+2][Ws+2][Cin] zero border; dy [B][2Hs+2*dy_pad][2Ws+2*dy_pad][Cout], dy_pad
 * even; dwph fp32 [Cout][16*Cin].  ducosy_unpack_upconv_wgrad folds dwph back to the OIHW 3x3 gradient (times gs[1]). */
size_t ducosy_upconv2x_wgrad_workspace_bytes(int B, int Hs, int Ws, int Cin, int Cout);
int ducosy_upconv2x_wgrad_nhwc(const void* src_pad, const void* dy, int dy_pad, float* dwph, int B, int Hs, int Ws, int Cin,
                               int Cout, void* workspace, size_t workspace_bytes, int dtype, ducosy_stream_t stream);
int ducosy_unpack_upconv_wgrad(const float* dwph, float* g_oihw, int Cout, int Cin, const float* gs, ducosy_stream_t stream);

/* InstanceNorm(+activation) backward on NHWC 16-bit maps: forward n = y*scale + shift, a = act(n); given da returns
 * dy = rstd*(g - mean(g) - n*mean(g*n)), g = da*act'(n), written with a zero border of `pad` pixels (ready for the
 * phase / dgrad convolutions).  scratch: ducosy_in_backward_scratch_bytes. */
size_t ducosy_in_backward_scratch_bytes(int B, int H, int W, int C);
int ducosy_in_backward_pad(const void* da, const void* y, const float* scale, const float* shift, void* dy_pad, float* scratch,
                           int B, int H, int W, int C, int pad, int act, int dtype, ducosy_stream_t stream);
/* The same with the padding adjoint of a pad-1 convolution folded into the loads: da_pad1 [B][H+2][W+2][C] is the gradient with
 * respect to the PADDED map (the output of ducosy_conv3x3s1_dgrad_nhwc); fold_mode DUCOSY_PAD_REFLECT | DUCOSY_PAD_ZERO.
 * Replaces ducosy_pad_fold + ducosy_in_backward_pad for a map that has no other consumer: da_pad1 is CONSUMED (with
 * DUCOSY_PAD_REFLECT the mirrored border terms are first added into its interior cells, in place).  H >= 5, W a power of two >= 8. */
int ducosy_in_backward_pad_folded(void* da_pad1, int fold_mode, const void* y, const float* scale, const float* shift,
                                  void* dy_pad, float* scratch, int B, int H, int W, int C, int pad, int act, int dtype,
                                  ducosy_stream_t stream);
/* Input gradient of Conv2d(Cin, Cout, k, stride 2, padding 1), k = 4 (PatchGAN) or 3 (generator down convs,
 * modules/model.py:96-98), as four 2x2 phase convolutions over the zero-padded output gradient dy_pad [B][Ho+2][Wo+2][Cout]
 * -> dx [B][2Ho][2Wo][Cin]; w_dgrad from ducosy_pack_dgrad_s2_weight ([4*Cin][4*Cout], zero taps where k = 3 has none). */
int ducosy_pack_dgrad_s2_weight(const float* w_oihw, void* packed, int Cout, int Cin, int ksize, int dtype,
                                ducosy_stream_t stream);
int ducosy_convs2_dgrad_nhwc(const void* dy_pad, const void* w_dgrad, void* dx, int B, int Ho, int Wo, int Cin, int Cout,
                             int dtype, ducosy_stream_t stream);
/* Power-of-two scaling of a loss gradient for the 16-bit backward maps: gs[0] = 2^e with max|g|*2^e in [1,2), gs[1] = 2^-e.
 * The functions below take `gs` (device pointer, NULL = no scaling): dgrad entry points multiply by gs[0], the fp32
 * parameter / input gradients are multiplied by gs[1]. */
int ducosy_grad_scale(const float* g, long long n, float* gs, ducosy_stream_t stream);
/* Input gradient of the 3x3 stride-1 convs (modules/model.py:60-62,73-79) w.r.t. their PADDED input: dy_pad2
 * [B][H+4][W+4][Cout] (zero border 2) -> dxpad [B][H+2][W+2][Cin]; w_dgrad [Cin][9*Cout] from ducosy_pack_dgrad_s1_weight.
 * ducosy_pad_fold is the adjoint of ReflectionPad2d / zero padding: dxpad -> dx [B][H][W][C]. */
int ducosy_pack_dgrad_s1_weight(const float* w_oihw, void* packed, int Cout, int Cin, int dtype, ducosy_stream_t stream);
int ducosy_conv3x3s1_dgrad_nhwc(const void* dy_pad2, const void* w_dgrad, void* dxpad, int B, int H, int W, int Cin, int Cout,
                                int dtype, ducosy_stream_t stream);
int ducosy_pad_fold(const void* dxpad, void* dx, int B, int H, int W, int C, int pad, int pad_mode, int dtype,
                    ducosy_stream_t stream);
/* same, plus an element-wise addend [B][H][W][C] (the residual skip gradient of modules/model.py:65,87); add may be NULL */
int ducosy_pad_fold_add(const void* dxpad, const void* add, void* dx, int B, int H, int W, int C, int pad, int pad_mode,
                        int dtype, ducosy_stream_t stream);
/* Backward of Upsample(x2 nearest)+Conv3x3(pad 1) (modules/model.py:108-109): input gradient dy_pad2 [B][2Hs+4][2Ws+4][Cout]
 * (zero border 2) -> dsrc [B][Hs][Ws][Cin] with w_dgrad [Cin][16*Cout] from ducosy_pack_upconv_dgrad_weight; for the weight
 * gradient materialise the upsampled, zero-padded source with ducosy_upsample2x_pad and call ducosy_conv2d_wgrad_nhwc (3x3, s1). */
int ducosy_pack_upconv_dgrad_weight(const float* w_oihw, void* packed, int Cout, int Cin, int dtype, ducosy_stream_t stream);
int ducosy_upconv2x_dgrad_nhwc(const void* dy_pad2, const void* w_dgrad, void* dsrc, int B, int Hs, int Ws, int Cin, int Cout,
                               int dtype, ducosy_stream_t stream);
int ducosy_upsample2x_pad(const void* src_pad, void* up_pad, int B, int Hs, int Ws, int C, int dtype, ducosy_stream_t stream);
/* packed fp32 weight gradient [Cout][taps*Cin] -> OIHW [Cout][Cin][taps] (times gs[1]). */
int ducosy_unpack_wgrad(const float* packed, float* g_oihw, int Cout, int Cin, int taps, const float* gs,
                        ducosy_stream_t stream);
/* First (1->64) and last (512->1) discriminator layers, backward (modules/model.py:122,128). */
size_t ducosy_disc_last_backward_scratch_bytes(void);
int ducosy_disc_last_backward(const float* dout, const void* w5_packed, const void* p4, void* da4, float* dw5, float* db5,
                              float* scratch, const float* gs, int B, int Hs, int Ws, int dtype, ducosy_stream_t stream);
size_t ducosy_disc_first_backward_scratch_bytes(int B, int H, int W);
int ducosy_disc_first_backward(const void* da1, const void* p1, const float* x, const float* w1, float* dw1, float* db1,
                               float* dx, float* scratch, const float* gs, int B, int H, int W, int dtype,
                               ducosy_stream_t stream);

/* ---------------------------------------------------------------- PatchGAN discriminator forward / backward (modules/model.py:118-131) */

size_t ducosy_discriminator_packed_bytes(void);
size_t ducosy_discriminator_workspace_bytes(int B, int H, int W);
/* params_host: host array of the 10 DEVICE fp32 tensors model.{0,2,5,8,12}.{weight,bias} in state_dict order. */
int ducosy_discriminator_pack(const float* const* params_host, int num_params, void* packed, int dtype,
                              ducosy_stream_t stream);
/* Discriminator.forward: x fp32 [B][1][H][W] (H, W multiples of 256) -> out fp32 [B][1][H/16][W/16]. */
int ducosy_discriminator_forward(const void* packed, const float* x, float* out, int B, int H, int W, void* workspace,
                                 size_t workspace_bytes, int dtype, ducosy_stream_t stream);
/* Backward of that forward (loss.backward() of modules/trainer.py:518-524): fwd_workspace is the untouched workspace of the
 * forward call; grads_host = host array of 10 DEVICE fp32 buffers shaped like the parameters (overwritten); dx optional. */
size_t ducosy_discriminator_backward_workspace_bytes(int B, int H, int W);
int ducosy_discriminator_backward(const void* packed, const float* x, const float* dout, const void* fwd_workspace,
                                  float* const* grads_host, float* dx, int B, int H, int W, void* workspace,
                                  size_t workspace_bytes, int dtype, ducosy_stream_t stream);

/* ---------------------------------------------------------------- image-quality metrics over volumes
 * calculate.py:232-271,360-381: normalize, calculate_mae, calculate_psnr, calculate_ssim (skimage structural_similarity
 * defaults: 7x7 uniform window, sample covariance, K1 0.01, K2 0.03, valid region), calculate_cs, calculate_ed -- as two-stage
 * fixed-order float64 reductions.  a, b: [S][n] volumes of `in_type` (n = H*W); int16 volumes reproduce numpy's int16
 * wrap-around in (img1 - img2) and (img1 - img2)**2.
 *   stats [S][12] = per slice: sum|a-b|, sum(a-b)^2, sum ab, sum aa, sum bb, sum a, sum b, min a, max a, min b, max b, 0
 *   scratch: S * max(ducosy_metrics_chunks(n) * 12, ducosy_metrics_ssim_tiles(H, W)) doubles */
enum { DUCOSY_IN_I16 = 0, DUCOSY_IN_F32 = 1, DUCOSY_IN_F64 = 2 };
int ducosy_metrics_chunks(long long n);
int ducosy_metrics_slice_stats(const void* a, const void* b, int in_type, int S, long long n, double* stats, double* scratch,
                               ducosy_stream_t stream);
/* ed_sums[S] = sum over the slice of ((a - min a)/(range a + 1e-8) - (b - min b)/(range b + 1e-8))^2 (calculate.py:373-379);
 * `stats` from ducosy_metrics_slice_stats. */
int ducosy_metrics_ed(const void* a, const void* b, int in_type, int S, long long n, const double* stats, double* ed_sums,
                      double* scratch, ducosy_stream_t stream);
/* out[i] = (in[i] - minmax[0]) / (minmax[1] - minmax[0]) in float64, zeros when the range is zero (calculate.py:232-238);
 * minmax: 2 doubles in DEVICE memory. */
int ducosy_metrics_normalize(const void* in, int in_type, double* out, long long total, const double* minmax,
                             ducosy_stream_t stream);
/* calculate_emd (calculate.py:320-337) for int16 volumes, integer-exact: cdf_abs_sums[S] = sum_v |C1(v) - C2(v)| over the
 * R = global max - global min + 1 value bins (vmin = global min of both volumes); hist: S*2*R ZEROED uint32 counters.
 * The per-slice distance is cdf_abs_sums / n / (range + 1e-8), the reference's scaled value that / n again. */
int ducosy_metrics_emd_i16(const int16_t* a, const int16_t* b, int S, long long n, int vmin, int R, unsigned int* hist,
                           double* cdf_abs_sums, ducosy_stream_t stream);
/* calculate_ts (calculate.py:340-358): ts_stats[S][3] = per slice (sum |sobel(a) - sobel(b)|, max sobel(a), max sobel(b)) with
 * skimage.filters.sobel's arithmetic (PARITY UNPINNED: scikit-image absent); scratch: S * ceil(H / 8) * 3 doubles. */
int ducosy_metrics_ts(const void* a, const void* b, int in_type, int S, int H, int W, double* ts_stats, double* scratch,
                      ducosy_stream_t stream);
int ducosy_metrics_ssim_tiles(int H, int W);
/* ssim_sums[S] = sum of the SSIM map over the valid region [3, H-3) x [3, W-3) of each slice pair; mean = / ((H-6)(W-6)). */
int ducosy_metrics_ssim(const void* a, const void* b, int in_type, int S, int H, int W, double data_range, double* ssim_sums,
                        double* scratch, ducosy_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* DUCOSY_H_ */
