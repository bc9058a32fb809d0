/*
 * ideas_b200.h -- C ABI of the B200-native IDEAS GAN-conv hot path (libideas_b200.so).
 *
 * Drop-in boundary.  The reference reaches native code through two
 * pybind11 modules built at import time:
 *     fused.fused_bias_act(input, bias, refer, act, grad, alpha, scale)
 *         stylegan2/op/fused_bias_act.cpp:11-20, kernel fused_bias_act_kernel.cu:18-99
 *     upfirdn2d_op.upfirdn2d(input, kernel, up_x, up_y, down_x, down_y, pad_x0..pad_y1)
 *         stylegan2/op/upfirdn2d.cpp:12-22,   kernel upfirdn2d_kernel.cu:17-369
 * and delegates every convolution to cuDNN through torch.nn.functional
 *     F.conv2d / F.conv_transpose2d   stylegan2/model.py:115,258,267,273 ; models.py:32
 * This header declares the C entry points that replace those three native units, plus the
 * integer bit-path kernels that restate utils.py:74-97 / train.py:285.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to dense fp32 (or the stated integer type) memory,
 *     borrowed for the duration of the call; the CALLER allocates all outputs and
 *     workspaces (the host side uses the torch caching allocator), nothing is owned here;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, no call
 *     synchronises; the library is re-entrant and holds no mutable global state except a
 *     mutex-guarded cache of TMA descriptors and the thread-local last-error string;
 *   - return value 0 = success, negative = error (IDEAS_ERR_*); ideas_last_error() returns
 *     the message of the calling thread's last failure.  Unlike the reference (no shape
 *     checks, no post-launch check, 32-bit indices -- SURVEY.md §8b "Errors") arguments are
 *     validated, launches are checked with cudaGetLastError and element counts are 64-bit;
 *   - activations are NHWC: a tensor (N, H, W, C) with C contiguous.  The reference's own
 *     native view (major, H, W, minor) of upfirdn2d_kernel.cu:228-231 is this layout with
 *     major = N, minor = C (the reference always passes minor = 1, i.e. NCHW planes; both
 *     are accepted).
 */
#ifndef IDEAS_B200_H_
#define IDEAS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IDEAS_OK 0
#define IDEAS_ERR_INVALID (-1)   /* bad argument (shape, alignment, null pointer)        */
#define IDEAS_ERR_CUDA (-2)      /* CUDA runtime / driver error, see ideas_last_error() */
#define IDEAS_ERR_UNSUPPORTED (-3) /* valid request outside the implemented envelope     */

/* implementation selector of the convolution entry points */
#define IDEAS_IMPL_AUTO 0   /* tcgen05 when the shape qualifies, SIMT otherwise */
#define IDEAS_IMPL_SIMT 1   /* fp32 FFMA implicit GEMM (exact-fp32 differential reference) */
#define IDEAS_IMPL_UMMA 2   /* tcgen05.mma kind::tf32 + TMA; fails with UNSUPPORTED if not eligible */

/* activation codes of the conv epilogue */
#define IDEAS_ACT_NONE 0
#define IDEAS_ACT_LRELU 1   /* y = (v > 0 ? v : alpha*v) * gain */

int ideas_abi_version(void);
const char* ideas_last_error(void);
/* number of CUDA kernels this library has launched in the calling process (all threads) */
unsigned long long ideas_launch_count(void);
/* 1 when the tcgen05/TMA convolution kernels are compiled into this build */
int ideas_umma_available(void);
/* tuning knobs; "tma_tf32" = 1: tensor maps use the TFLOAT32 data type (TMA converts on load) */
int ideas_set_option(const char* name, int value);
/* compute capability major*10+minor of the current device, or a negative error */
int ideas_device_cc(void);

/* ------------------------------------------------------------------------------------
 * A3  fused bias + leaky ReLU          replaces fused_bias_act.cpp:11-20
 *
 * out[i] = f(x[i] + bias[(i / step_b) % size_b]) * scale with f chosen by act*10+grad
 * exactly as fused_bias_act_kernel.cu:36-47 (act 1 = linear, 3 = lrelu; grad 0 = forward,
 * 1 = first derivative masked by the sign of `ref`, 2 = second derivative = 0).
 * `bias` / `ref` may be NULL ("absent" = the reference's empty tensor, kernel.cu:62-63).
 * NHWC or (B, C) data: step_b = 1;  NCHW data: step_b = H*W.
 * ---------------------------------------------------------------------------------- */
int ideas_fused_bias_act(float* out, const float* x, const float* bias, const float* ref,
                         int act, int grad, float alpha, float scale,
                         int64_t n, int64_t step_b, int size_b, void* stream);

/* Backward of the forward activation in ONE pass (the reference needs two:
 * fused_act.py:29-38): gx = gy * (out > 0 ? 1 : alpha) * scale and
 * gbias[c] += sum over everything but the channel of gx.  `gbias` (size_b floats) must
 * be zero-initialised by the caller; it may be NULL to skip the reduction. */
int ideas_bias_act_backward(float* gx, float* gbias, const float* gy, const float* out,
                            float alpha, float scale, int64_t n, int64_t step_b, int size_b,
                            void* stream);

/* Backward of the modulated-conv epilogue out = lrelu(d[n,c]*u + bias[c]) * gain (NHWC, P pixels
 * per sample) in one pass over (gy, out):
 *   g1  = gy * (out > 0 ? 1 : alpha) * gain        g1d[n,p,c] = g1 * d[n,c]   (d may be NULL = 1)
 *   gsum[n,c] += sum_p g1                          dotz[n,c] += sum_p g1 * lrelu^-1(out)
 * so that dL/dbias = sum_n gsum and dL/dd = (dotz - bias*gsum)/d.  gsum/dotz zero-initialised
 * by the caller.  Restates the autograd chain of stylegan2/model.py:245-248,375 + fused_act.py:20-49. */
int ideas_modconv_act_backward(float* g1d, float* gsum, float* dotz, const float* gy, const float* out,
                               const float* d, float alpha, float gain, int N, int64_t P, int C,
                               void* stream);

/* ------------------------------------------------------------------------------------
 * A4  upfirdn2d                          replaces upfirdn2d.cpp:12-22
 *
 * Same contract as upfirdn2d_op(): input viewed (major, in_h, in_w, minor), `kernel`
 * (kernel_h, kernel_w) un-flipped as the caller holds it, output (major, out_h, out_w,
 * minor) with out = (in*up + pad0 + pad1 - k)/down + 1 (upfirdn2d_kernel.cu:236-239).
 * Negative pads crop.  Optional fused epilogue (B200 addition): if `bias` != NULL the
 * result goes through (v + bias[minor_idx]) -> lrelu(alpha) -> * gain, which is the
 * Blur -> FusedLeakyReLU pair of an upsampling StyledConv (stylegan2/model.py:261,375).
 * ---------------------------------------------------------------------------------- */
int ideas_upfirdn2d(float* out, const float* x, const float* kernel,
                    int major, int in_h, int in_w, int minor, int kernel_h, int kernel_w,
                    int up_x, int up_y, int down_x, int down_y,
                    int pad_x0, int pad_x1, int pad_y0, int pad_y1,
                    const float* bias, float alpha, float gain, void* stream);
/* Same with a residual merged into the result: out = (upfirdn2d(x) + residual) * res_scale, `residual` of out's
 * shape.  Served by the fused up-sampling path only (up = 2, down = 1, <= 4x4 kernel, channels % 4 == 0, no bias):
 * the (out + skip)/sqrt(2) of an up-sampling residual block, whose skip path ends in this blur (models.py:78-95,
 * :178).  Other parameter sets return IDEAS_ERR_UNSUPPORTED. */
int ideas_upfirdn2d_res(float* out, const float* x, const float* kernel,
                        int major, int in_h, int in_w, int minor, int kernel_h, int kernel_w,
                        int up_x, int up_y, int down_x, int down_y,
                        int pad_x0, int pad_x1, int pad_y0, int pad_y1,
                        const float* bias, float alpha, float gain,
                        const float* residual, float res_scale, void* stream);

/* Backward of the pair  y = act(u) -> z = Blur(y)  (conv1 -> FusedLeakyReLU -> Blur of a down-sampling
 * ResBlock, reference models.py:213-227) in one pass:
 *   gx = upfirdn2d(g, kernel, pads) * (ref > 0 ? 1 : alpha) * gain,   gbias[c] += sum of gx over all but the channel
 * `g` (major, in_h, in_w, minor) is the gradient w.r.t. z, `kernel` / pads are those of the blur's own backward
 * (flipped kernel, upfirdn2d.py:31-42), `ref` = y has the shape of gx.  Replaces UpFirDn2dBackward followed by
 * FusedLeakyReLUFunctionBackward (upfirdn2d.py:19-42, fused_act.py:20-49) without writing the blurred gradient.
 * gbias may be NULL; IDEAS_ERR_UNSUPPORTED outside the fast-path envelope (caller runs the two ops). */
int ideas_blur_act_backward(float* gx, float* gbias, const float* g, const float* ref, const float* kernel,
                            int major, int in_h, int in_w, int minor, int kernel_h, int kernel_w,
                            int pad_x0, int pad_x1, int pad_y0, int pad_y1, float alpha, float gain, void* stream);

/* Backward of  u -> d[n,c]*u -> Blur  (up-sampling ModulatedConv2d, stylegan2/model.py:250-261) in one pass:
 * v = upfirdn2d(g, kernel [already flipped], gradient pads); gx = v * scale[n,c] (scale may be NULL);
 * dot[n,c] += sum over pixels of v * ref (ref has gx's shape; dot (major x minor) zero-initialised by the caller).
 * Fast path only (<= 4x4 kernel, channels % 4 == 0, power-of-two channel count): IDEAS_ERR_UNSUPPORTED otherwise. */
int ideas_blur_scale_dot_backward(float* gx, float* dot, const float* g, const float* ref, const float* scale,
                                  const float* kernel, int major, int in_h, int in_w, int minor,
                                  int kernel_h, int kernel_w, int pad_x0, int pad_x1, int pad_y0, int pad_y1,
                                  void* stream);

/* Blur -> FusedLeakyReLU (stylegan2/model.py:261,375) with a second output for the next layer:
 * out = gain*lrelu(upfirdn2d(x, kernel, pads) + bias[c]); out2 = out * post[n,c] (post: major x minor).
 * up = down = 1 fast path only; IDEAS_ERR_UNSUPPORTED otherwise. */
int ideas_blur_bias_act_post(float* out, float* out2, const float* x, const float* kernel, const float* bias,
                             const float* post, int major, int in_h, int in_w, int minor,
                             int kernel_h, int kernel_w, int pad_x0, int pad_x1, int pad_y0, int pad_y1,
                             float alpha, float gain, void* stream);

/* ------------------------------------------------------------------------------------
 * A1/A5  convolutions (new native unit; the reference calls cuDNN here)
 *
 * Weights are consumed in the PACKED layout  wp[tap][K][C]  (tap = kh*kw_count + kw,
 * K = output channels of the GEMM, C = its reduction channels, C contiguous), produced
 * by ideas_pack_weight from the module parameter:
 *   src OIHW (O,I,kh,kw)  -> dst[tap][o][i] = scale * src[o][i][tap]          (transpose=0, flip=0)
 *   the dgrad operand     -> dst[tap'][i][o] = scale * src[o][i][tap]         (transpose=1, flip=1,
 *                            tap' = kh*kw-1-tap)
 * For an IOHW parameter (EqualConvTranspose2d, models.py:17-19) swap the meaning of O and I.
 * ---------------------------------------------------------------------------------- */
int ideas_pack_weight(float* dst, const float* src, int O, int I, int kh, int kw,
                      int transpose, int flip, float scale, void* stream);
/* wpt[taps-1-t][c][k] = wp[t][k][c] : the dgrad operand derived from a packed forward weight */
int ideas_repack_dgrad(float* wpt, const float* wp, int K, int C, int taps, void* stream);
/* inverse packing for gradients: dst OIHW[o][i][tap] (+)= scale * src[tap][o][i]  (transpose=0)
 *                                                      or scale * src[tap][i][o]  (transpose=1) */
int ideas_unpack_weight_grad(float* dst, const float* src, int O, int I, int kh, int kw,
                             int transpose, float scale, int accumulate, void* stream);

/* Forward:  y[n,oy,ox,k] = epilogue( out_scale[n,k] * sum_{kh,kw,c} wp[kh,kw][k][c] *
 *                                    in_scale[n,c] * x[n, oy*stride+kh-pad, ox*stride+kw-pad, c] )
 * epilogue(v) = act(v + bias[k]) (bias / in_scale / out_scale may be NULL).
 * x (N,H,W,C), y (N,OH,OW,K) with OH = (H + 2*pad - kh)/stride + 1.
 * With in_scale = style s and out_scale = demodulation d this is ModulatedConv2d
 * (stylegan2/model.py:236-277) without materialising per-sample weights. */
int ideas_conv2d_forward(float* y, const float* x, const float* wp,
                         const float* in_scale, const float* out_scale, const float* bias,
                         int N, int H, int W, int C, int K, int kh, int kw, int stride, int pad,
                         int act, float alpha, float gain, int impl, void* stream);
/* Same, merging a residual after the activation: y = (epilogue(..) + residual) * res_scale, with
 * `residual` of y's shape (NULL = plain forward).  This is the (out + skip)/sqrt(2) of the residual
 * blocks (models.py:178,227) done in the convolution's epilogue instead of a pass of its own. */
int ideas_conv2d_forward_res(float* y, const float* x, const float* wp,
                             const float* in_scale, const float* out_scale, const float* bias,
                             const float* residual, float res_scale,
                             int N, int H, int W, int C, int K, int kh, int kw, int stride, int pad,
                             int act, float alpha, float gain, int impl, void* stream);

/* Data gradient of the forward above == transposed convolution:
 *   dx[n,iy,ix,c] = in_scale[n,c] * sum_{kh,kw,k} wpt[kh',kw'][c][k] * out_scale[n,k] * dy[n,oy,ox,k]
 * over all (oy,kh) with oy*stride + kh - pad = iy.  `wpt` is the transpose=1, flip=1 packing.
 * dy (N,OH,OW,K) -> dx (N,H,W,C); H, W are given explicitly (output_padding is implied).
 * Used as: conv dgrad, and the FORWARD of conv_transpose2d (stylegan2/model.py:258, models.py:32)
 * with optional bias / activation epilogue. */
int ideas_conv2d_dgrad(float* dx, const float* dy, const float* wpt,
                       const float* in_scale, const float* out_scale, const float* bias,
                       int N, int H, int W, int C, int K, int kh, int kw, int stride, int pad,
                       int OH, int OW, int act, float alpha, float gain, int impl, void* stream);

/* Weight gradient in packed layout (zero-initialised by the caller, accumulated with fp32 adds):
 *   dwp[kh,kw][k][c] += sum_{n,oy,ox} out_scale[n,k]*dy[n,oy,ox,k] * in_scale[n,c]*x[n,oy*stride+kh-pad, ox*stride+kw-pad, c] */
int ideas_conv2d_wgrad(float* dwp, const float* x, const float* dy,
                       const float* in_scale, const float* out_scale,
                       int N, int H, int W, int C, int K, int kh, int kw, int stride, int pad,
                       int OH, int OW, int impl, void* stream);

/* ------------------------------------------------------------------------------------
 * NHWC helpers used by the modulation path (stylegan2/model.py:239-248)
 * ---------------------------------------------------------------------------------- */
/* out[n,p,c] = x[n,p,c] * s[n,c]   (P = pixels per sample) */
int ideas_scale_channels(float* out, const float* x, const float* s, int N, int64_t P, int C,
                         void* stream);
/* dot[n,c] += sum_p a[n,p,c]*b[n,p,c];  if out != NULL also out[n,p,c] = b[n,p,c]*s[n,c].
 * `dot` must be zero-initialised by the caller. */
int ideas_channel_dot(float* dot, float* out, const float* a, const float* b, const float* s,
                      int N, int64_t P, int C, void* stream);
/* ------------------------------------------------------------------------------------
 * A6  EqualLinear GEMM (reference: F.linear -> cuBLAS, stylegan2/model.py:151-161), exact fp32 FFMA.
 *   C[i*ldc + j] = alpha * sum_{r<R} a[i*sa_i + r*sa_r] * b[j*sb_j + r*sb_r]        i < M, j < N
 * Strided operands make forward (x W^T), input gradient (dY W) and weight gradient (dY^T X) one
 * kernel.  c_is_zero != 0 promises a zero-initialised C and allows a split reduction merged with
 * fp32 atomics (long R, few tiles). */
int ideas_gemm_nt(float* c, const float* a, const float* b, int M, int N, int R,
                  int64_t sa_i, int64_t sa_r, int64_t sb_j, int64_t sb_r, int64_t ldc,
                  float alpha, int c_is_zero, void* stream);

/* out = (a + b) * gain   (the residual merge (out+skip)/sqrt(2), models.py:178,227) */
int ideas_add_scale(float* out, const float* a, const float* b, float gain, int64_t n, void* stream);

/* nn.ReflectionPad2d(pad) on NHWC data (reference models.py:102-108: reflect-padded 3x3 convs of the encoder,
 * structure-generator and extractor ResBlocks).  backward = 0: out (N, H+2p, W+2p, C) from x (N, H, W, C);
 * backward = 1: out (N, H, W, C) = gradient w.r.t. the un-padded tensor from x = gradient (N, H+2p, W+2p, C).
 * H and W are always the UN-padded sizes. */
int ideas_reflect_pad2d(float* out, const float* x, int N, int H, int W, int C, int pad, int backward, void* stream);

/* ------------------------------------------------------------------------------------
 * A10 / SURVEY §8(f)-1  patchify: n_crop boxes (device int32 (n_crop,4) = y,x,h,w, shared by the batch)
 * cut from every NHWC image and resized to (th,tw) with F.interpolate(bilinear, align_corners=False)
 * semantics -- replaces the Python loop of utils.py:127-149.  out (B*n_crop, th, tw, C), crops of one
 * image adjacent.  The backward scatters with fp32 atomics into the zero-initialised `gimg`.
 * ---------------------------------------------------------------------------------- */
int ideas_patchify_forward(float* out, const float* img, const int* boxes, int B, int H, int W, int C,
                           int n_crop, int th, int tw, void* stream);
int ideas_patchify_backward(float* gimg, const float* gout, const int* boxes, int B, int H, int W, int C,
                            int n_crop, int th, int tw, void* stream);

/* ------------------------------------------------------------------------------------
 * A9  bit path as integer kernels        restates utils.py:74-97, train.py:285
 * ---------------------------------------------------------------------------------- */
/* z[b,l] = step*(n+0.5) - 1 + (u*r*2 - r), n = the sigma bits of group l, MSB first.
 * bits: packed uint32, bit j of word w of row b = message bit 32*w+j; u: uniform draws (may be
 * NULL = 0.5, i.e. no jitter).  L groups per row, words_per_row = ceil(sigma*L/32). */
int ideas_bits_encode(float* z, const uint32_t* bits, const float* u, int B, int L, int sigma,
                      float delta, void* stream);
/* threshold peeling of utils.py:86-97 with its exact fp32 op order; writes packed bits */
int ideas_bits_decode(uint32_t* bits, const float* z, int B, int L, int sigma, void* stream);
/* *errors (uint64, zero-initialised by caller) += popcount(a xor b) over n_words words */
int ideas_bits_count_errors(unsigned long long* errors, const uint32_t* a, const uint32_t* b,
                            int64_t n_words, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* IDEAS_B200_H_ */
