/*
 * ddpmir.h -- C ABI of the B200-native (sm_100a) DDPM image-restoration hot path.
 *
 * The reference (Azure0413/DDPM_Image_Restoration) has no FFI / plugin layer: its only interface is a set of
 * Python classes and functions whose device work goes through stock PyTorch ops.  Each entry point below
 * therefore replaces one *op call site* of the reference's hot path (file:line given per function) and is what
 * the Python classes in ddpm_image_restoration_b200/ bind through ctypes (see INTEGRATION.md for the stub a
 * reference maintainer would add).
 *
 * Conventions
 *   - plain pointers + sizes only; every pointer is DEVICE memory unless its name ends in _host.
 *   - the caller owns all buffers (outputs and workspaces); nothing here allocates, frees or synchronises.
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); calls are asynchronous.
 *   - return value: DDPMIR_OK (0) or a negative DDPMIR_ERR_* code; ddpmir_last_error() gives the text.
 *   - activations inside the UNet are NHWC ("pixel-major"): [B, H, W, C] with C contiguous, element type
 *     selected by `dtype` (DDPMIR_BF16 = production, DDPMIR_F32 = fp32 check mode).  Sampler-level images
 *     are the reference's NCHW fp32 tensors in [-1, 1].
 *   - weights are [N, K] row-major ("K-major") in the activation dtype; for 3x3 convolutions
 *     K = 9 * Cin ordered (kh, kw, cin)  (pre-packed once on the host side from the checkpoint's OIHW fp32).
 */
#ifndef DDPMIR_H
#define DDPMIR_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DDPMIR_OK 0
#define DDPMIR_ERR_INVALID (-1)
#define DDPMIR_ERR_CUDA (-2)
#define DDPMIR_ERR_UNSUPPORTED (-3)

#define DDPMIR_F32 0
#define DDPMIR_BF16 1
#define DDPMIR_F16 2 /* binary16: only as a GEMM epilogue OUTPUT format and as the qkv of ddpmir_attention_prescaled_f16 */

#define DDPMIR_ACT_NONE 0
#define DDPMIR_ACT_RELU 1
#define DDPMIR_ACT_LRELU02 2
#define DDPMIR_ACT_SIGMOID 3
#define DDPMIR_ACT_SILU 4
#define DDPMIR_ACT_GELU 5
#define DDPMIR_ACT_TANH 6

/* kernel family selector for ops that have more than one implementation */
#define DDPMIR_IMPL_AUTO 0
#define DDPMIR_IMPL_SIMT 1   /* generic fp32-FMA kernels (the fp32 check mode; any shape) */
#define DDPMIR_IMPL_TENSOR 2 /* tensor-core kernels (bf16 only; shape constraints apply) */

typedef void* ddpmir_stream_t;

int ddpmir_version(void);
const char* ddpmir_last_error(void);

/* ------------------------------------------------------------------------------------------------------ */
/* Sampler-level kernels (NCHW fp32 images)                                                                */
/* ------------------------------------------------------------------------------------------------------ */

/* DDRM update, webp_inference.py:584-592,600 (avif_inference.py:441-449,457; svd.ipynb#c1:L367-375,383):
 *     x' = x_theta - codec + y
 *     out = eta_b*x' + (1-eta_b)*x_theta + eta*(z * (t[b]*sigma_scale))      (last_step: out = x')
 * evaluated in the reference's own operation order with unfused fp32 ops, so it is bit-exact for an injected z.
 * `codec` is either the reference's fp32 NCHW tensor (codec_u8_hwc = 0) or the decoder's raw uint8 HWC pixels
 * (codec_u8_hwc = 1; converted as ToTensor + sub(0.5).mul(2.0) would, webp_inference.py:524-528).
 * z == NULL: z is generated in-kernel, Philox4x32-10 keyed by (seed, step, noise_offset + flat NCHW element index)
 * -- see ddpmir_philox_normal; noise_offset (a multiple of 4) lets a micro-batch address its slice of the
 * whole-batch noise stream.  t is [B].  The scalars arrive as the Python doubles of the reference's signature and
 * are rounded to fp32 exactly where its tensor ops would ((1-eta_b) is formed in double first). */
int ddpmir_ddrm_update(const float* x_theta, const void* codec, int codec_u8_hwc, const float* y, const float* z,
                       const float* t, float* out, int B, int C, int H, int W, double sigma_scale, double eta,
                       double eta_b, int last_step, uint64_t seed, uint32_t step, uint64_t noise_offset,
                       ddpmir_stream_t stream);

/* GaussianMixtureSampler step, 0409_method.ipynb#c1:L411-447:
 *     if svd_prior: pred = (1-g)*pred + g*(y - svd_prior)
 *     last_step:    out = x_t + pred
 *     else:         x0 = x_t + pred ; mean = use_first ? 0.9*x0 + 0.1*x_t : 1.1*x0 - 0.1*x_t ; out = mean + noise_scale*z
 * z == NULL -> in-kernel Philox as above. */
int ddpmir_gmm_update(const float* x_t, const float* pred, const float* y, const float* svd_prior, float g,
                      const float* z, float* out, int64_t n, int use_first, float noise_scale, int last_step,
                      uint64_t seed, uint32_t step, ddpmir_stream_t stream);

/* Generic out = wa*a + wb*b (+ sigma*z): covers the classical DDPM posterior mean of
 * experiments/code/ddpm.ipynb#c5:L63-76 (a = x_t, b = eps).  b may be NULL; z NULL + sigma != 0 -> Philox. */
int ddpmir_lincomb(const float* a, float wa, const float* b, float wb, const float* z, float sigma, float* out,
                   int64_t n, uint64_t seed, uint32_t step, ddpmir_stream_t stream);

/* Standard normals from Philox4x32-10: element e takes word e%4 of Philox(counter=(e/4 mod 2^32, step, e/4 >> 32, 0),
 * key=(seed_lo, seed_hi)); u = ((w >> 9) + 0.5) * 2^-23; Box-Muller on word pairs (0,1) and (2,3).
 * Replaces torch.randn_like(x_t), webp_inference.py:589. */
int ddpmir_philox_normal(float* out, int64_t n, uint64_t seed, uint32_t step, ddpmir_stream_t stream);

/* (x*127.5+127.5).clamp(0,255).to(uint8) -- truncating -- and NCHW -> HWC, webp_inference.py:509,514.
 * out is [B, H, W, C] uint8 (device memory or mapped pinned host memory). */
int ddpmir_quantize_u8_hwc(const float* x, uint8_t* out, int B, int C, int H, int W, ddpmir_stream_t stream);

/* Inverse direction for the decoder's pixels: ToTensor() then .sub(0.5).mul(2.0), webp_inference.py:524-528.
 * in [B, H, W, C] uint8 -> out [B, C, H, W] fp32. */
int ddpmir_u8_hwc_to_nchw(const uint8_t* in, float* out, int B, int C, int H, int W, ddpmir_stream_t stream);

/* phase_consistency(x, ref, alpha), webp_inference.py:531-550, per [H, W] plane (H, W powers of two <= 1024).
 * ddpmir_phase_reference caches exp(i*angle(fft2(ref))) once per trajectory (ref = y is constant);
 * ddpmir_phase_consistency then does fft2(x) -> |.|*phasor -> ifft2 -> alpha*x + (1-alpha)*real.
 * phasor and ws are [planes, H, W] complex64 (2 floats per element). */
int ddpmir_phase_reference(const float* ref, int planes, int H, int W, float* phasor, float* ws,
                           ddpmir_stream_t stream);
int ddpmir_phase_consistency(const float* x, const float* phasor, float alpha, int planes, int H, int W,
                             float* out, float* ws, ddpmir_stream_t stream);

/* svd_structure_preservation, 0409_method.ipynb#c0:L321-346: rank-k truncation of every [H, W] plane
 * (one-sided Jacobi on the rows, one CTA per plane).  ws: planes * (H*W + H*H + 2*H) floats.  sweeps <= 0 -> 30
 * (it stops early once every row pair is orthogonal to fp32 round-off). */
int ddpmir_svd_lowrank(const float* x, int planes, int H, int W, int k, float* out, float* ws, int sweeps,
                       ddpmir_stream_t stream);

/* channel-weighted L1 of color_preservation_loss / color_loss on clamped [0,1] images,
 * 0409_method.ipynb#c0:L66-76, conv_deep.ipynb#c0:L60-73.  out_scalar[0] = 0.25 L1_R + 0.5 L1_G + 0.25 L1_B.
 * ws: 3 doubles. */
int ddpmir_color_l1(const float* pred, const float* target, int B, int H, int W, float* out_scalar, double* ws,
                    ddpmir_stream_t stream);

/* Forward values of the remaining loss terms (no gradients yet: the training step's backward kernels are a later
 * round).  ddpmir_mse: F.mse_loss(a, b), webp_training.py:108.  ws: 1 double. */
int ddpmir_mse(const float* a, const float* b, int64_t n, float* out_scalar, double* ws, ddpmir_stream_t stream);

/* pytorch_msssim.ssim(x*0.5+0.5, y*0.5+0.5, data_range=1.0, size_average=True) (gaussian window 11, sigma 1.5,
 * 'valid' separable filtering, K=(0.01,0.03)); clamp01 != 0 clamps the [0,1] images first (0409_method.ipynb#c0:L67-68,79;
 * webp_training.py:111-112,129 does not clamp).  x, y: [planes, H, W] fp32 in [-1,1].  ws: 1 double.
 * PARITY UNPINNED: the third-party package is not installed and the reference pins no version. */
int ddpmir_ssim(const float* x, const float* y, int planes, int H, int W, int clamp01, float* out_scalar, double* ws,
                ddpmir_stream_t stream);

/* Frequency terms of frequency_aware_loss, webp_training.py:114-126, over all planes at once:
 * acc2[0] = sum (|rfft2 P| - |rfft2 T|)^2, acc2[1] = sum (angle rfft2 P - angle rfft2 T)^2 with P = pred*0.5+0.5,
 * T = target*0.5+0.5 (H, W powers of two).  ws_pred / ws_target: [planes, H, W] complex64 each. */
int ddpmir_freq_loss_terms(const float* pred, const float* target, int planes, int H, int W, float* ws_pred,
                           float* ws_target, double* acc2, ddpmir_stream_t stream);

/* The same two sums over the FULL fft2 spectrum, the frequency terms of avif_frequency_aware_loss (avif.py:148-158:
 * torch.fft.fft2 instead of rfft2).  Computed on the half spectrum with the mirrored columns counted twice. */
int ddpmir_fft2_loss_terms(const float* pred, const float* target, int planes, int H, int W, float* ws_pred,
                           float* ws_target, double* acc2, ddpmir_stream_t stream);

/* gradient_loss of avif_frequency_aware_loss (avif.py:136-146) on the [0,1] images: acc2[0] = sum over the H-1 vertical
 * neighbour pairs of (|x - x_down| - |y - y_down|)^2, acc2[1] = the same over the W-1 horizontal pairs; the two F.mse_loss
 * means are acc2[0] / (planes (H-1) W) and acc2[1] / (planes H (W-1)). */
int ddpmir_edge_loss(const float* pred, const float* target, int planes, int H, int W, double* acc2, ddpmir_stream_t stream);

/* ------------------------------------------------------------------------------------------------------ */
/* UNet kernels                                                                                            */
/* ------------------------------------------------------------------------------------------------------ */

/* TimeEmbedding.forward, webp_inference.py:145-151: sin/cos features of t -> Linear -> SiLU -> Linear.
 * w0 [4*dim, dim], w1 [dim, 4*dim] fp32 (checkpoint layout). ws: B*5*dim floats. out [B, dim] fp32. */
int ddpmir_time_embed(const float* t, int B, int dim, const float* w0, const float* b0, const float* w1,
                      const float* b1, float* ws, float* out, ddpmir_stream_t stream);

/* Only the sinusoidal features [B, dim] of TimeEmbedding (the training step keeps the MLP's pre-activations). */
int ddpmir_time_features(const float* t, int B, int dim, float* out, ddpmir_stream_t stream);

/* out[r, n] = act(bias[n] + sum_k in[r,k] * w[n,k]) for a few rows (time_proj webp_inference.py:308; the pooled
 * multi-scale gates avif_inference.py:193-201).  All fp32. */
int ddpmir_linear_rows(const float* in, int rows, int K, const float* w, const float* bias, int N, int act,
                       float* out, ddpmir_stream_t stream);

/* GroupNorm statistics, nn.GroupNorm webp_inference.py:281,290,363 (eps 1e-5, biased variance).
 * x is NHWC (nchw = 0) or the NCHW fp32 network input (nchw = 1, dtype must be F32).
 * mean_rstd: [B, G, 2] fp32.  ws: B*G*2 doubles. */
int ddpmir_groupnorm_stats(const void* x, int dtype, int nchw, int B, int HW, int C, int G, float eps,
                           float* mean_rstd, double* ws, ddpmir_stream_t stream);

/* y = act((x - mean) * rstd * gamma[c] + beta[c]) on NHWC; act in {NONE, GELU, SILU} (webp_inference.py:304,
 * 311-312, 363-364).  x in `dtype`, out in `out_dtype`; raw_copy (optional, out_dtype) receives x itself cast to
 * out_dtype -- the operand copy of the fp32 residual stream that the 1x1 shortcut GEMM consumes. */
int ddpmir_groupnorm_apply(const void* x, int dtype, int B, int HW, int C, int G, const float* mean_rstd,
                           const float* gamma, const float* beta, int act, void* out, int out_dtype, void* raw_copy,
                           ddpmir_stream_t stream);

/* Convolution of the 3-channel NCHW fp32 network input (conv1 and the 1x1 shortcut of down1,
 * webp_inference.py:282,301,337), optionally folding norm1 (mean_rstd/gamma/beta non-NULL):
 * out[b,h,w,n] = bias[n] + row_bias[b,n] + sum w[n,c,kh,kw] * gn(x)[b,c,h+kh-p,w+kw-p]   (zero padding after GN).
 * w is the checkpoint's OIHW fp32; ksize 1 or 3; out is NHWC `dtype`. */
int ddpmir_conv_input(const float* x, int B, int Cin, int H, int W, const float* mean_rstd, const float* gamma,
                      const float* beta, const float* w, const float* bias, const float* row_bias, int N,
                      int ksize, int dtype, void* out, ddpmir_stream_t stream);

/* Epilogue shared by ddpmir_conv3x3 and ddpmir_gemm; m = flat pixel index (b, h, w), n = output channel.
 * The `dtype` argument of those functions is the OPERAND type (activations in, weights); the result can be
 * stored as fp32 (the residual stream, which only element-wise/normalisation kernels read) and/or as bf16 (the
 * next GEMM's operand): out in out_dtype, plus an optional second copy out2 in out2_dtype.  mul / res carry
 * their own dtypes.  A NULL epilogue means "store acc in the operand dtype".
 *     v = acc + bias[n] (+ row_bias[b, n])                      bias2 replaces bias on high-frequency pixels (freq_mode 2)
 *     v = act(v)
 *     freq_mode 1:  v = 0 unless (n < N/2) == is_low(h, w)      hidden layer of the stacked low/high gate MLP
 *     freq_mode 2:  v *= is_low ? 1 : img_scale[b]              high_boost, webp_inference.py:263-264
 *     otherwise  :  v *= img_scale[b]  (if given)
 *     v *= mul[m, n] (if given);  v += res[m, n] (if given)
 * is_low(h, w) restates the block loop of webp_inference.py:241-252 with block size `bs` and low size `low`. */
typedef struct {
    const float* bias;
    const float* bias2;
    const float* row_bias;
    const float* img_scale;
    const void* mul;
    const void* res;
    void* out2;
    int act;
    int freq_mode;
    int bs;
    int low;
    int out_dtype;
    int out2_dtype;
    int mul_dtype;
    int res_dtype;
} ddpmir_epilogue_t;

/* nn.Conv2d(Cin, N, 3, padding=1) on NHWC as an implicit GEMM (webp_inference.py:282,292,229; the AVIF edge
 * convs avif_inference.py:213-215).  Cin % 16 == 0.  w: [N, 9*Cin] (kh,kw,cin). */
int ddpmir_conv3x3(const void* x, int dtype, int B, int H, int W, int Cin, const void* w, int N,
                   const ddpmir_epilogue_t* epi, void* out, int impl, ddpmir_stream_t stream);

/* 1x1 convolutions / nn.Linear over pixels (shortcut :301, gate MLPs :215-225, MHA in/out projections :295):
 * a is [B*H*W, K], w is [N, K].  K % 16 == 0. */
int ddpmir_gemm(const void* a, int dtype, int B, int H, int W, int K, const void* w, int N,
                const ddpmir_epilogue_t* epi, void* out, int impl, ddpmir_stream_t stream);

/* Self-attention core of nn.MultiheadAttention(C, heads, batch_first=True) over L = H*W tokens
 * (webp_inference.py:317-319): qkv is the in_proj output [B, L, 3C] (q | k | v, head h owns channels
 * [h*hd, (h+1)*hd)); out [B, L, C] = softmax(q k^T / sqrt(hd)) v, never materialising the L x L scores. */
int ddpmir_attention(const void* qkv, int dtype, int B, int L, int C, int heads, void* out, int impl,
                     ddpmir_stream_t stream);

/* Same result for the inference path (bf16 only) when the q third of qkv was produced ALREADY multiplied by
 * log2(e)/sqrt(hd) (the factor is folded into in_proj_weight/bias rows [0, C) when the weights are packed, so it
 * costs no extra rounding).  head_dim 8/16 with L % 64 == 0 take the bounded-softmax kernel: the per-row offset of
 * the softmax is a Cauchy-Schwarz bound fixed before the first key instead of a running maximum (see attn_mma.cu);
 * rows whose bound is too large for fp32 are redone by the exact kernel, other shapes go to the exact kernels.
 * workspace: ddpmir_attention_prescaled_workspace(B, L, heads) bytes of device memory, caller-owned. */
size_t ddpmir_attention_prescaled_workspace(int B, int L, int heads);
int ddpmir_attention_prescaled(const void* qkv, int B, int L, int C, int heads, void* workspace, void* out,
                               ddpmir_stream_t stream);

/* The same op for the long sequences of the full-resolution blocks (webp_inference.py:295, 317-321 at H*W >= 1024;
 * head_dim 8 or 16, L % 128 == 0): qkv is BINARY16 (the in_proj GEMM writes it with epilogue out_dtype DDPMIR_F16, q rows
 * pre-scaled as above), out is bf16.  Three tiers per 128-row tile, chosen on the device from the tile's Cauchy-Schwarz
 * logit bound: <= 11 (exp2 domain) scores and probabilities stay in binary16 on the tensor cores (attn_tc16.cu); <= 60
 * bf16 probabilities (attn_tc.cu); beyond that the exact online-maximum kernel.  workspace:
 * ddpmir_attention_prescaled_f16_workspace(B, L, C, heads) bytes (flags + room for a bf16 copy of qkv that is written
 * only when a tile leaves the first tier).
 * Tier 0, ahead of those three and chosen per (image, head): when the logit bound max_i |q'_i| * max_j |k_j| is <= 2, the
 * softmax weights are evaluated with a minimax polynomial of the logit (degree 2-4, relative error <= 2.5e-3 per weight, the
 * same polynomial the quadratic tiers use on the FMA pipe) THROUGH ITS MONOMIAL FEATURE MAP, which turns the attention of
 * that (image, head) into two O(L) contractions instead of an O(L^2) one (attn_lin.cu). */
size_t ddpmir_attention_prescaled_f16_workspace(int B, int L, int C, int heads);
int ddpmir_attention_prescaled_f16(const void* qkv, int B, int L, int C, int heads, void* workspace, void* out,
                                   ddpmir_stream_t stream);

/* out = x + y * s[b, c] on fp32 NHWC tensors ([B, HW, C]); FrequencyAwareBlock.forward of the 0409 UNet
 * (experiments/code/0409_method.ipynb#c0:L256-263: x + x_freq * attn, attn a per-image per-channel gate).  out2 (optional)
 * receives a copy in out2_dtype for the GEMMs that read the result.  C % 8 == 0. */
int ddpmir_channel_scale_add(const float* x, const float* y, const float* s, int B, long long HW, int C, float* out,
                             void* out2, int out2_dtype, ddpmir_stream_t stream);

/* Bit-exact baseline-JPEG round trip on the device: the pixels Pillow returns for
 * `img.save(buf, "JPEG", quality=q, subsampling=4:4:4 if q > 30 else 4:2:0)` + `Image.open(buf)` -- jpeg_compress,
 * svd.ipynb#c1:L20-44 / 0409_method.ipynb#c0:L44-62 -- computed with libjpeg-turbo's integer arithmetic (colour
 * conversion, h2v2 downsampling, islow DCT, quantisation, islow IDCT, fancy upsampling), entropy coding skipped because
 * it is lossless.  rgb, out: uint8 [B, H, W, 3], any H and W (planes are padded to whole MCUs as libjpeg pads them).
 * workspace: ddpmir_jpeg_roundtrip_workspace(B, H, W) bytes. */
size_t ddpmir_jpeg_roundtrip_workspace(int B, int H, int W);
int ddpmir_jpeg_roundtrip_u8(const uint8_t* rgb, uint8_t* out, int B, int H, int W, int quality, int subsample_420,
                             void* workspace, ddpmir_stream_t stream);

/* DCT-domain JPEG projection: DCTProcessor.jpeg_compress, experiments/code/dct.ipynb#c2:L100-139, the
 * reference's pure-torch JPEG simulator: per channel and 8x8 block  c = DCT(x255 - 128);  c = round(c / Q) * Q  (luma
 * table for channel 0, chroma table otherwise, scaled by `quality` as at L105-112);  x255' = IDCT(c) + 128.  No colour
 * conversion and no clamping, as in the reference.  x255 = x * in_scale + in_offset and out = (x255' - in_offset) / in_scale:
 * (1, 0) projects 0..255 images like the reference, (127.5, 127.5) projects the sampler's [-1, 1] images.  fp32 NCHW,
 * H % 8 == 0, W % 8 == 0.  An opt-in replacement of the host codec round trip in the JPEG sampler -- NOT libjpeg. */
int ddpmir_jpeg_dct_project(const float* x, float* out, int B, int C, int H, int W, float quality, float in_scale,
                            float in_offset, ddpmir_stream_t stream);

/* Blockwise per-channel transform  out = alpha*x + beta * (T_c X T_c^T)  with zero padding to a multiple of bs
 * and crop back: DCTLayer.forward webp_inference.py:161-192 (per_channel = 0, T [bs,bs]) and
 * AVIFAdaptiveTransform avif_inference.py:140-177 (per_channel = 1, T [C,bs,bs]).  bs in {4, 8}. T fp32. */
int ddpmir_block_transform(const void* x, int dtype, int B, int H, int W, int C, const float* T, int bs,
                           int per_channel, float alpha, float beta, void* out, int out_dtype,
                           ddpmir_stream_t stream);

/* nn.MaxPool2d(2), webp_inference.py:342.  out [B, H/2, W/2, C]. */
int ddpmir_maxpool2(const void* x, int dtype, int B, int H, int W, int C, void* out, ddpmir_stream_t stream);

/* torch.cat([F.interpolate(lo, scale_factor=2, mode='bilinear', align_corners=False), skip], dim=1),
 * webp_inference.py:389-393.  lo [B,H,W,C1], skip [B,2H,2W,C2], out [B,2H,2W,C1+C2]. */
int ddpmir_upsample2_concat(const void* lo, const void* skip, int dtype, int B, int H, int W, int C1, int C2,
                            void* out, ddpmir_stream_t stream);

/* AdaptiveAvgPool2d(s) for s in {1,2,4,8}, avif_inference.py:195: out [85, B, C] fp32 (cell-major so that the
 * rows of one scale are contiguous for ddpmir_linear_rows), cells ordered s=1 (1), s=2 (4, row-major), s=4 (16),
 * s=8 (64).  Any H, W (PyTorch's adaptive window rule). */
int ddpmir_avgpool_pyramid(const void* x, int dtype, int B, int H, int W, int C, float* out,
                           ddpmir_stream_t stream);

/* enhanced + residual of AVIFFreqAwareBlock.forward, avif_inference.py:226-256:
 * out = h + xt * mean_s(bilinear_up(gates_s)) * color * edge, with gates [85,B,C] fp32 the sigmoid outputs of
 * the pooled MLPs (F.interpolate(..., mode='bilinear', align_corners=False) restated in-kernel) and color/edge
 * already multiplied by their boosts.  h is in h_dtype (the fp32 stream); xt, color, edge and out in `dtype`. */
int ddpmir_avif_combine(const void* h, int h_dtype, const void* xt, const float* gates, const void* color,
                        const void* edge, int dtype, int B, int H, int W, int C, void* out, ddpmir_stream_t stream);

/* out_conv tail, webp_inference.py:365-366: tanh(conv3x3(x) + bias), NHWC `dtype` in, NCHW fp32 out.
 * w is the checkpoint's [3, Cin, 3, 3] fp32. */
int ddpmir_out_conv_tanh(const void* x, int dtype, int B, int H, int W, int Cin, const float* w, const float* bias,
                         int N, float* out, ddpmir_stream_t stream);

/* ------------------------------------------------------------------------------------------------------ */
/* Training step (webp_training.py:476-537): backward kernels.  Gradients are fp32; `dtype` arguments name the   */
/* dtype of saved activations (operands may be bf16).  Data gradients of conv3x3 / gemm layers reuse the forward */
/* entry points with transposed, tap-flipped weights.                                                            */
/* ------------------------------------------------------------------------------------------------------ */

/* Weight gradient of conv3x3 (taps = 9) / 1x1 (taps = 1): out += dY^T * im2col(X) for the sub-block
 * [n_begin, +n_count) x [k_begin, +k_count) of the packed [N, taps*Cin] weight, leading dimension out_ld; oihw = 1
 * writes the checkpoint's OIHW layout of a 3x3 conv instead.  `out` is accumulated into (zero it for a fresh gradient). */
int ddpmir_wgrad(const void* dy, int dy_dtype, const void* x, int x_dtype, float* out, int B, int H, int W, int Cin, int N,
                 int taps, int n_begin, int n_count, int k_begin, int k_count, int out_ld, int oihw, ddpmir_stream_t stream);

/* Column sums of dY [B*H*W, N] accumulated into out_total[N] and/or out_img[B, N] (bias / time-embedding row-bias
 * gradients); cls = -1 all pixels, 1 / 0 only low- / high-frequency pixels (the class-dependent second-layer bias of the
 * stacked gate MLP).  Only columns [n_begin, n_begin + n_count) are summed; outputs are indexed from 0. */
int ddpmir_colsum(const void* dy, int dtype, int B, int H, int W, int N, int cls, int bs, int low, int n_begin, int n_count,
                  float* out_total, float* out_img, ddpmir_stream_t stream);

/* Backward of y = act(GroupNorm(x)) (act in NONE/GELU/SILU): dx (optionally accumulated), dgamma / dbeta accumulated.
 * x, dy, dx fp32 NHWC; ws: B*G*2 doubles. */
int ddpmir_groupnorm_backward(const float* x, const float* dy, int B, int HW, int C, int G, int act, const float* mean_rstd,
                              const float* gamma, const float* beta, float* dx, int accumulate_dx, float* dgamma, float* dbeta,
                              double* ws, ddpmir_stream_t stream);

/* Frequency gate e = h3 + g*s*d (g = sigmoid gate, s = 1 | boost[b], webp_inference.py:263-267): dz = de*s*d*g*(1-g),
 * dd = de*g*s.  g, d in op_dtype. */
int ddpmir_gate_backward(const float* de, const void* g, const void* d, int op_dtype, const float* boost, float* dz, float* dd,
                         int B, int H, int W, int C, int bs, int low, ddpmir_stream_t stream);

/* Hidden layer of the stacked gate MLP (freq_mode 1 epilogue): dpre = dg1 * class mask * LeakyReLU'(pre). */
int ddpmir_lrelu_mask_backward(const float* dg1, const void* g1, int op_dtype, float* dpre, int B, int H, int W, int N, int bs,
                               int low, ddpmir_stream_t stream);

/* ---- AVIF family: backward of AVIFFreqAwareBlock / AVIFAdaptiveTransform (avif.py:185-321), used by the training step
 * train_epoch_ddrm_avif (avif.py:528-590).  GEMM-shaped pieces reuse ddpmir_conv3x3 / ddpmir_gemm / ddpmir_wgrad. ---- */

/* Product rule of e = h + xt * A * color * edge (avif.py:318-321), A = 1/4 sum_s bilinear_up(gates_s) rebuilt from the
 * [85, B, C] gate pyramid exactly as ddpmir_avif_combine does.  color = boost_color[b] * sigmoid(zc), edge likewise
 * (avif.py:309-315).  Writes dxt = de*A*color*edge, the PRE-sigmoid gradients dz_color, dz_edge, and dattn = de*xt*color*edge/4
 * (gradient of the SUM of the up-sampled maps).  xt, color, edge in `dtype`; all outputs fp32 [B,H,W,C]. */
int ddpmir_avif_combine_backward(const float* de, const void* xt, const float* gates, const void* color, const void* edge,
                                 int dtype, const float* boost_color, const float* boost_edge, int B, int H, int W, int C,
                                 float* dxt, float* dz_color, float* dz_edge, float* dattn, ddpmir_stream_t stream);

/* Transposed F.interpolate(bilinear, align_corners=False) of the four gate maps (avif.py:297-300):
 * dgates[cell, b, c] = sum over pixels of weight(pixel, cell) * dattn[b, pixel, c];  dgates fp32 [85, B, C]. */
int ddpmir_avif_gates_backward(const float* dattn, int B, int H, int W, int C, float* dgates, ddpmir_stream_t stream);

/* Transposed nn.AdaptiveAvgPool2d(1|2|4|8) (avif.py:197): dx[b,h,w,c] (+)= sum over the cells containing (h,w) of
 * dpooled[cell,b,c] / |cell|.  dpooled fp32 [85, B, C]; dx fp32 [B,H,W,C], added to when accumulate != 0. */
int ddpmir_avgpool_pyramid_backward(const float* dpooled, int B, int H, int W, int C, float* dx, int accumulate,
                                    ddpmir_stream_t stream);

/* Weight gradient of the learned per-channel blockwise transform Z = T_c X T_c^T (AVIFAdaptiveTransform.transform_weights,
 * avif.py:191,216-223): dT[c] += sum over blocks of dZ^T (T X) + (dZ T) X^T.  x, dz fp32 [B,H,W,C] (zero padding / crop at ragged
 * edges as in the forward); T, dT fp32 [C, bs, bs]; bs = 8.  The data gradient is ddpmir_block_transform with T_c^T. */
int ddpmir_block_transform_wgrad(const float* x, const float* dz, const float* T, int bs, int B, int H, int W, int C, float* dT,
                                 ddpmir_stream_t stream);

/* Per-step operand packing of a checkpoint-layout weight w [N, Cin, kh, kw] (fp32, taps = kh*kw = 9 or 1) for the training step
 * (replaces the permute/flip/cat/cast chain over `state_dict` tensors; the weights change at every optimizer.step(),
 * webp_training.py:524): fwd[n*fwd_ld + fwd_off + tap*Cin + c] = w[n,c,tap] (the [N,(kh,kw,cin)] layout of ddpmir_conv3x3 /
 * ddpmir_gemm) and bwd[c*bwd_ld + bwd_off + (taps-1-tap)*N + n] = w[n,c,tap] (transposed, taps flipped: the data-gradient
 * operand).  Either output may be NULL; both are `dtype` (DDPMIR_BF16 or DDPMIR_F32). */
int ddpmir_pack_weight(const float* w, int N, int Cin, int taps, void* fwd, long long fwd_ld, long long fwd_off, void* bwd,
                       long long bwd_ld, long long bwd_off, int dtype, ddpmir_stream_t stream);

/* nn.ReLU backward from the saved OUTPUT y = relu(pre) (y in y_dtype): dpre = dy * [y > 0]. */
int ddpmir_relu_mask_backward(const float* dy, const void* y, int y_dtype, float* dpre, int64_t n, ddpmir_stream_t stream);

/* nn.Dropout(p) in train mode (webp_inference.py:291,313): out = keep ? in/(1-p) : 0 with a counter-hash mask keyed by
 * (seed, element); applying it to the incoming gradient with the same seed is its backward. */
int ddpmir_dropout(const void* in, int in_dtype, void* out, int out_dtype, int64_t n, float p, uint64_t seed,
                   ddpmir_stream_t stream);

int ddpmir_maxpool2_backward(const float* x, const float* dy, float* dx, int B, int H, int W, int C, ddpmir_stream_t stream);
int ddpmir_upsample2_concat_backward(const float* dy, float* dlo, float* dskip, int B, int H, int W, int C1, int C2,
                                     ddpmir_stream_t stream);

/* Attention for training: forward that also returns the row log-sum-exp lse [B, heads, L], and the backward
 * dqkv [B, L, 3C] fp32 from (qkv, o, dout, lse); delta is a [B, heads, L] fp32 workspace.  dout_op (optional): dout
 * cast to the operand dtype -- with bf16 operands, L % 64 == 0 and head_dim <= 64 the backward then runs on the tensor
 * cores (attn_bwd_mma.cu); otherwise the generic fp32-FMA kernels are used. */
int ddpmir_attention_train_forward(const void* qkv, int dtype, int B, int L, int C, int heads, void* out, float* lse,
                                   ddpmir_stream_t stream);
int ddpmir_attention_backward(const void* qkv, const void* o, int dtype, const float* dout, const void* dout_op,
                              const float* lse, float* delta, float* dqkv, int B, int L, int C, int heads, ddpmir_stream_t stream);

/* Small fp32 layers: element-wise activation forward / backward (u = pre-activation), row-wise linear backward
 * (dx may be NULL; dw, db accumulated), the 3-channel input convolution (weight gradient in OIHW + the gradients of the
 * folded GroupNorm affine; bias via ddpmir_colsum) and the out_conv + tanh tail. */
int ddpmir_act_forward(const float* x, int act, float* out, int64_t n, ddpmir_stream_t stream);
int ddpmir_act_backward(const float* dy, const float* u, int act, float* dx, int64_t n, ddpmir_stream_t stream);
int ddpmir_linear_rows_backward(const float* dy, const float* x, const float* w, int rows, int K, int N, float* dx,
                                int accumulate_dx, float* dw, float* db, ddpmir_stream_t stream);
int ddpmir_conv_input_backward(const float* x, const float* dh, int B, int Cin, int H, int W, int N, int ksize, const float* w,
                               const float* mean_rstd, const float* gamma, const float* beta, float* dw, float* dgamma,
                               float* dbeta, ddpmir_stream_t stream);
int ddpmir_out_conv_tanh_backward(const void* a, int dtype, const float* y, const float* dy, int B, int H, int W, int Cin, int N,
                                  const float* w, float* da, float* dw, float* dbias, ddpmir_stream_t stream);

/* Loss backward pieces of frequency_aware_loss (webp_training.py:105-132): da (+)= weight * d mse / da;
 * dpred += d/dpred [w_mag * sum (|P|-|T|)^2 + w_phase * sum (angle P - angle T)^2] (ws_*: [planes,H,W] complex64);
 * dx += weight * d ssim / dx (ws: planes*3*(H-10)*(W-10) floats). */
int ddpmir_mse_backward(const float* a, const float* b, int64_t n, float weight, float* da, int accumulate, ddpmir_stream_t stream);
int ddpmir_freq_loss_backward(const float* pred, const float* target, int planes, int H, int W, float w_mag, float w_phase,
                              float* ws_pred, float* ws_target, float* ws_grad, float* dpred, ddpmir_stream_t stream);
int ddpmir_ssim_backward(const float* x, const float* y, int planes, int H, int W, int clamp01, float weight, float* dx, float* ws,
                         ddpmir_stream_t stream);
/* Gradients of the two AVIF-specific terms (avif.py:126-164): the full-spectrum frequency terms (same contract as
 * ddpmir_freq_loss_backward) and dpred (+)= w_v d acc2[0]/dpred + w_h d acc2[1]/dpred of ddpmir_edge_loss. */
int ddpmir_fft2_loss_backward(const float* pred, const float* target, int planes, int H, int W, float w_mag, float w_phase,
                              float* ws_pred, float* ws_target, float* ws_grad, float* dpred, ddpmir_stream_t stream);
int ddpmir_edge_loss_backward(const float* pred, const float* target, int planes, int H, int W, float w_v, float w_h,
                              float* dpred, int accumulate, ddpmir_stream_t stream);

/* nn.HuberLoss(reduction='mean', delta) (0409_method.ipynb#c0:L438, 567) over n elements, and its gradient
 * da (+)= weight * d huber / da.  n % 4 == 0 for the forward; ws: one double. */
int ddpmir_huber(const float* a, const float* b, int64_t n, float delta, float* out_scalar, double* ws, ddpmir_stream_t stream);
int ddpmir_huber_backward(const float* a, const float* b, int64_t n, float delta, float weight, float* da, int accumulate,
                          ddpmir_stream_t stream);

/* Gradient of the colour term of color_preservation_loss (0409_method.ipynb#c0:L67-76; forward: ddpmir_color_l1):
 * dpred (+)= weight * d(0.25 L1_R + 0.5 L1_G + 0.25 L1_B)/dpred on NCHW fp32 [-1,1] images clamped to [0,1]. */
int ddpmir_color_l1_backward(const float* pred, const float* target, int B, int H, int W, float weight, float* dpred,
                             int accumulate, ddpmir_stream_t stream);

/* Optimiser (webp_training.py:521-524, 775): acc += sum x^2 (global gradient norm), and the fused
 * clip_grad_norm_(max_norm) + AdamW update of one tensor (grad_sumsq NULL = no clipping; step >= 1). */
int ddpmir_sumsq(const float* x, int64_t n, double* acc, ddpmir_stream_t stream);
int ddpmir_adamw_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                      float weight_decay, int step, const double* grad_sumsq, float max_norm, ddpmir_stream_t stream);

/* The same fused clip + AdamW update over many tensors in ONE launch.  The gradient and both moment estimates live in flat
 * buffers (all parameters' segments back to back); chunk c covers flat elements [chunk_start[c], chunk_start[c] + chunk_len[c])
 * of one tensor and chunk_param[c] points at that tensor's matching elements.  The three tables are device arrays; chunks of
 * parameters the caller wants untouched (torch skips .grad None) are simply left out. */
int ddpmir_adamw_multi(const int64_t* chunk_start, const int32_t* chunk_len, float* const* chunk_param, int n_chunks,
                       const float* g_flat, float* m_flat, float* v_flat, float lr, float beta1, float beta2, float eps,
                       float weight_decay, int step, const double* grad_sumsq, float max_norm, ddpmir_stream_t stream);

/* dtype conversion helper for weight pre-packing (fp32 -> bf16, round-to-nearest-even). */
int ddpmir_cast_f32_to_bf16(const float* in, void* out, int64_t n, ddpmir_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* DDPMIR_H */
