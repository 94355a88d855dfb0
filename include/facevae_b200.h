/* facevae_b200.h -- C ABI of libfacevae_b200.so (sm_100a only).
 *
 * The reference (Luh1124/face-vae) has no C/FFI boundary: its hot path is PyTorch nn.Modules whose arithmetic
 * dispatches into ATen/cuDNN.  This header is the boundary a maintainer binds instead: each
 * entry point names the reference call site (file:line in /root/reference) whose device work it replaces.  The
 * Python facade in face_vae_b200/ (same class names and signatures as the reference's modules.py / models.py /
 * losses.py) calls these through ctypes; INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless stated otherwise;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*; NULL = legacy default stream);
 *   - return 0 on success, non-zero on failure (bad shape / unsupported configuration / CUDA error); the message
 *     is available from fv_last_error() (thread-local).  Nothing is thrown, nothing falls back to another path;
 *   - the library never allocates, frees or retains caller memory;
 *   - sums over pixels (batch-norm statistics, bias / weight gradients, losses) are combined in a FIXED order, so every entry
 *     point is bitwise reproducible from run to run.  Entry points that reduce across thread blocks take `red_ws`: a scratch
 *     buffer of fv_reduce_ws_bytes() bytes whose first 512 bytes are zero before the first use (the kernels leave them zero
 *     again), private to one stream at a time; outputs of such reductions are WRITTEN, not accumulated;
 *   - activations are NHWC ("channels last") bf16 or fp32 with the channel count padded to a multiple of 16
 *     (Cp / Ci / Co_pad below); the NCHW fp32 tensors of the reference API are converted at the edges.
 */
#pragma once
#ifdef __cplusplus
extern "C" {
#endif

enum { FV_DT_BF16 = 0, FV_DT_F32 = 1 };
enum { FV_OUT_NHWC_BF16 = 0, FV_OUT_NHWC_F32 = 1, FV_OUT_NCHW_F32 = 2 };
enum { FV_MODE_NONE = 0, FV_MODE_POOL = 1, FV_MODE_UP = 2 };   /* fused 2x2 avg-pool / nearest 2x up-sample */
enum { FV_ACT_NONE = 0, FV_ACT_RELU = 1, FV_ACT_LEAKY = 2 };   /* LeakyReLU slope 0.2 (reference modules.py:29) */

#define FV_ABI_VERSION 2   /* bumped with every signature change; fv_abi_version() returns the value the library was built with */
const char* fv_last_error(void);
const char* fv_version(void);
int fv_abi_version(void);
int fv_device_ok(void);   /* 0 iff the current device is sm_10x */

/* ---- layout at the edges of the path -------------------------------------------------------------------- */
/* NCHW fp32 [N,C,H,W] -> NHWC [N,H,W,Cp] (bf16 or fp32), channels C..Cp-1 zero.  Frames enter the reference as
 * NCHW fp32 in [0,1] (dataset.py:116-129, logger.py:144-148). */
int fv_nchw_to_nhwc(const float* src, void* dst, int dst_dtype, int N, int C, int H, int W, int Cp, void* stream);
/* NHWC (channel stride Cs) -> NCHW fp32, first C channels; accumulate != 0 adds into dst. */
int fv_nhwc_to_nchw(const void* src, int src_dtype, float* dst, int N, int C, int H, int W, int Cs, int accumulate, void* stream);

/* F.interpolate(x, mode="bilinear", scale_factor=s, align_corners=False, recompute_scale_factor=True) of NCHW fp32 frames
 * (the input pre-scale of EFE_conv5 / EFE_conv6, models.py:764, 872): out [N,C,Ho,Wo], Ho = floor(H*s) chosen by the caller. */
int fv_bilinear_resize(const float* x, float* out, int N, int C, int H, int W, int Ho, int Wo, void* stream);

/* ---- convolution: nn.Conv2d inside _ConvBlock (modules.py:15,32), mid_conv (models.py:750,1096), out_conv
 *      (models.py:1099) and their autograd (aten::convolution_backward) ----------------------------------- */
/* nn.Conv2d weight [Co,Ci,R,S] fp32 -> wf bf16 [Co_pad][R*S][Ci_pad] (forward operand) and
 * wd bf16 [Ci_pad][R*S][Co_pad], taps rotated 180 degrees (data-gradient operand).  Either may be NULL. */
int fv_weight_prep(const float* w, void* wf, void* wd, int Co, int Ci, int R, int S, int Co_pad, int Ci_pad, void* stream);
/* The same for a table of layers in ONE launch (device array of n_layers descriptors; max_items = the largest
 * Co_pad*R*S*Ci_pad in the table): the per-step filter preparation of a whole network. */
typedef struct {
    const float* w;
    void* wf;                 /* kind 0: wf;  kind 1 (fv_weight_prep_up): wx2;  kind 2 (fv_weight_prep_s2): wf */
    void* wd;                 /* kind 0: wd;  kind 1: ws2;                      kind 2: wx2 */
    int Co, Ci, R, S, Co_pad, Ci_pad;
    int kind, reserved;
} fv_prep_desc;
int fv_weight_prep_batched(const fv_prep_desc* table_dev, int n_layers, long long max_items, void* stream);
/* The same with a flat grid sized by the layers: desc.reserved = index of the layer's first block, a block covering
 * fv_weight_prep_block_items() consecutive items of the layer's Co_pad * taps * Ci_pad (taps = R*S for kind 0, 16 otherwise);
 * total_blocks = sum over the layers.  (The 2-D grid of fv_weight_prep_batched gives every layer the same block count although
 * the layers differ 1000x in size: 58 us for the anchor's 13 convolutions.) */
int fv_weight_prep_block_items(void);
int fv_weight_prep_flat(const fv_prep_desc* table_dev, int n_layers, int total_blocks, void* stream);
/* The same through shared-memory tiles (coalesced filter reads, both operands written in their storage order): desc.reserved = running
 * sum of fv_weight_prep_tiled_blocks(kind, Co_pad, Ci_pad, R, S) over the preceding layers, total_blocks = the sum over all. */
int fv_weight_prep_tiled_blocks(int kind, int Co_pad, int Ci_pad, int R, int S);
int fv_weight_prep_tiled(const fv_prep_desc* table_dev, int n_layers, int total_blocks, void* stream);
/* y = conv(x, wf) + bias (+ residual); stride 1, odd square filter, pad = (R-1)/2.  tcgen05 implicit GEMM.
 * x: NHWC bf16 [N,H,W,Ci] (Ci = 16, 32 or a multiple of 64); wf: [Co_pad][R*S][Ci]; bias: fp32 [Co] or NULL;
 * residual: NHWC bf16 [N,H,W,Co_pad] or NULL (the `x +` of ResBlock2D, modules.py:124-125);
 * y: NHWC bf16 / NHWC fp32 with Co_pad channels, or NCHW fp32 [N,Co,H,W], per out_mode.
 * The data gradient is the same call with x := dY, wf := wd, (Ci, Co, Co_pad) := (Co_pad, Ci, Ci_pad). */
int fv_conv2d(const void* x, const void* wf, const float* bias, const void* residual, void* y, int out_mode, int N, int H, int W,
              int Ci, int Co, int Co_pad, int R, int S, int pad, void* stream);
/* bytes of the reduction scratch every `red_ws` parameter below expects (see Conventions) */
long long fv_reduce_ws_bytes(void);
/* fv_conv2d with the statistic pass of the following batch norm fused into the epilogue (replaces a separate fv_bn_stats
 * read of y): stats[0..Co_pad) = sum_pixels y, stats[Co_pad..2*Co_pad) = sum_pixels y^2 over the values as stored
 * (fp32; NHWC output modes only). */
int fv_conv2d_stats(const void* x, const void* wf, const float* bias, const void* residual, void* y, int out_mode, int N, int H,
                    int W, int Ci, int Co, int Co_pad, int R, int S, int pad, float* stats, void* red_ws, void* stream);
/* 1 when fv_conv2d_stats fuses the statistics into the epilogue for this shape, 0 when it runs a separate fv_bn_stats pass
 * over y (fusing costs the epilogue ~250 cycles per 16 channels and tile; it is done only where the tile's MMAs hide it). */
int fv_conv2d_fuses_stats(int out_mode, int N, int H, int W, int Ci, int Co_pad, int R, int S, int has_residual);
/* the same question for fv_conv2d_x2 (kind 1) / fv_conv2d_s2 (kind 2); H, W = the coarse grid */
int fv_conv2d_geom_fuses_stats(int kind, int out_mode, int N, int H, int W, int Ci, int Co_pad);
/* Weight gradient, tcgen05, split over pixels: part[split][Co_pad][R*S][Ci] (fp32) = this split's share of
 * sum_pixels x[pixel + tap] * dy[pixel]; splits = fv_conv2d_wgrad_splits(0, ...) slabs are written (plain stores). */
int fv_conv2d_wgrad_splits(int kind /* 0 same, 1 x2, 2 s2 */, int N, int H, int W, int Ci, int Co_pad, int R, int S);
int fv_conv2d_wgrad(const void* x, const void* dy, float* part, int splits, int N, int H, int W, int Ci, int Co_pad, int R, int S,
                    int pad, void* stream);
/* slabs added in split order -> nn.Conv2d layout grad [Co,Ci,R,S] fp32 (accumulate != 0 adds, as autograd does into .grad). */
int fv_wgrad_finish(const float* part, int splits, float* grad, int Co, int Ci, int R, int S, int Co_pad, int Ci_pad, int accumulate,
                    void* stream);
/* sums[C] = per-channel sums over P rows of an NHWC bf16 tensor: the bias gradient. */
int fv_colsum(const void* y, float* sums, long long P, int C, void* red_ws, void* stream);
/* out[n] (+)= sum over `slabs` partial vectors, slab_stride elements apart, in slab order. */
int fv_slab_sum(const float* part, int slabs, long long slab_stride, float* out, long long n, int accumulate, void* stream);

/* ---- UpBlock2D's nn.Upsample(x2, nearest) + 3x3 conv (modules.py:78-89) WITHOUT the up-sampled tensor, and the 4x4 stride-2
 *      convolution of Conv2dELR / EFE_conv6.efe_encoder (models_utils.py:632-744, models.py:845-852).  Output pixel
 *      (2i + a, 2j + b) of the up-sampled conv sees only the coarse pixels (i + u - 1 + a, j + v - 1 + b), u, v in {0, 1}: four
 *      2x2 "phase" convolutions on the coarse grid with summed taps (2.25x fewer MACs, 4x smaller input).  Its data gradient
 *      is a 4x4 stride-2 convolution of dY (the 2x2 sum of the up-sampling backward folded in), and vice versa. ----------- */
/* w [Co,Ci,3,3] fp32 -> wx2 bf16 [4][Co_pad][4][Ci_pad] (fv_conv2d_x2) and ws2 bf16 [Ci_pad][16][Co_pad] (fv_conv2d_s2 computing
 * the data gradient).  Either may be NULL. */
int fv_weight_prep_up(const float* w, void* wx2, void* ws2, int Co, int Ci, int Co_pad, int Ci_pad, void* stream);
/* w [Co,Ci,4,4] fp32 -> wf bf16 [Co_pad][16][Ci_pad] (fv_conv2d_s2) and wx2 bf16 [4][Ci_pad][4][Co_pad] (fv_conv2d_x2 computing
 * the data gradient).  Either may be NULL. */
int fv_weight_prep_s2(const float* w, void* wf, void* wx2, int Co, int Ci, int Co_pad, int Ci_pad, void* stream);
/* x NHWC bf16 [N,H,W,Ci] -> y [N,2H,2W,Co_pad] (out_mode as fv_conv2d) = bias + four 2x2 phase convolutions with
 * wp = [4][Co_pad][4*Ci].  stats (optional, NHWC outputs): [2][Co_pad] sum / sum of squares of y, needs red_ws. */
int fv_conv2d_x2(const void* x, const void* wp, const float* bias, void* y, int out_mode, int N, int H, int W, int Ci, int Co,
                 int Co_pad, float* stats, void* red_ws, void* stream);
/* x NHWC bf16 [N,2H,2W,Ci] -> y [N,H,W,Co_pad] = bias + 4x4 stride-2 pad-1 convolution with w = [Co_pad][16*Ci]. */
int fv_conv2d_s2(const void* x, const void* w, const float* bias, void* y, int out_mode, int N, int H, int W, int Ci, int Co,
                 int Co_pad, float* stats, void* red_ws, void* stream);
/* The general convolution entry: geometry kind (0 same, 1 x2, 2 s2; R, S, pad are used by kind 0 only), an optional activation
 * applied in the epilogue after the bias (FV_ACT_*: conv -> bias -> LeakyReLU(0.2) is Conv2dELR.forward, models_utils.py:712-742),
 * optional residual (kind 0) and fused statistics. */
int fv_conv2d_ex(int kind, const void* x, const void* w, const float* bias, const void* residual, void* y, int out_mode, int N, int H,
                 int W, int Ci, int Co, int Co_pad, int R, int S, int pad, int act, float* stats, void* red_ws, void* stream);
/* Conv2dELR weight path (models_utils.py:686-704): weff[co] = gain * w[co] / max(||w[co]||, 1e-12) (demod != 0: F.normalize over
 * dims 1..3) or gain * w; K = Ci*R*S elements per output channel; inv_norm[Co] (may be NULL) is kept for fv_demod_bwd, which maps
 * the gradient with respect to weff back to the gradient with respect to w. */
int fv_demod_fwd(const float* w, float* weff, float* inv_norm, int Co, int K, float gain, int demod, void* stream);
int fv_demod_bwd(const float* w, const float* inv_norm, const float* dweff, float* dw, int Co, int K, float gain, int demod, void* stream);
/* dy = g * act'(out) for an activation applied in a conv epilogue (NHWC bf16, n % 8 == 0 elements). */
int fv_act_bwd(const void* out, const void* g, void* dy, long long n, int act, void* stream);
/* x [N,H,W,Ci] coarse, dy [N,2H,2W,Co_pad] fine -> part[split][4][Co_pad][4][Ci]; splits = fv_conv2d_wgrad_splits(1, ...). */
int fv_conv2d_wgrad_x2(const void* x, const void* dy, float* part, int splits, int N, int H, int W, int Ci, int Co_pad, void* stream);
/* x [N,2H,2W,Ci] fine, dy [N,H,W,Co_pad] coarse -> part[split][Co_pad][16][Ci]; splits = fv_conv2d_wgrad_splits(2, ...);
 * fv_wgrad_finish(R = S = 4) completes it. */
int fv_conv2d_wgrad_s2(const void* x, const void* dy, float* part, int splits, int N, int H, int W, int Ci, int Co_pad, void* stream);
/* phase slabs of fv_conv2d_wgrad_x2 -> grad [Co,Ci,3,3] fp32 of the 3x3 filter. */
int fv_wgrad_finish_up(const float* part, int splits, float* grad, int Co, int Ci, int Co_pad, int Ci_pad, int accumulate, void* stream);

/* ---- out_conv: nn.Conv2d(32, 3, 7, padding 3) (models.py:1099) -> torch.sigmoid (models.py:1110) -> ReconLoss
 *      (losses.py:396-403, trainer.py:314), tap-folded tcgen05 schedule (csrc/fv_outconv.cu).  Shapes: 7x7 filter,
 *      Ci = 32, Co <= 4, W = 128 or 256 (fv_outconv_supported() != 0); other shapes use fv_conv2d. ----------------- */
int fv_outconv_supported(int N, int H, int W, int Ci, int Co, int R, int S);
/* weight [Co,Ci,7,7] fp32 -> wq bf16 [7][32][32] (forward operand: rows (s,co), K = ci) and wdq bf16 [7][32][32]
 * (data-gradient operand: rows ci, K = (s',co), taps rotated).  Either may be NULL. */
int fv_outconv_prep(const float* w, void* wq, void* wdq, int Co, int Ci, void* stream);
/* x: NHWC bf16 [N,H,W,32].  logits (NCHW fp32 [N,Co,H,W], optional when target is given) = conv(x) + bias.  With
 * target (NCHW fp32) the epilogue also produces pred = sigmoid(logits) (optional), loss_sum[1] = sum l(pred - target),
 * g4 = bf16 [N,H,W,4]: gscale * dloss/dlogits (channels Co..3 zero) and gsum[Co] = its per-channel sums (the bias gradient);
 * the fused loss needs red_ws. */
int fv_outconv_fwd(const void* x, const void* wq, const float* bias, float* logits, const float* target, float* pred, void* g4,
                   float* loss_sum, float* gsum, int N, int H, int W, int Ci, int Co, int l1, int use_sigmoid, float gscale,
                   void* red_ws, void* stream);
/* dx NHWC bf16 [N,H,W,32] = (*scale_ptr) * conv_transpose(g4, w) (scale_ptr may be NULL); g4 as written by fv_outconv_fwd. */
int fv_outconv_dgrad(const void* g4, const void* wdq, const float* scale_ptr, void* dx, int N, int H, int W, int Ci, int Co,
                     void* stream);
/* part[split][Co,32,7,7] fp32 (nn.Conv2d layout per slab) = this CTA's share of (*scale_ptr) * sum_pixels x[pixel + tap] *
 * g4[pixel]; splits = fv_outconv_wgrad_splits(N, H, W); fv_slab_sum adds the slabs. */
int fv_outconv_wgrad_splits(int N, int H, int W);
int fv_outconv_wgrad(const void* x, const void* g4, const float* scale_ptr, float* part, int splits, int N, int H, int W, int Ci,
                     int Co, void* stream);

/* ---- nn.SyncBatchNorm (modules.py:19) + ReLU/LeakyReLU (modules.py:27,29) + AvgPool2d (modules.py:62,70) /
 *      nn.Upsample (modules.py:81,89) ---------------------------------------------------------------------- */
/* sums[0..C) = sum y, sums[C..2C) = sum y^2 over P = N*H*W rows (all-reduced across ranks by the host before
 * fv_bn_finalize -- the stat exchange of torch/nn/modules/_functions.py:39-83). */
int fv_bn_stats(const void* y, int dtype, float* sums, long long P, int C, void* red_ws, void* stream);
/* stat[4][C] = mean, invstd (biased variance, eps), scale = gamma*invstd, shift = beta - mean*scale; updates
 * running_mean / running_var (unbiased variance, momentum) when given. */
int fv_bn_finalize(const float* sums, double count, const float* gamma, const float* beta, float* running_mean,
                   float* running_var, float momentum, float eps, float* stat, int C, void* stream);
/* eval mode: the same stat block from the running statistics. */
int fv_bn_eval_affine(const float* gamma, const float* beta, const float* running_mean, const float* running_var, float eps,
                      float* stat, int C, void* stream);
/* out = [pool2x2 | up2x]( act(scale*y + shift) ).  H, W are the input sizes.  out: NHWC (bf16/fp32) or NCHW fp32. */
int fv_bn_act_fwd(const void* y, int in_dtype, const float* stat, void* out, int out_dtype, int nchw_out, int N, int H, int W,
                  int C, int mode, int act, void* stream);
/* fv_bn_finalize + fv_bn_act_fwd in one launch (single-process training): every block derives scale / shift from the sums,
 * block 0 writes stat_out[4][C] (kept for backward) and updates the running statistics. */
int fv_bn_act_fwd_fin(const void* y, int in_dtype, const float* sums, double count, const float* gamma, const float* beta,
                      float* running_mean, float* running_var, float momentum, float eps, float* stat_out, void* out, int out_dtype,
                      int nchw_out, int N, int H, int W, int C, int mode, int act, void* stream);
/* backward pass 1: sums[0..C) = sum dz, sums[C..2C) = sum dz*xhat (all-reduced across ranks like
 * torch/nn/modules/_functions.py:144-159).  g is the gradient of the block output (pooled / up-sampled domain). */
int fv_bn_act_bwd_reduce(const void* y, int y_dtype, const void* g, int g_dtype, int g_nchw, const float* stat, float* sums,
                         int N, int H, int W, int C, int mode, int act, void* red_ws, void* stream);
/* dgamma/dbeta (+)= local sums; coef[2][C] = global sums / count. */
int fv_bn_bwd_finalize(const float* sums_local, const float* sums_global, double count, float* dgamma, float* dbeta, float* coef,
                       int C, int accumulate, void* stream);
/* backward pass 2: dy = scale*(dz - coef0 - xhat*coef1) (+ add), bf16 NHWC: the conv-output gradient. */
int fv_bn_act_bwd_apply(const void* y, int y_dtype, const void* g, int g_dtype, int g_nchw, const float* stat, const float* coef,
                        const void* add, void* dy, int N, int H, int W, int C, int mode, int act, void* stream);

/* fv_bn_bwd_finalize + fv_bn_act_bwd_apply in one launch (single-process training): coef = sums / count is derived in the
 * kernel, block 0 writes dgamma = sums[C..2C), dbeta = sums[0..C) (either may be NULL). */
int fv_bn_act_bwd_apply_fin(const void* y, int y_dtype, const void* g, int g_dtype, int g_nchw, const float* stat, const float* sums,
                            double count, float* dgamma, float* dbeta, const void* add, void* dy, int N, int H, int W, int C, int mode,
                            int act, void* stream);

/* ---- nn.InstanceNorm2d(C, affine=True) + ReLU / LeakyReLU (modules.py:21,27,29: the Discriminator's blocks, models.py:1120-1127) on
 *      NHWC bf16: statistics per (image, channel) over H*W, biased variance, no running statistics.
 *      stat [N][2][C] = mean | invstd;  sums [N][2][C] = sum dz | sum dz*xhat per image (dgamma / dbeta are their sums over N). */
int fv_in_stats(const void* y, float* stat, int N, int H, int W, int C, float eps, void* stream);
int fv_in_act_fwd(const void* y, const float* stat, const float* gamma, const float* beta, void* out, int N, int H, int W, int C, int act,
                  void* stream);
int fv_in_bwd_sums(const void* y, const void* g, const float* stat, const float* gamma, const float* beta, float* sums, int N, int H, int W,
                   int C, int act, void* stream);
int fv_in_bwd_apply(const void* y, const void* g, const float* stat, const float* sums, const float* gamma, const float* beta, void* dy,
                    int N, int H, int W, int C, int act, void* stream);

/* ---- first encoder layer: SameBlock2D(C <= 4 -> 32) on raw NCHW fp32 frames (modules.py:97-108 via models.py:749) ----
 * 1x1 conv + training-mode batch norm + ReLU is a per-pixel affine map whose statistics follow from the input moments.
 * sums are double: forward [C + C*C] = sum x_c | sum x_c x_d; backward [Co + Co*C] = sum dz | sum dz x_c (written;
 * all-reduced across ranks by the host).  coef [Co][C+1] = A | c with a = act(A x + c); stat [2][Co] = mean_y | invstd. */
int fv_pw_moments(const float* x_nchw, double* sums, int N, int C, int HW, void* red_ws, void* stream);
int fv_pw_prepare(const double* sums, double count, const float* w, const float* bias, const float* gamma, const float* beta,
                  float* running_mean, float* running_var, float momentum, float eps, float* coef, float* stat, int Co, int C, void* stream);
int fv_pw_fwd(const float* x_nchw, const float* coef, void* out_nhwc_bf16, int N, int C, int HW, int Co, int act, void* stream);
int fv_pw_bwd_reduce(const float* x_nchw, const void* g_nhwc_bf16, const float* coef, double* sums, int N, int C, int HW, int Co, int act,
                     void* red_ws, void* stream);
int fv_pw_bwd_finalize(const double* fsums, const double* bsums, double count, const float* w, const float* bias, const float* gamma,
                       const float* stat, float* dw, float* dgamma, float* dbeta, int Co, int C, void* stream);

/* ---- cross-rank statistic exchange over NVLink peer memory, fused with the finalize kernels ------------------------
 * Replaces the per-layer all_gather / all_reduce of SyncBatchNorm under DDP (torch/nn/modules/_functions.py:74-83,159;
 * reference modules.py:19, logger.py:55).  peer_bufs_dev: device array of `world` pointers to the ranks' symmetric
 * buffers (fv_xrank_buffer_floats() fp32 each, zero-initialised, mapped in every rank); epoch_ctr: device uint64,
 * zero-initialised, private to the rank.  mode 0: out = stat[4][C] (+ running stats), mode 1: out = coef[2][C] and
 * dgamma/dbeta from the LOCAL sums.  Every rank must issue the same sequence of calls. */
long long fv_xrank_buffer_floats(void);
int fv_bn_finalize_xrank(const float* sums_local, void* peer_bufs_dev, int rank, int world, void* epoch_ctr, int mode, double count,
                         const float* gamma, const float* beta, float* running_mean, float* running_var, float momentum, float eps,
                         float* out, float* dgamma, float* dbeta, int accumulate, int C, void* stream);
/* fv_bn_stats / fv_bn_act_bwd_reduce with the exchange and the finalize step run by the LAST block of the reduction itself (one
 * launch instead of two per batch-norm layer and pass on the data-parallel critical path).  `sums` still receives the local sums;
 * forward: stat[4][C] (+ running statistics); backward: coef[2][C], dgamma / dbeta (from the local sums, may be NULL). */
int fv_bn_stats_xrank(const void* y, int dtype, float* sums, long long P, int C, void* red_ws, void* peer_bufs_dev, int rank, int world,
                      void* epoch_ctr, double count, const float* gamma, const float* beta, float* running_mean, float* running_var,
                      float momentum, float eps, float* stat, void* stream);
int fv_bn_act_bwd_reduce_xrank(const void* y, int y_dtype, const void* g, int g_dtype, int g_nchw, const float* stat, float* sums, int N, int H,
                               int W, int C, int mode, int act, void* red_ws, void* peer_bufs_dev, int rank, int world, void* epoch_ctr,
                               double count, float* coef, float* dgamma, float* dbeta, void* stream);
/* The same exchange with all `world` ranks emulated as the blocks of ONE cooperative launch on one GPU (block r = rank r):
 * every per-rank array is the concatenation of the ranks' arrays; peer_bufs_dev points at `world` local buffers.  For the
 * single-GPU parity test of the protocol (waiting kernels must be co-resident, which separate launches do not guarantee).
 * A peer that never arrives is reported after FACEVAE_XRANK_TIMEOUT_S seconds of wall clock (default 600). */
int fv_bn_finalize_xrank_emulate(const float* sums_local, void* peer_bufs_dev, int world, void* epoch_ctr, int mode, double count,
                                 const float* gamma, const float* beta, float* running_mean, float* running_var, float momentum,
                                 float eps, float* out, float* dgamma, float* dbeta, int accumulate, int C, void* stream);

/* ---- gradient averaging of the data-parallel step (DistributedDataParallel's all-reduce, logger.py:55) over peer memory ----
 * In-place mean all-reduce of the ranks' flat gradient buffers, ONE kernel: rank r pulls slice r from all ranks over NVLink, adds
 * the `world` values in rank order (bitwise identical on every rank), multiplies by scale (1 / world) and pushes the result into
 * all buffers; epoch-flag barriers before the pull and after the push.  bufs_dev / flags_dev: device arrays of `world`
 * peer-mapped pointers -- buffers of n floats (n % 4 == 0, 16-byte aligned) and flag arrays of fv_grad_allreduce_flag_words()
 * zero-initialised 64-bit words; epoch_ctr (u64) and ticket (u32): zero-initialised device words owned by this rank.  world = 1
 * degenerates to the scaling.  A peer that never arrives traps after FACEVAE_XRANK_TIMEOUT_S seconds. */
long long fv_grad_allreduce_flag_words(void);
int fv_grad_allreduce(void* bufs_dev, void* flags_dev, int rank, int world, long long n, void* epoch_ctr, void* ticket, float scale,
                      void* stream);

/* ---- re-parameterisation (models.py:559-561) fused with KLDivergenceLoss (losses.py:385-393) ------------- */
/* mu/logstd: fp32 rows of Dz values (Dz % 4 == 0, 16-byte aligned), row_stride apart; z[N,Dz] = mu + exp(logstd)*eps (NULL
 * eps => z = mu; NULL z => KL only); kl_part[N][P] (may be NULL), P = fv_reparam_kl_parts(N, Dz): per-block partial sums of
 * sum_d(-0.5 - logstd + 0.5 mu^2 + 0.5 exp(2 logstd)) -- the caller adds the P partials of a row (no atomics). */
int fv_reparam_kl_parts(int N, int Dz);
int fv_reparam_kl_fwd(const float* mu, const float* logstd, long long row_stride, const float* eps, float* z, float* kl_part,
                      int N, int Dz, void* stream);
/* dmu = dz + k*mu (+dmu_ext); dlogstd = dz*eps*exp(logstd) + k*(exp(2 logstd)-1) (+dls_ext); k = kscale * (*kscale_ptr). */
int fv_reparam_kl_bwd(const float* mu, const float* logstd, long long row_stride, const float* eps, const float* dz,
                      const float* dmu_ext, const float* dls_ext, float kscale, const float* kscale_ptr, float* dmu, float* dls,
                      long long out_stride, int N, int Dz, void* stream);

/* ---- ReconLoss / nn.MSELoss (losses.py:396-403), nn.L1Loss (losses.py:128), torch.sigmoid (models.py:1110) - */
/* NCHW fp32 logits/target [N,C,H,W]; loss_sum[1] = sum l(pred - target); optional outputs: pred
 * (= sigmoid(logits) when use_sigmoid), gradient w.r.t. logits times gscale as fp32 NCHW and/or bf16 NHWC [N,H,W,Cp]. */
int fv_recon_loss(const float* logits, const float* target, float* pred_out, float* grad_f32, void* grad_nhwc, float* loss_sum,
                  int N, int C, int H, int W, int Cp, int l1, int use_sigmoid, float gscale, void* red_ws, void* stream);
/* same-shape flat fp32 tensors a, b of E elements: loss_sum[1] = sum l(a-b); grad (optional) = gscale * dl/da. */
int fv_recon_loss_flat(const float* a, const float* b, float* grad, float* loss_sum, long long E, int l1, float gscale, void* red_ws,
                       void* stream);
/* out = in * scale * (*scale_ptr) (scale_ptr may be NULL); n elements, n % 8 == 0. */
int fv_scale(const void* in, void* out, int dtype, long long n, const float* scale_ptr, float scale, void* stream);

/* ---- optimiser step of Logger.step (logger.py:60,160: Adam lr 5e-5, betas (0.5, 0.999)) ------------------------------------
 * torch.optim.Adam semantics (no amsgrad, no weight decay) over a device table of n_tensors descriptors in ONE launch;
 * *step_dev = the step count after this update (fp32, device memory: graph-capturable); max_n = the largest element count. */
typedef struct {
    float* p;
    const float* g;
    float* m;
    float* v;
    long long n;
} fv_adam_desc;
int fv_adam_multi(const fv_adam_desc* table_dev, int n_tensors, long long max_n, float lr, double beta1, double beta2, float eps,
                  const float* step_dev, void* stream);

/* ---- calibration (not on the product path) ---------------------------------------------------------------- */
/* cycles for `iters` back-to-back tcgen05.mma (M=128, K=16, N=n_cols) on shared-memory-resident operands. */
int fv_debug_mma_rate(int n_cols, int row_bytes, int iters, int a_distinct, int mn_major, int all_sms, long long* out_cycles_dev,
                      void* stream);

/* role-loop cycle counters of the ring kernels: registers a device buffer of 148*8 int64 counters (libraries built with
 * -DFV_TRACE only; NULL switches tracing off). */
int fv_debug_trace_set(long long* dev_counters);

#ifdef __cplusplus
}
#endif
