/* mmpl_b200.h -- C ABI of libmmpl_b200.so: the B200 (sm_100a) kernels behind the multimodal-PL dense hot path.
 *
 * Every entry point enqueues work on the caller's CUDA stream and returns immediately (no host synchronisation,
 * no device allocation).  All pointers are DEVICE pointers unless stated otherwise; the caller owns every buffer.
 * Return value: 0 = ok, negative = MMPL_E_*; the message is available from mmpl_last_error() (thread-local).
 * There is no CPU fallback: a device that is not compute capability 10.x yields MMPL_E_ARCH.
 *
 * Layout conventions
 *   activations : NDHWC ("channels last"), element type selected by `dtype` (MMPL_F32 | MMPL_BF16)
 *   image       : [N,1,D,H,W] fp32 (identical to NDHWC for one channel)
 *   logits      : [N,C,D,H,W] fp32, the layout the reference modules return (unet3D.py:713)
 *   weights     : master copy fp32 in the reference layout [Cout,Cin,kd,kh,kw] (unet3D.py:18); kernels consume
 *                 the standardised, tap-major packing produced by mmpl_ws_weight_fwd:
 *                   fprop packing  [tap][Cout][Cin]   (tap = (kd*k+kh)*k+kw)
 *                   dgrad packing  [tap'][Cin][Cout]  (tap' = flipped tap, so dgrad is a plain correlation)
 *   GN stats    : double [N][G][2] raw sums (sum x, sum x^2) per sample and group; consumers derive mean/rstd.
 *
 * Each function cites the reference code it replaces (path:line in TThuraya/multimodal-PL).
 */
#ifndef MMPL_B200_H_
#define MMPL_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* mmpl_stream_t; /* cudaStream_t */

enum { MMPL_F32 = 0, MMPL_BF16 = 1 };
enum { MMPL_ALGO_DIRECT = 0, MMPL_ALGO_TCGEN05 = 1, MMPL_ALGO_TCGEN05_PSPLIT = 2 };
enum {
  MMPL_OK = 0,
  MMPL_E_SHAPE = -1,
  MMPL_E_DTYPE = -2,
  MMPL_E_ALIGN = -3,
  MMPL_E_ARCH = -4,
  MMPL_E_CUDA = -5,
  MMPL_E_UNSUPPORTED = -6
};

int mmpl_version(void);
const char* mmpl_last_error(void);
/* 0 if the current device is sm_100-class, else MMPL_E_ARCH. */
int mmpl_check_device(void);
/* Number of kernel launches issued by this library in this process (for bench.py's gpu_launches). */
uint64_t mmpl_launch_count(void);

/* ---- weight standardisation: Conv3d.forward, unet3D.py:22-26 -------------------------------------------------
 * w [Cout][Cin][taps] fp32.  Outputs: w_hat fp32 (same layout), inv_std [Cout] fp32, and the two tap-major
 * packings in `dtype` (either may be NULL).  standardise=0 copies/packs w unchanged (plain nn.Conv3d). */
int mmpl_ws_weight_fwd(const float* w, int cout, int cin, int taps, int standardise, float* w_hat, float* inv_std,
                       void* packed_fprop, void* packed_dgrad, int dtype, mmpl_stream_t stream);
/* The same for every convolution of a network in ONE launch.  `table_dev` is a DEVICE array of `count` entries sorted
 * by first_block; a block serves 8 consecutive out-channels, entry i owns blocks [first_block, first_block +
 * ceil(cout/8)) of the grid, total_blocks = sum of ceil(cout/8).  16-byte aligned packings, cout % 8 == 0 for the
 * vectorised dgrad packing (any cout works, narrower stores).
 * stem_kch > 0 (Cin = 1 stem only): packed_fprop is [cout][stem_kch] with the 27 taps in columns 0..26 and, for
 * stem_kch = 64, again in 32..58 (the operand mmpl_stem_im2col pairs with); other columns are left untouched. */
typedef struct mmpl_ws_entry {
  const float* w;
  float* w_hat;
  float* inv_std;
  void* packed_fprop;
  void* packed_dgrad;
  int32_t cout, cin, taps, standardise, first_block, stem_kch;
} mmpl_ws_entry; /* 64 bytes */
int mmpl_ws_weight_fwd_batched(const mmpl_ws_entry* table_dev, int count, int total_blocks, int dtype,
                               mmpl_stream_t stream);
/* Backward of the standardisation: g_hat = dL/dw_hat given tap-major [tap][Cout][Cin] fp32 (as wgrad writes it);
 * dw (reference layout, fp32) = (g - mean(g) - w_hat * sum(g*w_hat)/(n-1)) * inv_std.  standardise=0: un-pack. */
int mmpl_ws_weight_bwd(const float* g_hat_tapmajor, const float* w_hat, const float* inv_std, int cout, int cin,
                       int taps, int standardise, float* dw, mmpl_stream_t stream);

/* ---- convolution: F.conv3d in Conv3d.forward, unet3D.py:27; k in {1,3}, pad = k/2, stride in {1,2} -----------
 * fprop : y[N,Do,Ho,Wo,Cout] = conv(x[N,D,H,W,Cin], w_fprop) (+ residual if non-NULL, same shape/dtype as y)
 * dgrad : dx[N,D,H,W,Cin]    = conv^T(dy[N,Do,Ho,Wo,Cout], w_dgrad) (+ addend if non-NULL)
 * wgrad : dw_tapmajor[tap][Cout][Cin] fp32 = sum_voxels dy * x_shifted   (overwritten, not accumulated)
 * algo MMPL_ALGO_TCGEN05 requires bf16 and channel counts of 32 or multiples of 64 (MMPL_E_UNSUPPORTED otherwise);
 * for stride-2 3x3x3 fprop/wgrad the activation argument is the parity-split copy written by mmpl_parity_split and
 * the algo is MMPL_ALGO_TCGEN05_PSPLIT.  MMPL_ALGO_DIRECT (CUDA cores, fp32 accumulate) handles every case. */
/* P[(p*N + n)][d'][h'][w'][c] = X[n][2d'+pd][2h'+ph][2w'+pw][c], p = pd*4+ph*2+pw, extents ceil(D/2) etc., zeros
 * where the source index is out of range.  bf16 only. */
int mmpl_parity_split(const void* x, void* p_out, int n, int d, int h, int w, int c, int dtype, mmpl_stream_t stream);
/* gn_stats_out (may be NULL): zero-initialised double [N][16][2]; receives the GroupNorm(16) raw sums of y -- fused into
 * the tcgen05 epilogue when one tile spans all output channels, otherwise computed by a following streaming pass. */
int mmpl_conv3d_fprop(const void* x, const void* w_fprop, const void* residual, void* y, int n, int d, int h, int w,
                      int cin, int cout, int ksize, int stride, int dtype, int algo, double* gn_stats_out,
                      mmpl_stream_t stream);
/* Optional fusion of the first pass of the GroupNorm+ReLU backward into the dgrad epilogue.  dx is the gradient
 * w.r.t. a = relu(gn(.)), this convolution's forward input (NoBottleneck.forward, unet3D.py:59-66).  Since
 * a = gamma*xhat + beta where the ReLU passes and 0 elsewhere, the sums the GroupNorm backward needs are
 *   S1_c = sum_v dx*[a > 0]      and      Q_c = gamma_c * sum_v g*xhat = sum_v dx*a - beta_c*S1_c,
 * which the epilogue accumulates while dx is still in registers (one extra row read of `a`).  `ws` is the
 * double [N][Cin][6] workspace of mmpl_gn_relu_bwd (zero on entry); head selects columns {0,1} or {2,3}.
 * *gn_fused_out = 1 if the request was honoured (tcgen05 algorithms), 0 if the caller must run the reduction pass. */
typedef struct mmpl_gn_bwd_fuse {
  const void* a;           /* forward input of this convolution, dtype of dx: [N,D,H,W,Cin], or its parity-split copy */
  const float* beta;       /* [Cin] GroupNorm bias of the node that produced a */
  double* ws;              /* [N][Cin][6], accumulated into */
  int a_is_parity_split;   /* 1: `a` is the mmpl_parity_split tensor (stride-2 3x3x3 convolutions keep only that) */
  int head;                /* 0, or 1 for the second head of a dual GroupNorm */
} mmpl_gn_bwd_fuse;
int mmpl_conv3d_dgrad(const void* dy, const void* w_dgrad, const void* addend, void* dx, int n, int d, int h, int w,
                      int cin, int cout, int ksize, int stride, int dtype, int algo, const mmpl_gn_bwd_fuse* gn,
                      int* gn_fused_out, mmpl_stream_t stream);
int mmpl_conv3d_wgrad(const void* x, const void* dy, float* dw_tapmajor, int n, int d, int h, int w, int cin,
                      int cout, int ksize, int stride, int dtype, int algo, void* workspace, size_t workspace_bytes,
                      mmpl_stream_t stream);
size_t mmpl_conv3d_wgrad_workspace(int n, int d, int h, int w, int cin, int cout, int ksize, int stride, int algo);

/* ---- stem (Cin = 1) and classifier (1x1x1 with bias, NCDHW fp32 logits): unet3D.py:594, :629-633 -------------- */
int mmpl_stem_conv_fwd(const float* image, const float* w_hat /*[Cout][27]*/, void* y, int n, int d, int h, int w,
                       int cout, int dtype, mmpl_stream_t stream);
/* ---- stem conv3x3x3(1 -> base) on tcgen05 with the 27-tap (hi/lo bf16) operand tile built in shared memory from an
 * fp32 image halo: no expanded image tensor in HBM (self.conv1, unet3D.py:594, :666).  image [N,1,D,H,W] fp32;
 * w_packed = the stem packing [cout][64] bf16 written by mmpl_ws_weight_fwd_batched (stem_kch = 64); y [N,D,H,W,cout]
 * bf16; gn_stats_out (may be NULL) zero-initialised double [N][16][2], receives the GroupNorm(16) raw sums of y.
 * wgrad: dy [N,D,H,W,cout] bf16 -> dw_tapmajor fp32 [27][cout] (zeroed by the call).  cout 32 or 64. */
int mmpl_stem_tc_fwd(const float* image, const void* w_packed, void* y, double* gn_stats_out, int n, int d, int h,
                     int w, int cout, mmpl_stream_t stream);
int mmpl_stem_tc_wgrad(const float* image, const void* dy, float* dw_tapmajor, int n, int d, int h, int w, int cout,
                       mmpl_stream_t stream);

/* bf16 path: x27 [N,D,H,W,channels] bf16 = the 27 shifted copies of the image; the stem is then a channels -> Cout
 * 1x1x1 convolution for mmpl_conv3d_fprop / mmpl_conv3d_wgrad (tcgen05).  channels = 32: taps in 0..26, zeros in
 * 27..31.  channels = 64: hi bf16 part in 0..26 and the lo part (x - hi) in 32..58, i.e. the image keeps 16 mantissa
 * bits and only the weights are rounded to bf16, as in every other layer. */
int mmpl_stem_im2col(const float* image, void* x27, int n, int d, int h, int w, int channels, mmpl_stream_t stream);
/* dw is tap-major [27][Cout] fp32; workspace (>= mmpl_stem_conv_wgrad_workspace bytes) holds a zero-padded image copy. */
size_t mmpl_stem_conv_wgrad_workspace(int n, int d, int h, int w);
int mmpl_stem_conv_wgrad(const float* image, const void* dy, float* dw_tapmajor /*[27][Cout]*/, int n, int d, int h,
                         int w, int cout, int dtype, void* workspace, size_t workspace_bytes, mmpl_stream_t stream);
int mmpl_cls_fwd(const void* a, const float* wc /*[C][Cin]*/, const float* bias, float* logits, int n, int64_t spatial,
                 int cin, int classes, int dtype, mmpl_stream_t stream);
/* gn_beta / gn_ws (both NULL or both set): `a` is the output of precls_conv.0/1 = GroupNorm+ReLU (unet3D.py:629-631);
 * the kernel then also accumulates that node's backward sums S1, Q into gn_ws[N][Cin][6] (see mmpl_gn_bwd_fuse). */
int mmpl_cls_bwd(const void* a, const float* wc, const float* dlogits, void* da, float* dwc, float* dbias,
                 const float* gn_beta, double* gn_ws, int n, int64_t spatial, int cin, int classes, int dtype,
                 mmpl_stream_t stream);
/* Training head fused: precls_conv.2 (unet3D.py:632) + EDiceLoss_partial.forward (loss_partial.py:71-99) without the
 * fp32 logits / dlogits tensors ever reaching memory.  a: [n][spatial][cin] bf16 (cin 32 or 64), classes <= 16; target,
 * class_weight, lut, per_sample, sums, loss, uce as in mmpl_partial_loss_fwd / _bwd.  The backward recomputes the logits,
 * writes da (bf16), dwc [classes][cin] and dbias [classes] and, with gn_beta / gn_ws, adds the first
 * pass of the GroupNorm+ReLU backward of `a` exactly like mmpl_cls_bwd. */
int mmpl_cls_loss_fwd(const void* a, const float* wc, const float* bias, const void* target, int target_is_u8,
                      const float* class_weight, const float* lut, int per_sample, double* sums, float* loss, int n,
                      int64_t spatial, int cin, int classes, int uce, mmpl_stream_t stream);
int mmpl_cls_loss_bwd(const void* a, const float* wc, const float* bias, const void* target, int target_is_u8,
                      const float* class_weight, const float* lut, int per_sample, const double* sums,
                      const float* grad_out, void* da, float* dwc, float* dbias, const float* gn_beta, double* gn_ws,
                      int n, int64_t spatial, int cin, int classes, int uce, mmpl_stream_t stream);

/* ---- GroupNorm(16)+ReLU: NoBottleneck.forward, unet3D.py:59-60,64-65 and downsample.0/1, :645-646 -------------
 * stats: double [N][G][2], must be zero before mmpl_gn_stats accumulates into it. */
int mmpl_gn_stats(const void* x, double* stats, int n, int64_t spatial, int c, int groups, int dtype,
                  mmpl_stream_t stream);
/* y = relu(gn(x; gamma,beta)); if gamma2 != NULL also y2 = relu(gn(x; gamma2,beta2)) from the same read of x
 * (gn1 and downsample.0 share their input, unet3D.py:59 and :69).
 * real_cpg (0 = all): how many of the C/groups channels of every group carry data.  Networks whose widths the tensor-core
 * kernels do not cover (the refiner unet3D_g: 24/48/96/192 channels in groups of 6/12/24/48, unet3D.py:1507-1559) run
 * zero-padded to 32/64/128/256 channels with every group padded in place; the zero channels add nothing to the raw sums
 * and real_cpg makes mean / variance (and the group means of the backward) divide by the real element count. */
/* layout (0 = plain NDHWC outputs; otherwise vol_d*vol_h*vol_w == spatial are the volume extents):
 *   bit 1 (2): the SECOND head is stored on the even voxels only, compact [N][ceil(D/2)][ceil(H/2)][ceil(W/2)][C] -- the
 *              1x1x1 stride-2 downsample (unet3D.py:645-651) reads nothing else, so 7/8 of that head's writes (and of its
 *              gradient's reads in the backward) never happen;
 *   bit 0 (1): the FIRST head is written in the parity-split layout P[pc*N + n][d/2][h/2][w/2][C], pc = (d&1)<<2 | (h&1)<<1 |
 *              (w&1), that the stride-2 3x3x3 tensor-core convolution reads (even extents only; forward only) -- what
 *              mmpl_parity_split would otherwise produce in an extra pass. */
int mmpl_gn_relu_fwd(const void* x, const double* stats, const float* gamma, const float* beta, void* y,
                     const float* gamma2, const float* beta2, void* y2, int n, int64_t spatial, int c, int groups,
                     int real_cpg, int vol_d, int vol_h, int vol_w, int layout, float eps, int dtype, mmpl_stream_t stream);
/* dx = d/dx of the one or two GN+ReLU heads (+ addend if non-NULL); dgamma/dbeta per head (fp32 [C]).
 * workspace: N*C*6 + 1 doubles: [N][C][6] = per head {S1 = sum g, Q = gamma * sum g*xhat} (g = dy*[relu gate]) in columns 0..3 and
 * scratch in 4..5.  reduced = 0: the call zeroes it and runs the reduction pass over (x, dy[, dy2]); reduced = 1: the
 * producers of dy (and dy2) already accumulated S1 and Q (mmpl_gn_bwd_fuse, mmpl_cls_bwd), only the apply pass runs.
 * In both cases the workspace is zero again when the call's work completes.  Channels with gamma == 0 (where Q
 * carries no information about sum g*xhat) get their dgamma from an exact accumulation inside the apply pass. */
int mmpl_gn_relu_bwd(const void* x, const double* stats, const float* gamma, const float* beta, const void* dy,
                     const float* gamma2, const float* beta2, const void* dy2, const void* addend, void* dx,
                     float* dgamma, float* dbeta, float* dgamma2, float* dbeta2, double* workspace, int reduced, int n,
                     int64_t spatial, int c, int groups, int real_cpg, int vol_d, int vol_h, int vol_w, int layout, float eps,
                     int dtype, mmpl_stream_t stream);

/* ---- class-token attention maps + token EMA of unet3D_with_feam3 (unet3D.py:142-212, :1051-1068, :1127-1175) --------
 * mmpl_ln_rows_*: LayerNorm over the channel axis of `rows` NDHWC voxel rows without affine (biased variance, eps under
 * the square root); fwd writes y = xhat and rstd [rows]; bwd: dx = rstd (g - mean(g) - xhat mean(g xhat)).  The attention
 * map itself is mmpl_cls_fwd on xhat with the folded 15 x C matrix (csrc/eam.cu explains the algebra).
 * mmpl_token_stats: sums [ntok][C], counts [ntok] (zeroed by the call) of the feature rows whose nearest-neighbour
 * down-sampled label (mask [N][dm][hm][wm], fp32 or uint8 class ids) is l + 1; mmpl_token_ema: token[l] <- (1-alpha)
 * token[l] + alpha sums[l]/counts[l] where counts[l] > 0 (renew_token without host synchronisation). */
int mmpl_ln_rows_fwd(const void* x, void* y, float* rstd, int64_t rows, int c, float eps, int dtype, mmpl_stream_t stream);
int mmpl_ln_rows_bwd(const void* xhat, const float* rstd, const void* g, void* dx, int64_t rows, int c, int dtype,
                     mmpl_stream_t stream);
int mmpl_token_stats(const void* x, const void* mask, int mask_is_u8, float* sums, float* counts, int n, int d, int h,
                     int w, int c, int dm, int hm, int wm, int ntok, int dtype, mmpl_stream_t stream);
int mmpl_token_ema(float* token, const float* sums, const float* counts, int ntok, int c, float alpha,
                   mmpl_stream_t stream);

/* ---- device-side input pipeline: AMOSDataSet_newatlas.__getitem__, MOTSDataset.py:299-395, and the
 * intensity augmentations of get_train_transform, :33-52.  Volumes are [h][w][d] (d fastest) as the reference holds them,
 * src_dtype 0 = fp32, 1 = int16, 2 = uint8; outputs are [crop_d][crop_h][crop_w].
 * mmpl_volume_moments: moments[2] = {sum, sum of squares} in fp64 (zeroed by the call) for the MRI z-score.
 * mmpl_prepare_patch: zero-pad to crop + 5 (:370-372), scale (mode 0: CT clip +-325 HU / 325; mode 1: MRI (x - mean) / std
 *   over the padded volume, `moments` from mmpl_volume_moments; mode 2: copy, for labels; :171-186), crop at (b, c, a)
 *   (:377-383), transpose (:389-391).  Integer sources are scaled in fp64 and rounded once, like numpy promotes them.
 * mmpl_atlas_patch: nearest-neighbour resize of atlas [k][ha][wa][da] to (h, w, d) (:357), then pad / crop / transpose.
 * mmpl_augment_patch (in place): x += N(0, noise_std) from a counter-based generator (seed), x = x * mult + add,
 *   x = clip((x - mean) * contrast + mean, min, max) with stats = {sum, min, max} from mmpl_patch_stats (double[4]);
 *   neutral parameters (0, 1, 0, 1) skip a step.  mmpl_blur_axis: one axis of scipy.ndimage.gaussian_filter ('reflect'). */
int mmpl_volume_moments(const void* volume, int src_dtype, int64_t count, double* moments, mmpl_stream_t stream);
int mmpl_prepare_patch(const void* volume, int src_dtype, void* out, int out_is_u8, int h, int w, int d, int b, int c, int a,
                       int crop_h, int crop_w, int crop_d, int mode, const double* moments, mmpl_stream_t stream);
int mmpl_atlas_patch(const float* atlas, float* out, int k, int ha, int wa, int da, int h, int w, int d, int b, int c, int a,
                     int crop_h, int crop_w, int crop_d, mmpl_stream_t stream);
int mmpl_patch_stats(const float* x, int64_t n, double* stats /*[4]*/, mmpl_stream_t stream);
int mmpl_augment_patch(float* x, int64_t n, float noise_std, uint64_t seed, float mult, float add, float contrast,
                       const double* stats, mmpl_stream_t stream);
int mmpl_blur_axis(const float* src, float* dst, int d, int h, int w, int axis, const float* taps_dev, int radius,
                   mmpl_stream_t stream);

/* ---- helpers for the refiner unet3D_g and the discriminator norm_style_discriminator_output (SURVEY 8f-f3;
 * unet3D.py:1507-1623, :1907-1947; csrc/aux_nets.cu explains how their convolutions map onto the tcgen05 kernels).
 * mmpl_space_to_depth2: [N,D,H,W,C] (bf16 or fp32) -> [N,D/2,H/2,W/2,cp], channel (pd*4+ph*2+pw)*C + c, zero beyond 8*C
 *   (inverse = 1: the transpose, x is the s2d tensor and y the full-resolution one; used as the backward).
 * mmpl_bias_lrelu_*: y = leaky_relu(x + bias[c], slope) over `rows` NDHWC rows; bwd reads the gate from y (slope > 0),
 *   writes dx and dbias [C] (zeroed by the call).
 * mmpl_upsample2x_ncdhw_*: nn.Upsample(scale_factor=2, mode='trilinear') (align_corners=False) of `planes` fp32
 *   [d][h][w] volumes (the refiner's final up-sampling of its logits, unet3D.py:1621) and its transpose. */
int mmpl_space_to_depth2(const void* x, void* y, int n, int d, int h, int w, int c, int cp, int inverse, int dtype,
                         mmpl_stream_t stream);
int mmpl_bias_lrelu_fwd(const void* x, const float* bias, void* y, int64_t rows, int c, float slope, int dtype,
                        mmpl_stream_t stream);
int mmpl_bias_lrelu_bwd(const void* y, const void* dy, void* dx, float* dbias, int64_t rows, int c, float slope, int dtype,
                        mmpl_stream_t stream);
int mmpl_upsample2x_ncdhw_fwd(const float* x, float* y, int64_t planes, int d, int h, int w, mmpl_stream_t stream);
int mmpl_upsample2x_ncdhw_bwd(const float* dy, float* dx, int64_t planes, int d, int h, int w, mmpl_stream_t stream);

/* ---- trilinear x2 upsample (align_corners=False) + skip add: unet3D.py:608, :686-687 -------------------------- */
/* gn_stats (optional, may be NULL): double [N][16][2], zeroed by the caller; receives the GroupNorm(16) raw sums
 * (sum, sum of squares per group) of y, i.e. the statistics of the next block's gn1 / downsample.0 (unet3D.py:59, :645). */
int mmpl_upsample2x_add_fwd(const void* x_lo, const void* skip, void* y, int n, int d, int h, int w, int c,
                            int dtype, void* gn_stats, mmpl_stream_t stream);
int mmpl_upsample2x_bwd(const void* dy, void* dx_lo, int n, int d, int h, int w, int c, int dtype,
                        mmpl_stream_t stream);

/* ---- partial-label loss: EDiceLoss_partial.forward + DiceLoss.forward, loss_partial.py:38-57, :71-99 ----------
 * logits [N,C,S] fp32; target [N,S] class ids, fp32 like the reference's label tensors (target_is_u8 = 0) or uint8
 * (target_is_u8 = 1; a quarter of the bytes over PCIe and HBM).
 * per_sample = 0 (the reference): Dice sums pooled over batch and voxels, class_weight [C] = mask[0] (:87, :92); lut
 *   (may be NULL) [C] is the cmask remap of train_amos_atlas_final.py:252-255 applied to the target on the fly.
 * per_sample = 1 (mixed CT/MRI batches): class_weight [N][C] and lut [N][C]; the reference formula is
 *   evaluated per sample with that sample's weights (= the reference called once per sample) and averaged over N.
 * sums: double [G][4][C] = I, Z, Y, E per group (G = N when per_sample else 1) followed by ONE more double the forward
 * uses as its last-block ticket, i.e. a workspace of G*4*C+1 doubles (zeroed by the call; per call, so launches on
 * different streams / devices never share state and nothing is allocated by the library); loss: one fp32.
 * classes <= 32.  16-byte loads (4 voxels per thread and class plane) when S % 4 == 0 and the planes are aligned. */
int mmpl_partial_loss_fwd(const float* logits, const void* target, int target_is_u8, const float* class_weight,
                          const float* lut, int per_sample, double* sums, float* loss, int n, int64_t spatial,
                          int classes, int uce, mmpl_stream_t stream);
int mmpl_partial_loss_bwd(const float* logits, const void* target, int target_is_u8, const float* class_weight,
                          const float* lut, int per_sample, const double* sums,
                          const float* grad_out /*device scalar*/, float* dlogits, int n, int64_t spatial, int classes,
                          int uce, mmpl_stream_t stream);

/* ---- binary Dice over a voxel gate (+ BCE-with-logits): DiceLoss._dice_loss / EDiceLoss_full2.forward,
 * loss_partial.py:24-36, :150-170 (the pseudo-label terms of get_loss, losses.py:165-176) --------------------------
 * x, target, gate: fp32 [voxels]; gate may be NULL (= all voxels) and is a 0/1 mask otherwise.  sigmoid = 1: the score
 * is sigmoid(x), else x itself.  uce = 1 adds mean BCE-with-logits over ALL voxels.  sums: double[5] = {I, Y, Z, E,
 * last-block ticket}, written by fwd and read by bwd.  dtarget may be NULL. */
int mmpl_masked_dice_fwd(const float* x, const float* target, const float* gate, double* sums, float* loss,
                         int64_t voxels, int sigmoid, int uce, mmpl_stream_t stream);
int mmpl_masked_dice_bwd(const float* x, const float* target, const float* gate, const double* sums,
                         const float* grad_out, float* dx, float* dtarget, int64_t voxels, int sigmoid, int uce,
                         mmpl_stream_t stream);

/* ---- SGD with momentum: torch.optim.SGD at train_amos_atlas_final.py:132-135,378 ------------------------------
 * d = grad*grad_scale + wd*p; buf = first ? d : mom*buf + d; p -= lr*buf.  lr is read from a device scalar so a
 * captured CUDA graph can be replayed while the poly schedule (utils.py:53-60) changes it. */
int mmpl_sgd_step(float* p, const float* grad, float* buf, int64_t count, const float* lr_dev, float momentum,
                  float weight_decay, float grad_scale, int first_step, mmpl_stream_t stream);

/* ---- sliding-window blend + argmax/Dice: predict_sliding evaluate_amos.py:261-279, get_dice :128-141 ----------
 * acc and wsum [D][H][W] in `acc_bytes` per element (4 = fp32, 8 = fp64 like the reference).  acc is class-major
 * [C][D][H][W] (d_outer = 0, the layout predict_sliding returns) or depth-major [D][C][H][W] (d_outer = 1: a depth slab is
 * one contiguous block, which the multi-GPU path exchanges plane-wise along D).  wsum may be NULL (not accumulated). */
int mmpl_sw_blend(void* acc, void* wsum, const float* tile_logits /*[C][td][th][tw]*/, const float* gauss, int c,
                  int d, int h, int w, int td, int th, int tw, int d0, int h0, int w0, int acc_bytes, int d_outer,
                  mmpl_stream_t stream);
/* Classifier + blend in one kernel (bf16 activations a [td*th*tw][cin], cin 32/64, fp32 acc): acc += gauss * (a W^T + b)
 * at the tile origin read from DEVICE memory (origin_dev = {d0, h0, w0}), so one captured CUDA graph serves every tile.
 * Replaces mmpl_cls_fwd + mmpl_sw_blend on the inference path (no fp32 logits tile in HBM). */
int mmpl_cls_blend(const void* a, const float* wc /*[C][cin]*/, const float* bias, const float* gauss, float* acc,
                   float* wsum /*may be NULL*/, const int* origin_dev, int classes, int d, int h, int w, int td, int th,
                   int tw, int cin, int d_outer, mmpl_stream_t stream);
/* normalise -> argmax -> Dice counts over `voxels` = Dl*plane voxels.  plane = 0: acc is class-major [C][voxels];
 * plane = H*W: acc is depth-major [Dl][C][plane] (any depth slab).  wsum may be NULL (acc already normalised; the fp32
 * fast path never divides -- the argmax does not depend on a positive normaliser).  label fp32 or uint8 class ids.
 * out_logits (may be NULL) [C][voxels] fp32 = acc/wsum; argmax uint8 [voxels]; counts int64 [3][C] = |P&T|,|P|,|T|
 * (zeroed by the call; pass NULL with label NULL to skip the metrics). */
int mmpl_sw_finalize(const void* acc, const void* wsum, const void* label, int label_is_u8, float* out_logits,
                     uint8_t* argmax, long long* counts, int c, int64_t voxels, int64_t plane, int acc_bytes,
                     mmpl_stream_t stream);
/* dst[i] += src[i], fp32: the slab owner of the sharded sliding window adds the partial accumulator planes
 * its peers send (evaluate._sliding_blend; the reference sums full_probs on one device, evaluate_amos.py:268-272). */
int mmpl_accumulate_f32(float* dst, const float* src, int64_t n, mmpl_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MMPL_B200_H_ */
