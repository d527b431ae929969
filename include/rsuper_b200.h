/* rsuper_b200.h — C-ABI of librsuper_b200.so (hand-written sm_100a kernels for the R-Super
 * 3D segmentation train step).
 *
 * The reference (MrGiovanni/R-Super @ 7f811ce1) is pure Python and has NO native interface of its
 * own: every entry point below replaces an ATen / cuDNN *library call* that
 * the reference makes, and each declaration cites that call site (file:line under
 * /root/reference/rsuper_train).  The reference-side binding a maintainer would add is a ctypes
 * stub; it is shown in INTEGRATION.md and implemented in r-super_b200/rsuper_b200/_lib.py.
 *
 * Conventions (all entry points)
 *   - plain pointers + sizes; no torch types; every pointer is a DEVICE pointer unless noted.
 *   - the caller (PyTorch caching allocator) owns all memory; kernels never allocate, free, or
 *     keep pointers after return.
 *   - `stream` is a cudaStream_t passed as void*; launches are asynchronous on it, there is no
 *     internal synchronisation and no default-stream use (=> CUDA-graph capturable).
 *   - return 0 on success, negative on error; rsb_last_error() gives the message (thread-local).
 *   - activations are channels-last NDHWC with an explicit channel pitch (elements between
 *     consecutive voxels), so a kernel can read/write a channel slice of a wider (concatenated)
 *     buffer: torch.cat at unet_utils.py:71 costs nothing.
 *   - `dtype`: RSB_BF16 (0) = bf16 storage, RSB_F32 (1) = fp32 storage.  Tensor-core operands are
 *     always bf16 with fp32 accumulation in TMEM.
 *   - InstanceNorm statistics travel as raw per-(n,c) pairs (sum, sumsq) in fp32.  A statistics
 *     array ALWAYS shares the channel pitch of the activation buffer it describes:
 *     stats[(n*pitch + c)*2 + {0,1}], so the statistics of a channel slice are the same slice of
 *     the statistics array (pointer offset by 2*c0 floats).  Producers accumulate them in their
 *     epilogue, consumers derive mean / rstd with eps = 1e-4 (conv_layers.py:39-42).
 */
#ifndef RSUPER_B200_H_
#define RSUPER_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RSB_BF16 0
#define RSB_F32 1

/* ---------------------------------------------------------------------------------------------
 * library
 * ------------------------------------------------------------------------------------------ */
const char* rsb_version(void);
const char* rsb_last_error(void);
/* number of SMs of the current device (grid sizing for the persistent kernels) */
int rsb_num_sms(void);

/* ---------------------------------------------------------------------------------------------
 * 3x3x3 convolution, stride 1, pad 1, no bias — nn.Conv3d inside ConvNormAct
 * (model/dim3/conv_layers.py:29-38), with the residual add of BasicBlock (conv_layers.py:92) and the
 * next layer's InstanceNorm statistics fused into the epilogue.  The operand `a` is the bf16 tensor
 * a = act(instnorm(x)) produced by rsb_norm_act (the pre-activation of conv_layers.py:47-49).
 * TMA-fed implicit GEMM on tcgen05 (M = 128 voxels, N = merged (kd, Cout tile), K = 9*Cin per plane),
 * accumulators in TMEM.  The same kernel is the data gradient (dgrad) when given flipped/transposed
 * weights; its epilogue then applies act'(xhat) and reduces the InstanceNorm-backward sums.
 * ------------------------------------------------------------------------------------------ */

/* N tile the kernels use for a given Cout (packing and launch agree on it) */
int rsb_conv3_n_tile(int Cout);
/* bytes of the packed bf16 weight image for a (Cout, Cin) 3x3x3 conv; parts = 1, 3 for the split-precision
 * image [hi | hi | lo], or 6 for the three-piece image [hi | lo | hi | lo2 | hi | lo] */
size_t rsb_conv3_packed_weight_bytes(int Cout, int Cin, int parts);

/* fp32 OIDHW [Cout][Cin][3][3][3] (the nn.Parameter layout) -> packed bf16 UMMA tiles.
 * transpose_flip = 0: forward operand.  1: dgrad operand (taps flipped, Cin/Cout swapped: the
 * packed image then describes a conv with Cin' = Cout, Cout' = Cin). */
int rsb_conv3_pack_weights(const float* w_oidhw, void* packed, int Cout, int Cin,
                           int transpose_flip, int parts, void* stream);

/* Batched packing: one launch for every conv of a network (replaces the per-layer rsb_conv3_pack_weights calls and
 * the torch.cat of the merged conv1 || shortcut weights — conv_layers.py:85-92).  A job's logical OIDHW tensor is
 * [w_a (rows_a rows) ; w_b (Cout - rows_a rows)].  Fill {w_a, w_b, packed, rows_a, Cout, Cin, transpose_flip, parts}
 * on the host, call rsb_conv3_pack_plan (fills the derived fields and the block count), copy the table to device
 * memory once, then launch rsb_conv3_pack_weights_batched on it every step. */
typedef struct RsbPackJob {
  const float* w_a;
  const float* w_b;
  void* packed;
  int rows_a, Cout, Cin;
  int transpose_flip, parts;
  /* derived (rsb_conv3_pack_plan) */
  int co_eff, ci_eff, NT, ntiles, nchunks;
  unsigned int block_begin;
  unsigned long long total;
  /* 1 => w_a / w_b are 1x1x1 kernels [rows][Cin]: embedded at the centre tap of the packed 3x3x3 image, other taps zero
   * (consumed with RsbConv3Args.pointwise = 1, or as an ordinary 3x3x3 image) */
  int pointwise;
} RsbPackJob;
int rsb_conv3_pack_plan(RsbPackJob* jobs_host, int n_jobs, unsigned int* total_blocks);
int rsb_conv3_pack_weights_batched(const RsbPackJob* jobs_device, int n_jobs, unsigned int total_blocks, void* stream);

typedef struct RsbConv3Args {
  /* geometry */
  int N, D, H, W;
  int Cin, Cout;
  int dtype; /* storage dtype of y, res, mask_x */
  /* conv operand, bf16 NDHWC (see rsb_norm_act); a_lo != NULL selects the 3-pass split-precision
   * product a_hi*w_hi + a_lo*w_hi + a_hi*w_lo (weights packed with parts = 3) */
  const void* a;
  int a_pitch;
  const void* a_lo;
  const void* a_lo2; /* third piece: 6-pass product carrying ~24 mantissa bits of both operands (weights packed with parts = 6) */
  /* packed weights from rsb_conv3_pack_weights */
  const void* w_packed;
  /* output */
  void* y;
  int y_pitch;
  /* epilogue (all optional) */
  const void* res; /* y = conv + res  (BasicBlock residual) */
  int res_pitch;
  float* out_stats; /* [N][y_pitch][2] += (sum, sumsq) of y  (fp32, pre-rounding) */
  /* dgrad epilogue: y = conv * act'(xhat), xhat from (mask_x, mask_stats);
   * bwd_sums[(n*mask_x_pitch+c)*2+{0,1}] += (sum y, sum y*xhat) — InstanceNorm backward
   * reductions; mask_stats uses the same indexing */
  const void* mask_x;
  int mask_x_pitch;
  const float* mask_stats;
  float* bwd_sums;
  float eps;   /* 1e-4 */
  float slope; /* 0 => ReLU (reference default), 0.01 => LeakyReLU */
  /* tuning (0 = auto) */
  int planes_per_item; /* PZ in {1,2,4} */
  int max_ctas;        /* 0 => number of SMs */
  /* 1 => w_packed is a 1x1x1 convolution ([Cout][Cin][1][1][1]: Bottleneck conv1 / conv3, conv_layers.py:104-108) embedded at the
   * centre tap of a 3x3x3 image whose other 26 taps are zero: only the centre tap is streamed and multiplied (same result,
   * 1/27 of the tensor work) */
  int pointwise;
} RsbConv3Args;

int rsb_conv3_forward(const RsbConv3Args* args, void* stream);
/* profiling aid: per-CTA pipeline wait-cycle counters of subsequent rsb_conv3_forward launches
 * (16 int64 per CTA; NULL disables).  Not used by the product path. */
int rsb_debug_set_timing_buffer(void* device_ptr);

/* weight gradient of the same conv: dW[co][ci][tap] = sum_v dy[v][co] * a[v+tap-1][ci], both
 * operands bf16 NDHWC (a = the rsb_norm_act operand tensor of the forward conv), fetched by TMA and
 * consumed MN-major by tcgen05 (K = voxels).  Writes fp32 OIDHW (overwrites or accumulates). */
typedef struct RsbConv3WgradArgs {
  int N, D, H, W;
  int Cin, Cout;
  const void* a;      /* bf16 [N][D][H][W][a_pitch] */
  int a_pitch;
  const void* dy;     /* bf16 [N][D][H][W][dy_pitch] */
  int dy_pitch;
  float* dw_oidhw;    /* [Cout][Cin][27] fp32 */
  int accumulate;     /* 0: overwrite, 1: += */
  void* workspace;    /* rsb_conv3_wgrad_workspace_bytes() */
  size_t workspace_bytes;
  int max_ctas;
} RsbConv3WgradArgs;

size_t rsb_conv3_wgrad_workspace_bytes(int Cout, int Cin, int max_ctas);
int rsb_conv3_wgrad(const RsbConv3WgradArgs* args, void* stream);
int rsb_debug_set_wgrad_timing_buffer(void* device_ptr); /* profiling aid, see above */

/* ---------------------------------------------------------------------------------------------
 * Conv operand producer: hi = bf16(act((x - mean) * rstd)) — the pre-activation
 * nn.InstanceNorm3d(eps=1e-4, affine=False) + nn.ReLU of ConvNormAct (conv_layers.py:39-49;
 * slope = 0 => ReLU, 0.01 => LeakyReLU) as one fused, 128-bit vectorised pass.  stats == NULL =>
 * plain cast.  lo != NULL additionally writes lo = bf16(value - hi) for the split-precision
 * parity mode; lo2 != NULL (needs lo) a third piece lo2 = bf16(value - hi - lo), together ~24 mantissa bits.
 * x has storage dtype `dtype`; hi / lo / lo2 are always bf16.  full != NULL additionally writes the activation
 * itself in the storage dtype — the block OUTPUT of a post-activation SingleConv = ConvNormAct(preact=False)
 * (conv_layers.py:50-68); hi may then be NULL.
 * ------------------------------------------------------------------------------------------ */
int rsb_norm_act(const void* x, int x_pitch, int dtype, const float* stats, float eps, float slope,
                 void* hi, int hi_pitch, void* lo, int lo_pitch, void* lo2, int lo2_pitch, void* full, int full_pitch,
                 int N, int D, int H, int W, int C, void* stream);

/* Backward of the post-activation a = act(instnorm(y)) when d(a) comes from pooling / upsampling / the head:
 * g = d * act'(yhat); bwd_sums[(n, c)] += (sum g, sum g * yhat).  rsb_instnorm_backward_apply(g, y, ...) then yields d(y)
 * (autograd of nn.InstanceNorm3d + nn.ReLU in ConvNormAct(preact=False), conv_layers.py:50-52). */
int rsb_act_backward_stats(const void* d, int d_pitch, const void* y, int y_pitch, const float* y_stats,
                           float* bwd_sums, void* g, int g_pitch, int dtype, float eps, float slope, int N, int D,
                           int H, int W, int C, void* stream);

/* ---------------------------------------------------------------------------------------------
 * stem conv 3x3x3 with Cin = 1 (inconv.conv1, model/dim3/unet_utils.py:15,18) — direct
 * CUDA-core kernel (K = 27 is too small for the tensor pipe); fp32 NCDHW(C=1) in, NDHWC out,
 * InstanceNorm statistics of the output accumulated in the epilogue.
 * ------------------------------------------------------------------------------------------ */
int rsb_stem_conv_forward(const float* x, const float* w_oidhw, void* y, int y_pitch, int dtype,
                          float* out_stats, int N, int D, int H, int W, int Cout, void* stream);
/* dW[co][tap] = sum_v dy[v][co] * x[v+tap-1] */
int rsb_stem_conv_wgrad(const float* x, const void* dy, int dy_pitch, int dtype, float* dw_oidhw,
                        int N, int D, int H, int W, int Cout, void* stream);

/* ---------------------------------------------------------------------------------------------
 * 1x1x1 head with bias (UNet.outc, model/dim3/unet.py:47,62): NDHWC raw features -> fp32 NCDHW
 * logits, and its backward (d features, dW, db).
 * ------------------------------------------------------------------------------------------ */
int rsb_head_forward(const void* x, int x_pitch, int dtype, const float* w /*[C][Cin]*/,
                     const float* bias, float* logits_ncdhw, int N, int D, int H, int W, int Cin,
                     int C, void* stream);
int rsb_head_backward(const void* x, int x_pitch, int dtype, const float* w,
                      const float* dlogits_ncdhw, void* dx, int dx_pitch, float* dw, float* db,
                      int N, int D, int H, int W, int Cin, int C, void* stream);

/* ---------------------------------------------------------------------------------------------
 * nn.MaxPool3d(2) (unet_utils.py:36) on NDHWC, output statistics fused; backward routes the
 * gradient to the first maximum in ATen scan order (d, h, w) and adds an optional second
 * gradient stream (the decoder skip) in the same pass.
 * ------------------------------------------------------------------------------------------ */
int rsb_maxpool2_forward(const void* x, int x_pitch, void* y, int y_pitch, int dtype,
                         float* out_stats, int N, int D, int H, int W, int C, void* stream);
int rsb_maxpool2_backward(const void* x, int x_pitch, const void* dy, int dy_pitch,
                          const void* dskip, int dskip_pitch, void* dx, int dx_pitch, int dtype,
                          int N, int D, int H, int W, int C, void* stream);

/* ConvTranspose3d(C_in, C, kernel_size = 2, stride = 2) up-sampling (the transposed-conv variant of up_block; semantics of
 * model/dim3/vnet.py:108) = a 1x1x1 conv to 8 * C channels at the INPUT resolution (rsb_conv3_forward, pointwise = 1, weight
 * rows (4a + 2b + c) * C + co = w[ci][co][a][b][c]) followed by this rearrangement:
 *   up[n, 2z+a, 2y+b, 2x+c, co] = q[n, z, y, x, (4a+2b+c) * C + co] + bias[co]   (+ InstanceNorm statistics of up)
 * D, H, W = the INPUT (low) resolution; rsb_space_to_depth2 is the inverse copy (the adjoint, used on the gradient). */
int rsb_depth_to_space2(const void* q, int q_pitch, const float* bias, void* up, int up_pitch, int dtype, float* out_stats,
                        int N, int D, int H, int W, int C, void* stream);
int rsb_space_to_depth2(const void* up, int up_pitch, void* q, int q_pitch, int dtype, int N, int D, int H, int W, int C,
                        void* stream);

/* ---------------------------------------------------------------------------------------------
 * F.interpolate(mode='trilinear', align_corners=True) (unet_utils.py:69), NDHWC, output
 * statistics fused, and its adjoint (deterministic gather).
 * ------------------------------------------------------------------------------------------ */
int rsb_upsample_trilinear_forward(const void* x, int x_pitch, void* y, int y_pitch, int dtype,
                                   float* out_stats, int N, int Di, int Hi, int Wi, int Do, int Ho,
                                   int Wo, int C, void* stream);
/* workspace: N*Di*(Ho + Hi)*Wo*C elements of the storage dtype for the separable adjoint (z pass -> [N][Di][Ho][Wo][C],
 * y pass -> [N][Di][Hi][Wo][C], x pass -> dx), or NULL for the single-pass 3-D gather. */
int rsb_upsample_trilinear_backward(const void* dy, int dy_pitch, void* dx, int dx_pitch,
                                    int dtype, int N, int Di, int Hi, int Wi, int Do, int Ho,
                                    int Wo, int C, void* workspace, void* stream);

/* ---------------------------------------------------------------------------------------------
 * InstanceNorm backward (affine=False), applied after the dgrad epilogue produced
 * g = dL/da * act'(xhat) and the reductions S1 = sum g, S2 = sum g*xhat:
 *   dx = rstd * (g - S1/V - xhat * S2/V)  [+ add]
 * (autograd of F.instance_norm as called by conv_layers.py:39-49).
 * ------------------------------------------------------------------------------------------ */
int rsb_instnorm_backward_apply(const void* g, int g_pitch, const void* x, int x_pitch,
                                const float* x_stats, const float* bwd_sums, const void* add,
                                int add_pitch, void* dx, int dx_pitch, int dtype, float eps,
                                int N, int D, int H, int W, int C, void* stream);

/* layout helpers: fp32 NCDHW <-> NDHWC(dtype, pitch) */
int rsb_ncdhw_to_ndhwc(const float* src, void* dst, int dst_pitch, int dtype, int N, int C,
                       int D, int H, int W, void* stream);
int rsb_ndhwc_to_ncdhw(const void* src, int src_pitch, int dtype, float* dst, int N, int C,
                       int D, int H, int W, void* stream);
/* per-(n,c) (sum, sumsq) of an NDHWC tensor (used by tests and as a fallback producer) */
int rsb_channel_stats(const void* x, int x_pitch, int dtype, float* stats, int N, int D, int H,
                      int W, int C, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Segmentation loss: mean(BCEWithLogits(r, l) * known) + DiceLossMultiClass(r, l, known)
 * (training/losses_foundation.py:945-956, 541-607), fp32 NCDHW logits.
 *   pass 1: one read of (logits, label, known) -> per-(b,c) partial sums {bce, TP, FP, FN}
 *   finalize: alpha_c (batch-coupled, clamped [0.2,0.8], carries gradient), loss scalars
 *   pass 2: dlogits = dloss * d(loss)/d(logits)   (includes d/d alpha terms)
 * label / known are uint8 (0/1).  class_weights may be NULL ([B][C] otherwise).
 * ------------------------------------------------------------------------------------------ */
typedef struct RsbSegLossArgs {
  int B, C;
  long long V; /* D*H*W */
  const float* logits;
  const uint8_t* label;
  const uint8_t* known; /* NULL => all ones */
  const float* class_weights;
  const float* bce_weight_map; /* optional fp32 [B][C][V] per-voxel weight of the BCE term (Ball loss: GWRP foreground
                                  weights + background indicator, losses_foundation.py:1775-1811); NULL => 1 */
  float* partials;  /* [B*C*4] workspace, zeroed by forward */
  float* coef;      /* [B*C*4] workspace (per-(b,c) backward coefficients) */
  float* loss_out;  /* [3]: total, bce, dice */
} RsbSegLossArgs;
int rsb_seg_loss_forward(const RsbSegLossArgs* args, void* stream);
/* dlogits (+)= grad_scale[0] * d(bce term)/dlogits + grad_scale[1] * d(dice term)/dlogits ; grad_scale points to
 * two device floats (pass the same value twice for the plain segmentation loss) */
int rsb_seg_loss_backward(const RsbSegLossArgs* args, const float* grad_scale, float* dlogits,
                          int accumulate, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Binary ball dilation — dilate_volume / dilate_volume_conv / create_ball_kernel
 * (losses_foundation.py:22-99, 1161-1232): iterated passes with the exact ball structuring
 * elements the reference builds; uint8 0/1 volumes [n_vol][D][H][W].
 * ------------------------------------------------------------------------------------------ */
int rsb_dilate_ball(const uint8_t* src, uint8_t* dst, uint8_t* tmp, int n_vol, int D, int H, int W,
                    int kernel_size, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Report-supervised losses (training/losses_foundation.py): Volume loss (:250-395) and Ball loss
 * (:1336-1864).  Rows are contiguous runs of V = D*H*W elements.  The scalar glue (per-tumour loop,
 * formulas on [B][L] values) is host code, as in the reference.
 * ------------------------------------------------------------------------------------------ */
/* get_lesion_channels (:204-248) for single-channel groups: dst[r] = src[row_map[r]]; and its adjoint */
int rsb_rows_gather(const void* src, const int* row_map, void* dst, int n_rows, long long V, int elem_bytes, void* stream);
int rsb_rows_scatter_add(const float* src, const int* row_map, float* dst, int n_rows, long long V, void* stream);
/* get_lesion_channels with groups that merge several channels of one organ (:215-220, torch.stack(...).max(dim=0)):
 * group_map[r * group_size + k] = source row of member k of output row r, -1 = none (member 0 always exists);
 * dst[r] = max over the members (uint8 masks: OR).  rows_scatter_add_max is its backward for the logits: the gradient of
 * row r is added to the member row that attained the maximum (the first one on ties), x = the tensor that was gathered. */
int rsb_rows_gather_max(const void* src, const int* group_map, int group_size, void* dst, int n_rows, long long V, int elem_bytes,
                        void* stream);
int rsb_rows_scatter_add_max(const float* grad_rows, const float* x, const int* group_map, int group_size, float* dst, int n_rows,
                             long long V, void* stream);
/* mask algebra on uint8 0/1 volumes; op: 0 a|b, 1 a&b, 2 a&~b, 3 ~(a|b), 4 ~a   (to_penalize :1605, borders :1721-1737) */
int rsb_u8_binary(const uint8_t* a, const uint8_t* b, uint8_t* out, int op, long long n, void* stream);
/* counts[r] = number of non-zero voxels of row r (the `.sum() > 0` / `.sum() < t` tests) */
int rsb_u8_row_count(const uint8_t* a, long long* counts, int n_rows, long long V, void* stream);
/* Volume loss reduction (:330-340): sums[r] = scale[r] * sum_v sigmoid(x[r][v]) * mask[r][v], and
 * dx[r][v] (+)= coef[r] * scale[r] * mask * sigmoid'(x) */
int rsb_masked_sigmoid_sum(const float* x, const uint8_t* mask, const float* scale, float* sums, int n_rows, long long V, void* stream);
int rsb_masked_sigmoid_grad(const float* x, const uint8_t* mask, const float* scale, const float* coef, float* dx, int accumulate,
                            int n_rows, long long V, void* stream);
/* Ball loss (:1680-1719): x_iter = sigmoid(x) * seg ; x_iter *= (1 - mask) */
int rsb_ball_prepare(const float* x, const uint8_t* seg, float* x_iter, long long V, void* stream);
int rsb_ball_remove(float* x_iter, const uint8_t* mask, long long V, void* stream);
/* isolate_tumor (:1387-1420): argmax of the cross-correlation of x_iter with a (Gaussian) ball given as n_taps int4
 * {dz, dy, dx, float bits of the weight}; argmax_out receives (score bits << 32 | 0xFFFFFFFF - flat index), first
 * maximum in flattened order */
size_t rsb_ball_workspace_bytes(int D, int H, int W);
int rsb_ball_correlate_argmax(const float* x_iter, const void* taps, int n_taps, int kernel_half, void* workspace,
                              long long* argmax_out, int D, int H, int W, void* stream);
/* The same argmax in three separable stages (rows -> discs -> planes; csrc/report_loss.cu): gauss = device float[R + 1],
 * g(d) = exp(-d^2 / 2 sigma^2); wtab = device int[(R + 1)^2], wtab[a * (R + 1) + b] = largest dx with a^2 + b^2 + dx^2 <= r^2
 * or -1 (the ball of create_ball_kernel as x ranges); R = floor(radius).  The normalisation of the reference's kernel is a
 * positive factor and is dropped.  gauss_host / wtab_host: the same two tables in HOST memory (optional; with them and
 * R <= 16 the disc stage runs from shared-memory tiles with the tables passed by value).
 * workspace: rsb_ball_sep_workspace_bytes(D, H, W, R) bytes. */
size_t rsb_ball_sep_workspace_bytes(int D, int H, int W, int R);
int rsb_ball_correlate_argmax_sep(const float* x_iter, const float* gauss, const int* wtab, const float* gauss_host,
                                  const int* wtab_host, int R, void* workspace, long long* argmax_out, int D, int H, int W,
                                  void* stream);
/* candidates {value bits, voxel index}: mode 0 = voxels of the ball (centre, grid half-width, radius^2 — insert_ball
 * :1336-1385) with x > 0 (also writes the 0/1 ball volume); mode 1 = voxels with mask != 0, value sigmoid(x) */
int rsb_ball_candidates(const float* x, const uint8_t* mask, int mode, int cz, int cy, int cx, int half, float radius2,
                        void* cand, int* n_cand, int max_cand, uint8_t* ball_out, int D, int H, int W, void* stream);
/* torch.topk membership x3 (:1470-1490) by exact ranking (value desc, index asc): m_i[idx] = rank < k_i */
int rsb_ball_rank_select(const void* cand, const int* n_cand, int max_cand, int k0, int k1, int k2, uint8_t* m0, uint8_t* m1,
                         uint8_t* m2, void* stream);
/* GlobalWeightedRankPooling weights (:442-537, hard cutoff) times N, scattered to voxel order */
int rsb_ball_rank_gwrp(const void* cand, const int* n_cand, int max_cand, float concentration, float* wmap, void* stream);
/* wmap = (pseudo ? wmap : 0) + (1 - dilated)   — foreground GWRP weights + background indicator (:1775-1811) */
int rsb_ball_weight_map(float* wmap, const uint8_t* pseudo, const uint8_t* dilated, long long V, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Optimizer end of the train step as two multi-tensor launches over a device table of
 * tensors:  torch.nn.utils.clip_grad_norm_(net.parameters(), max_norm)      (train_ddp.py:352)
 *           torch.optim.AdamW(lr, betas, eps = 1e-5, weight_decay).step()   (training/utils.py:46-51, train_ddp.py:353)
 *           update_ema_variables: ema = a * ema + (1 - a) * param           (training/utils.py:154-158)
 * Launch 1 reduces every gradient to per-block partial sums of squares (fixed order, no atomics); launch 2 re-sums
 * them in every block, derives clip = min(1, max_norm / (norm + 1e-6)) and updates g (clipped in place, like
 * clip_grad_norm_), p, exp_avg, exp_avg_sq and the EMA copy in one pass (40 B per parameter).
 * The table lives in device memory; chunk_begin[i] = sum_{j<i} ceil(n_j / rsb_opt_chunk_elems()).
 * ------------------------------------------------------------------------------------------ */
typedef struct RsbOptTensor {
  float* p;     /* parameter (fp32) */
  float* g;     /* gradient; multiplied by the clip coefficient in place.  NULL = no gradient this step: the row is left
                   out of the norm and of AdamW (m, v unused) and only its EMA copy moves */
  float* m;     /* exp_avg */
  float* v;     /* exp_avg_sq */
  float* ema;   /* EMA copy of the parameter, or NULL when has_ema = 0 */
  long long n;  /* elements */
  long long chunk_begin;
} RsbOptTensor;
long long rsb_opt_chunk_elems(void);
/* upper bound of the grid (= floats the `partials` workspace must hold) */
int rsb_opt_max_blocks(void);
/* step counts from 1 (bias corrections 1 - beta^step); max_norm <= 0 disables clipping (partials may then be NULL and
 * norm_out receives 0); norm_out (optional) receives the total gradient norm BEFORE clipping, the return value of
 * clip_grad_norm_.  Scalars are doubles like the python floats torch's AdamW computes with. */
int rsb_clip_adamw_ema_step(const RsbOptTensor* table_device, int n_tensors, long long total_chunks, int has_ema,
                            float* partials, float* norm_out, double max_norm, double lr, double beta1, double beta2,
                            double eps, double weight_decay, long long step, double ema_alpha, const float* hyper_device,
                            void* stream);
/* CUDA-graph mode: hyper_device != NULL points to rsb_opt_hyper_floats() floats in DEVICE memory holding the step-dependent
 * scalars; the kernel reads them at run time, so a captured launch stays valid while the host rewrites them before every
 * replay (bias corrections, LR schedule, EMA warm-up).  rsb_opt_fill_hyper computes that block on the HOST from the same
 * arguments (the by-value scalars of the call above are then ignored, except max_norm <= 0 deciding whether the norm
 * kernel is launched). */
int rsb_opt_hyper_floats(void);
int rsb_opt_fill_hyper(float* hyper_host, double max_norm, double lr, double beta1, double beta2, double eps,
                       double weight_decay, long long step, double ema_alpha);

/* ---------------------------------------------------------------------------------------------
 * Sliding-window inference post-processing: inference/inference3d.py:28-107 and
 * predict_abdomenatlas.py:637-710.  Volumes are fp32 / uint8 NCDHW like the reference's tensors.
 * ------------------------------------------------------------------------------------------ */
/* out[:, :, win] += sigmoid(pred) ; count[:, 0, win] += 1  for the window at (d0, h0, w0) of size (wd, wh, ww)
 * (inference3d.py:80-100).  pred = fp32 [B][C][wd][wh][ww] logits; pred == NULL: a gated-out window — zeros are added,
 * only the count moves (:92-100).  out fp32 [B][C][D][H][W], count fp32 [B][D][H][W]. */
int rsb_sigmoid_window_accumulate(const float* pred, float* out, float* count, int B, int C, int D, int H, int W, int wd,
                                  int wh, int ww, int d0, int h0, int w0, void* stream);
/* prob = acc / count (inference3d.py:102; prob may alias acc or be NULL) ; mask = prob > threshold (uint8 0/1, may be
 * NULL) — the `> 0.5` of predict_abdomenatlas.py:672 */
int rsb_blend_finalize(const float* acc, const float* count, float* prob, uint8_t* mask, float threshold, int B, int C,
                       long long V, void* stream);
/* ndi.binary_dilation(organ, structure = np.ones((3, 3, 3))) with border value 0 (predict_abdomenatlas.py:675) */
int rsb_dilate_box3(const uint8_t* src, uint8_t* dst, int n_vol, int D, int H, int W, void* stream);
/* prob *= organ (0/1)   (predict_abdomenatlas.py:678-680) */
int rsb_gate_by_mask(float* prob, const uint8_t* organ, long long n, void* stream);
/* Face-connected (6-neighbour) components of one uint8 volume — sitk.ConnectedComponentImageFilter with its default
 * FullyConnected = off (predict_abdomenatlas.py:690-695).  labels[v] = smallest linear index of v's component, -1 for
 * background (sorting the distinct roots reproduces the raster-scan numbering 1..n); *n_components = their number.
 * largest != NULL additionally writes keep_largest_component (:686-710): the first component (raster order) of maximal
 * size, all zeros when there is none; that needs `workspace` of rsb_cc_workspace_bytes(). */
size_t rsb_cc_workspace_bytes(int D, int H, int W);
int rsb_cc_label(const uint8_t* mask, int* labels, int* n_components, uint8_t* largest, void* workspace, int D, int H, int W,
                 void* stream);

/* ---------------------------------------------------------------------------------------------
 * Batch assembly from the on-disk crop format: label / unknown / chosen-segment masks are stored
 * np.packbits(bool[C][D][H][W], axis = 0) (dataset_abdomenatlas_UFO.py:952-975) and loaded with
 * np.unpackbits(...)[:C] (:1006-1015).  packed = uint8 [B][ceil(C/8)][V] uploaded as stored (8x fewer bytes than uint8
 * masks, 64x fewer than the int64 labels of train_ddp.py:262); out = uint8 [B][C][V] 0/1 (channel c = bit 7 - (c & 7)
 * of byte plane c >> 3), complemented when invert != 0.
 * ------------------------------------------------------------------------------------------ */
int rsb_unpack_masks(const uint8_t* packed, uint8_t* out, int B, int C, long long V, int invert, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Online intensity augmentations of the loader: training/augmentation.py:17-168 as applied by
 * training/dataset/dim3/dataset_abdomenatlas_UFO.py:1048-1061.  One fp32 volume of n voxels; the random draws
 * (factor, offset, gamma, sigma, noise) are made by the host exactly as the reference makes them.
 * ------------------------------------------------------------------------------------------ */
size_t rsb_aug_workspace_bytes(void);
/* stats4 = {min, max, mean, unbiased std} of the volume (augmentation.py:118-126, 152-158): deterministic two-stage
 * reduction in double */
int rsb_aug_stats(const float* x, long long n, void* workspace, float* stats4, void* stream);
/* y = ((x * mul) + add) + noise * noise_std with torch's separate roundings; each stage optional:
 * brightness_multiply (:86-103), brightness_additive (:69-83), gaussian_noise (:17-19; noise = N(0,1) samples or NULL) */
int rsb_aug_affine(const float* x, float* y, long long n, float mul, int has_mul, float add, int has_add,
                   const float* noise, float noise_std, void* stream);
/* gamma (:106-138): y = pow((x - min) / (max - min), gamma) * (max - min) + min with stats4 = rsb_aug_stats(x); then
 * (retain_stats) rsb_aug_renorm(y, stats(y), stats(x)): y = (y - mean_y) / std_y * std_x + mean_x, in place */
int rsb_aug_gamma(const float* x, float* y, long long n, const float* stats4, float gamma, void* stream);
int rsb_aug_renorm(float* y, long long n, const float* stats_y, const float* stats_x, void* stream);
/* contrast (:140-168): y = clamp((x - mean) * factor + mean, min, max) */
int rsb_aug_contrast(const float* x, float* y, long long n, const float* stats4, float factor, void* stream);
/* gaussian_blur (:48-66): one axis (0 = D, 1 = H, 2 = W) of the separable form of the reference's normalised 3-D
 * Gaussian, zero padding; taps_host = ntaps (odd, <= 33) weights in HOST memory (passed to the kernel by value) */
int rsb_aug_blur_axis(const float* x, float* y, int n_vol, int D, int H, int W, int axis, const float* taps_host,
                      int ntaps, void* stream);

/* ---------------------------------------------------------------------------------------------
 * MedFormer voxel-side kernels (rsuper_train/model/dim3/medformer_utils.py, conv_layers.py).
 * NDHWC tensors with a channel pitch, dtype RSB_BF16 / RSB_F32; small map-side tensors are fp32.
 * ------------------------------------------------------------------------------------------ */
/* nn.Conv3d(C, C, 3, padding=1, groups=C, bias=False) (conv_layers.py:125-157, 192-230): w fp32 [C,1,3,3,3];
 * flip = 1 mirrors the taps (the data gradient).  rsb_dwconv3_wgrad overwrites dw [C,1,3,3,3]. */
int rsb_dwconv3_forward(const void* a, int a_pitch, const float* w, void* y, int y_pitch, int dtype, int flip, int N, int D,
                        int H, int W, int C, void* stream);
int rsb_dwconv3_wgrad(const void* a, int a_pitch, const void* dy, int dy_pitch, int dtype, float* dw, int N, int D, int H,
                      int W, int C, void* stream);
/* SEBlock (conv_layers.py:159-173): y[n,v,c] = x[n,v,c] * s[n,c]; out[n,c] = sum_v a[n,v,c] * b[n,v,c] (overwritten) */
int rsb_scale_channels(const void* x, int x_pitch, const float* s, void* y, int y_pitch, int dtype, int N, long long V, int C,
                       void* stream);
int rsb_channel_dot(const void* a, int a_pitch, const void* b, int b_pitch, int dtype, float* out, int N, long long V, int C,
                    void* stream);
/* scratch (floats) for the column statistics of `rows` independent (sample[, head]) problems */
size_t rsb_colstats_workspace_floats(int rows);
/* SemanticMapGeneration (medformer_utils.py:222-235): smap[n,c,k] = sum_v feat[n,v,c] * softmax_v(logit[n,:,k])[v], K <= 27;
 * ms [N,64] receives the column (max, sum of exp) pairs the backward pass reuses.
 * backward: dS [N,C,K], tk[n,k] = sum_c dS[n,c,k] * smap[n,c,k]  ->  dfeat, dlogit (columns K..Kp-1 zeroed). */
int rsb_softmax_pool_forward(const void* feat, int feat_pitch, const void* logit, int logit_pitch, int dtype, float* ms,
                             float* workspace, float* smap, int N, long long V, int C, int K, void* stream);
int rsb_softmax_pool_backward(const void* feat, int feat_pitch, const void* logit, int logit_pitch, int dtype, const float* ms,
                              const float* dS, const float* tk, void* dfeat, int dfeat_pitch, void* dlogit, int dlogit_pitch,
                              int N, long long V, int C, int K, int Kp, void* stream);
/* BidirectionAttention (medformer_utils.py:67-103): q / fv = voxel-side query / value (channel c = dim * heads + head),
 * mq / mv = map-side query / value fp32 [N, heads, J, dim_head]; A = scale * q . mq; feat_out = softmax_j(A) mv (voxel side),
 * map_out = softmax_i(A)^T fv (fp32 [N, heads, J, dim_head]); ms [N*heads, 64] keeps the column statistics.
 * backward: tj[n,h,j] = dmap_out[n,h,j,:] . map_out[n,h,j,:]. */
int rsb_biattention_forward(const void* q, const void* fv, int qv_pitch, int dtype, const float* mq, const float* mv, float* ms,
                            float* workspace, void* feat_out, int feat_out_pitch, float* map_out, int N, long long V,
                            int heads, int dim_head, int J, float scale, void* stream);
int rsb_biattention_backward(const void* q, const void* fv, int qv_pitch, int dtype, const float* mq, const float* mv,
                             const float* ms, const void* dfeat_out, int dfeat_out_pitch, const float* dmap_out,
                             const float* tj, void* dq, void* dfv, int dqv_pitch, float* dmq, float* dmv, int N, long long V,
                             int heads, int dim_head, int J, float scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RSUPER_B200_H_ */
