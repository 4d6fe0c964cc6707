/* libvptr_b200.so -- C-ABI of the B200-native VPTR stage-2 hot path.
 *
 * The reference (XiYe20/VPTR) is pure PyTorch and has no FFI of its own: every kernel it runs is an ATen /
 * cuBLAS / cuDNN call made from model/*.py.  Each entry point below therefore cites the reference call site
 * (file:line under the reference root) whose library kernels it replaces.  The host side that mirrors the
 * reference's nn.Module API (vptr_b200/model) binds these with ctypes (INTEGRATION.md shows the stub).
 *
 * Conventions
 *   - all tensors fp32, device pointers borrowed for the duration of the call (never retained, allocated or
 *     freed here); activations are token-major / channel-last: rows ordered (n, t, h, w), C contiguous.
 *   - every function is asynchronous on `stream` and returns int: 0 ok, <0 invalid shape / alignment /
 *     unsupported configuration, >0 a cudaError_t.  vptr_last_error() gives the thread-local message.
 *   - nothing synchronises the device and nothing allocates, so every call is CUDA-graph capturable.
 */
#ifndef VPTR_B200_H
#define VPTR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* vptr_stream_t; /* == cudaStream_t */

int vptr_version(void);
const char* vptr_last_error(void);

/* ---- dense contractions (tcgen05 kind::tf32, TMEM accumulators, TMA operands) -------------------------
 * D[M,N] (+)= act(alpha * sum_k Aop[m,k]*Bop[n,k] + bias[n]) + residual[m,n]
 *   a_mn=0: A is [M][K] (pitch lda)   a_mn=1: A is [K][M]      b_mn=0: B is [N][K] (pitch ldb)   b_mn=1: B is [K][N]
 *   act: 0 none, 1 exact GELU, 2 ReLU.  flags: bit0 atomic accumulate into D (split-K allowed: k_splits 0 = auto),
 *   bit1 round stored values to tf32 (round-to-nearest) so a following tf32 contraction reads exact operands.
 * Replaces: q/k/v/out nn.Linear (model/MultiHeadAttentionRPE.py:543-545,688), nn.MultiheadAttention in/out
 * projections (model/VidHRFormer_modules.py:79-84,185-187,204-205), MlpDWBN 1x1 convs (:424-442), linear1/linear2
 * (:87-89,190-192), NCE_projector (model/VPTR_modules.py:133-135), and -- on im2col operands -- the ResNet 3x3
 * Conv2d / ConvTranspose2d (model/ResNetAutoEncoder.py:26-47,74-90), with their autograd dgrad / wgrad. */
int vptr_gemm_tf32(const float* A, long long lda, int a_mn, const float* B, long long ldb, int b_mn, float* D, long long ldd,
                   int M, int N, int K, const float* bias, const float* residual, long long ldr, float alpha, int act,
                   int flags, int k_splits, const float* rowscale, int rows_per_group, unsigned long long drop_seed, float drop_p,
                   vptr_stream_t stream);
/* debug aid: device buffer that CTA 0 of the next vptr_gemm_tf32 launches fills with clock64() stamps per tile; NULL disables */
int vptr_gemm_debug_buffer(long long* buf);
/* Branch regularisation shared by the entry points below: out = rowscale[row / rows_per_group] * dropout_p(...) (+ residual).
 * rowscale = DropPath keep-scales per clip (reference model/VidHRFormer_modules.py:563-575; vptr_droppath_scales) or NULL;
 * dropout masks come from a counter-based RNG keyed by (drop_seed, element index), so the backward regenerates them.
 * same contract on fp32 FFMA; for pitches TMA cannot address and as the on-device cross-check in tests */
int vptr_gemm_simt(const float* A, long long lda, int a_mn, const float* B, long long ldb, int b_mn, float* D, long long ldd,
                   int M, int N, int K, const float* bias, const float* residual, long long ldr, float alpha, int act,
                   int flags, int k_splits, const float* rowscale, int rows_per_group, unsigned long long drop_seed, float drop_p,
                   vptr_stream_t stream);

/* ---- LayerNorm over C (model/VidHRFormer_modules.py:44-56,137-161,25-26,114-115) ---------------------
 * y = LN(x)*gamma+beta [relu]; y2 = y + add[(row/add_div) % add_mod] (positional add of :75-84,176-178,200). */
int vptr_layernorm_fwd(const float* x, const float* gamma, const float* beta, float* y, float* y2, const float* add,
                       int add_div, int add_mod, float* mean, float* rstd, long long rows, int C, float eps, int relu,
                       int round_tf32, vptr_stream_t stream);
/* dx = dres + dLN(dy1 + dy2); dgamma/dbeta accumulated (+=); any of dy2, dres, dx, dgamma may be NULL */
int vptr_layernorm_bwd(const float* dy1, const float* dy2, const float* x, const float* gamma, const float* beta,
                       const float* mean, const float* rstd, const float* dres, float* dx, float* dgamma, float* dbeta,
                       long long rows, int C, int relu, vptr_stream_t stream);

/* ---- MlpDWBN norms (model/VidHRFormer_modules.py:397-400,424-442) ------------------------------------- */
int vptr_bn_stats(const float* x, long long rows, int ch, float* mean, float* rstd, float* running_mean, float* running_var,
                  float eps, float momentum, double* ws /* 2*ch doubles */, vptr_stream_t stream);
int vptr_bn_eval_stats(const float* running_mean, const float* running_var, float* mean, float* rstd, int ch, float eps,
                       vptr_stream_t stream);
int vptr_group_stats(const float* x, int groups, long long gsize, float* mean, float* rstd, float eps, vptr_stream_t stream);
/* y = GELU(norm(x)) (+res). mode 0 BatchNorm (per channel), 1 LayerNorm((ch,H,W)) per frame, affine laid [hw][ch] */
int vptr_norm_act_fwd(const float* x, float* y, const float* res, const float* mean, const float* rstd, const float* gamma,
                      const float* beta, long long rows, int ch, int hw, int mode, int round_tf32, const float* rowscale,
                      int rows_per_group, unsigned long long drop_seed, float drop_p, vptr_stream_t stream);
/* mode 0 train BatchNorm, 1 frame LayerNorm, 2 eval BatchNorm. ws: 2*ch (modes 0,2) or 2*frames (mode 1) floats */
int vptr_norm_act_bwd(const float* dy, const float* x, const float* mean, const float* rstd, const float* gamma,
                      const float* beta, float* dx, float* dgamma, float* dbeta, long long rows, int ch, int hw, int mode,
                      float* ws, int round_tf32, const float* rowscale, int rows_per_group, unsigned long long drop_seed, float drop_p,
                      vptr_stream_t stream);

/* ---- attention cores ---------------------------------------------------------------------------------
 * mode 0: local-window attention + relative-position bias (model/VidHRFormer_modules.py:321-357,503-525;
 *         model/MultiHeadAttentionRPE.py:586-590,623,635-650,677-686); F_or_N = frames.
 * mode 1: temporal / encoder-decoder attention per pixel (model/VidHRFormer_modules.py:79-84,185-187,204-205),
 *         causal = FAR mask of :78; F_or_N = clips.
 * Q,K,V,O are token-major with row pitches ld*; head h uses columns [h*d, (h+1)*d). */
int vptr_attn_fwd(const float* Q, long long ldq, const float* K, long long ldk, const float* V, long long ldv, float* O,
                  long long ldo, const float* rpe_table, int mode, int F_or_N, int H, int W, int ws, int Tq, int Tk, int nhead,
                  int d, int causal, float scale, int round_tf32, unsigned long long drop_seed, float drop_p, vptr_stream_t stream);
/* tcgen05 / TMA / TMEM forward of the same core (head_dim 66, groups <= 32 tokens, 4x4 windows or temporal sequences): same
 * arguments as vptr_attn_fwd; returns -3 (unsupported) outside its domain.  TF32 operands (~4e-4 relative when q/k/v were
 * rounded to tf32 by their producers).  vptr_attn_fwd routes here when VPTR_ATTN_TC=1; the default is the 3xTF32 mma path. */
int vptr_attn_fwd_tcgen05(const float* Q, long long ldq, const float* K, long long ldk, const float* V, long long ldv, float* O,
                          long long ldo, const float* rpe_table, int mode, int F_or_N, int H, int W, int ws, int Tq, int Tk, int nhead,
                          int d, int causal, float scale, int round_tf32, unsigned long long drop_seed, float drop_p,
                          vptr_stream_t stream);
int vptr_attn_bwd(const float* Q, long long ldq, const float* K, long long ldk, const float* V, long long ldv, const float* dO,
                  long long ldo, float* dQ, long long lddq, float* dK, long long lddk, float* dV, long long lddv,
                  const float* rpe_table, float* d_rpe_table, int mode, int F_or_N, int H, int W, int ws, int Tq, int Tk,
                  int nhead, int d, int causal, float scale, int round_tf32, unsigned long long drop_seed, float drop_p,
                  vptr_stream_t stream);
/* same, and dbq / dbk / dbv ([nhead*d], any may be NULL) += column sums of dQ / dK / dV: the bias gradients of the q / k / v
 * projections (in_proj_bias, q/k/v_proj.bias) come out of the attention backward itself instead of a pass over dq / dk / dv */
int vptr_attn_bwd_bias(const float* Q, long long ldq, const float* K, long long ldk, const float* V, long long ldv, const float* dO,
                  long long ldo, float* dQ, long long lddq, float* dK, long long lddk, float* dV, long long lddv,
                  const float* rpe_table, float* d_rpe_table, int mode, int F_or_N, int H, int W, int ws, int Tq, int Tk,
                  int nhead, int d, int causal, float scale, int round_tf32, unsigned long long drop_seed, float drop_p,
                  float* dbq, float* dbk, float* dbv, vptr_stream_t stream);
/* integer artefacts from the kernels' own index functions (bit-exact contract): relative_position_index
 * (model/MultiHeadAttentionRPE.py:373-387) as int64 [L][L]; window token map (model/VidHRFormer_modules.py:503-513)
 * as int64 [L][B]; causal mask (model/VidHRFormer_modules.py:78) as uint8 [T][T] */
int vptr_window_index_maps(int F, int H, int W, int ws, long long* rpi, long long* wmap, vptr_stream_t stream);
int vptr_causal_mask(int T, unsigned char* mask, vptr_stream_t stream);

/* ---- depthwise 3x3 of MlpDWBN (model/VidHRFormer_modules.py:405-410) ----------------------------------- */
int vptr_dwconv3x3(const float* x, const float* w9, const float* bias, float* y, int F, int H, int W, int ch, int flip,
                   vptr_stream_t stream);
/* forward depthwise conv + per-frame (sum, sum of squares) of its outputs into sums[0:F], sums[F:2F] (fp64, caller-zeroed): the
 * statistics of the LayerNorm((ch,H,W)) that follows (MlpDWBN norm2) without another pass.  -3 outside the streaming kernel's domain. */
int vptr_dwconv3x3_stats(const float* x, const float* w9, const float* bias, float* y, int F, int H, int W, int ch, double* sums,
                         vptr_stream_t stream);
/* mean / rstd per group from such (sum, sum of squares) pairs */
int vptr_group_stats_finalize(const double* sums, int groups, long long gsize, float* mean, float* rstd, float eps, vptr_stream_t stream);
int vptr_dwconv3x3_wgrad(const float* x, const float* dy, float* dw9, float* dbias, int F, int H, int W, int ch,
                         vptr_stream_t stream);

/* ---- elementwise / layout helpers ----------------------------------------------------------------------- */
int vptr_axpby(const float* a, const float* b, float* out, long long n, float alpha, float beta, vptr_stream_t stream);
int vptr_add_rows(const float* x, const float* add, float* out, long long rows, int C, int div, int mod, int round_tf32,
                  vptr_stream_t stream);
int vptr_rowgroup_sum(const float* dy, float* out, long long group_elems, int reps, vptr_stream_t stream);
int vptr_gelu_fwd(const float* x, float* y, long long n, int round_tf32, unsigned long long drop_seed, float drop_p,
                  vptr_stream_t stream);
int vptr_gelu_bwd(const float* dy, const float* x, float* dx, long long n, int round_tf32, unsigned long long drop_seed, float drop_p,
                  vptr_stream_t stream);
/* y = x rounded to nearest tf32: operands of vptr_gemm_tf32 are pre-rounded by their producers (or by this copy) so the
 * tensor core's mantissa truncation is exact and unbiased */
int vptr_round_copy(const float* x, float* y, long long n, int do_round, const float* rowscale, long long group_elems,
                    unsigned long long drop_seed, float drop_p, vptr_stream_t stream);
/* one launch for all GEMM weights of a pass: table = n device source pointers (as int64) followed by n cumulative end
 * offsets (float4 units) inside dst */
int vptr_round_copy_multi(const long long* table, int n, float* dst, long long total4, vptr_stream_t stream);
int vptr_droppath_scales(float* out, int n, unsigned long long seed, float p, vptr_stream_t stream);
int vptr_relu_fwd(const float* x, float* y, long long n, vptr_stream_t stream);
int vptr_relu_bwd(const float* dy, const float* y, float* dx, long long n, vptr_stream_t stream);
int vptr_colsum(const float* x, float* out, long long rows, int C, long long ld, vptr_stream_t stream);
int vptr_transpose(const float* in, float* out, int batch, int R, int C, int accumulate, vptr_stream_t stream);
/* n transposes in one launch: device table of n x 5 int64 {src, dst, R, C, cumulative 32x32-tile count}; dst[c][r] (+)= src[r][c]
 * (LayerNorm((ch,H,W)) affine weights and depthwise 3x3 weights into the engine's channel-last layout and their gradients back:
 * reference model/VidHRFormer_modules.py:386-442) */
int vptr_transpose_multi(const long long* table, int n, int total_tiles, int accumulate, vptr_stream_t stream);
/* PadBlock (model/VidHRFormer_modules.py:527-561): dir 0 centre zero-pad, dir 1 crop */
int vptr_pad_crop(const float* in, float* out, int F, int H, int W, int Hp, int Wp, int ph0, int pw0, int C, int dir,
                  vptr_stream_t stream);
/* Producers of dY that also emit its column sums (= the bias gradient of the Linear / 1x1 conv consuming dY), so that no separate
 * pass re-reads dY: colsum[c] += sum_r out[r][c] (caller provides an accumulator; the flat gradient buffer's bias slice). */
int vptr_round_copy_colsum(const float* x, float* y, long long rows, int C, int do_round, const float* rowscale, int group_rows,
                           unsigned long long drop_seed, float drop_p, float* colsum, vptr_stream_t stream);
int vptr_gelu_bwd_colsum(const float* dy, const float* x, float* dx, long long rows, int C, int round_tf32, unsigned long long drop_seed,
                         float drop_p, float* colsum, vptr_stream_t stream);
int vptr_norm_act_bwd_colsum(const float* dy, const float* x, const float* mean, const float* rstd, const float* gamma, const float* beta,
                             float* dx, float* dgamma, float* dbeta, long long rows, int ch, int hw, int mode, float* ws, int round_tf32,
                             const float* rowscale, int rows_per_group, unsigned long long drop_seed, float drop_p, float* colsum,
                             vptr_stream_t stream);
/* gradient clipping pieces (train_NAR.py:85): sqnorm += sum x^2 ; x *= min(1, max_norm/(sqrt(sqnorm)+1e-6)) */
int vptr_sqnorm_accumulate(const float* x, long long n, double* sqnorm_out, vptr_stream_t stream);
int vptr_clip_scale(float* x, long long n, const double* sqnorm, float max_norm, vptr_stream_t stream);

/* Implicit-GEMM 3x3 stride-1 convolution on the tcgen05 kernel: A tiles are 4-D TMA boxes of the padded NHWC activation
 * (no im2col matrix).  Replaces the 18 ResnetBlock convs (model/ResNetAutoEncoder.py:138,151) + their BN/ReLU/residual.
 * Returns -3 (unsupported) when H x W does not tile into 128-pixel boxes; callers then use vptr_im2col + vptr_gemm_tf32. */
int vptr_conv3x3_tf32(const float* xpad, const float* w, float* out, int F, int H, int W, int C, int Cout, const float* bias,
                      const float* residual, int act, int flags, int w_planes /* 1, or 2 = [hi|lo] tf32 weight split */,
                      vptr_stream_t stream);
/* the same convolution for H, W multiples of 8 beyond 8x8 (16x16 grid of 128x128 frames) on the raw-tile kernel: xq from
 * vptr_pad_nhwc_quad = every 8x8 quadrant with its own halo, [F*(H/8)*(W/8)][10][10][C] */
int vptr_conv3x3_tf32_quad(const float* xq, const float* w, float* out, int F, int H, int W, int C, int Cout, const float* bias,
                           const float* residual, int act, int flags, int w_planes, vptr_stream_t stream);
int vptr_pad_nhwc_quad(const float* x, float* out, int F, int H, int W, int C, int pad_mode, int round_tf32, vptr_stream_t stream);
/* the ResnetBlock convolution (reference model/ResNetAutoEncoder.py:138,151) with both operands as two bf16 planes (x = hi + lo,
 * w = hi + lo) and three bf16 tensor-core passes (hi*hi + lo*hi + hi*lo, fp32 accumulate): ~2^-16 per product instead of tf32's
 * 2^-11, at 1.5 instead of 2 TF32-pass equivalents.  H, W multiples of 8 (8x8 included), C % 8 == 0.
 *   xq2: vptr_pad_nhwc_quad_bf16x2 -> [2][F*(H/8)*(W/8)][10][10][C] bf16;  w2: vptr_split_bf16x2 -> [Cout][2][9*C] bf16 */
int vptr_conv3x3_bf16x3(const void* xq2, const void* w2, float* out, int F, int H, int W, int C, int Cout, const float* bias,
                        const float* residual, int act, int flags, vptr_stream_t stream);
int vptr_pad_nhwc_quad_bf16x2(const float* x, void* out, int F, int H, int W, int C, int pad_mode, vptr_stream_t stream);
int vptr_split_bf16x2(const float* w, void* out, long long rows, long long K, vptr_stream_t stream);
/* out[r] = [ rna_tf32(w[r]) | rna_tf32(w[r] - hi) ]: the two tf32 planes of a weight matrix (rows of K -> rows of 2K) */
int vptr_split_tf32(const float* w, float* out, long long rows, long long K, vptr_stream_t stream);
int vptr_pad_nhwc(const float* x, float* out, int F, int H, int W, int C, int pad, int pad_mode, int round_tf32, vptr_stream_t stream);

/* ---- ResNet encoder / decoder (model/ResNetAutoEncoder.py:26-48,70-98) ---------------------------------- */
int vptr_im2col(const float* x, const float* mask, float* col, int F, int H, int W, int Cin, int k, int stride, int pad,
                int pad_mode /* 0 zero, 1 reflect, 2 replicate */, int round_tf32, vptr_stream_t stream);
int vptr_convT_gather(const float* col, const float* shift, float* out, int F, int H, int W, int Cout, int relu,
                      vptr_stream_t stream);
int vptr_bn_fold(const float* gamma, const float* beta, const float* running_mean, const float* running_var, float eps,
                 float* scale, float* shift, int C, vptr_stream_t stream);
int vptr_pack_conv_weight(const float* w, const float* scale, float* out, int Co, int Ci, int k, int mode, vptr_stream_t stream);
int vptr_stem_conv7x7(const float* x, const float* wpk, const float* shift, float* out, int F, int Ci, int H, int W, int Co,
                      vptr_stream_t stream);
int vptr_head_conv7x7_fwd(const float* x, const float* wpk, const float* bias, float* out, int F, int Ci, int Co, int H, int W,
                          int act /* 0 none, 1 tanh, 2 sigmoid */, vptr_stream_t stream);
/* ws: F*(H+6)*(W+6)*Ci + 49*Co*Ci floats of scratch */
int vptr_head_conv7x7_bwd(const float* dout, const float* out, const float* w, float* dx, int F, int Ci, int Co, int H, int W,
                          int act, float* ws, vptr_stream_t stream);

/* ---- tail of the iteration: losses, global-norm clip, AdamW ---------------------------------- */
/* MSELoss + GDL(alpha=1) of cal_lossT (model/criterion.py:105-204, train_NAR.py:33-36; train_FAR.py:32-34) over `planes` = N*T*C
 * image planes of H x W.  sums: 3 device doubles zeroed by the caller; loss3 = {total, mse, gdl}. */
int vptr_mse_gdl_fwd(const float* pred, const float* target, long long planes, int H, int W, double* sums, float* loss3,
                     vptr_stream_t stream);
/* dpred = dloss * d(mse + gdl)/d(pred); dloss: device scalar or NULL (= 1) */
int vptr_mse_gdl_bwd(const float* pred, const float* target, const float* dloss, float* dpred, long long planes, int H, int W,
                     vptr_stream_t stream);
/* BiPatchNCE of cal_lossT (model/criterion.py:206-259 applied to F.normalize'd features, train_NAR.py:36), fused: one CTA per
 * frame; gt / pred are [F][L][C] channel-last rows, L = h*w <= 64.  S_save (F*64*64 floats) and stats_save (F*256 floats) carry the
 * cosine matrix and the row / column statistics to the backward; loss_sum: device double zeroed by the caller. */
int vptr_bipatch_nce_fwd(const float* gt, const float* pred, int F, int L, int C, float temperature, float* S_save, float* stats_save,
                         double* loss_sum, float* loss_out, vptr_stream_t stream);
int vptr_bipatch_nce_bwd(const float* gt, const float* pred, const float* S_save, const float* stats_save, const float* dloss, int F,
                         int L, int C, float temperature, float* dgt, float* dpred, vptr_stream_t stream);
/* sum of squares over n tensors with one launch (torch.nn.utils.clip_grad_norm_, train_NAR.py:85).  table: device int64
 * [n pointers][n cumulative unit ends]; vec != 0: units are float4 (16-byte aligned tensors, numel % 4 == 0), else floats */
int vptr_sqnorm_multi(const long long* table, int n, long long total_units, int vec, double* out, vptr_stream_t stream);
/* one AdamW step (torch.optim.AdamW semantics: decoupled weight decay, bias correction; optimizer_T.step(), train_NAR.py:86) over
 * n tensors with one launch.  table: device int64 [param][grad][exp_avg][exp_avg_sq][cumulative unit ends], n entries each.
 * sqnorm != NULL: the clip_grad_norm_(max_norm) coefficient min(1, max_norm/(sqrt(*sqnorm)+1e-6)) is applied to the gradients on
 * the fly, so the clip costs no pass of its own. */
int vptr_adamw_multi(const long long* table, int n, long long total_units, int vec, float lr, float beta1, float beta2, float eps,
                     float weight_decay, long long step, const double* sqnorm, float max_norm, vptr_stream_t stream);

/* ---- data-parallel gradient reduction (replaces DistributedDataParallel, train_NAR_mp.py:118,167-168) ------------ */
/* NCCL is bound at run time (dlopen of the process's libnccl.so.2).  A communicator is created from a 128-byte unique id made on
 * one rank (vptr_nccl_unique_id) and shipped to the others by the host (torch.distributed broadcast, MPI, a file ...). */
int vptr_nccl_unique_id(unsigned char* out128);
int vptr_nccl_comm_init(void** comm, int world, int rank, const unsigned char* id128);
int vptr_nccl_comm_destroy(void* comm);
/* flat[0..n) <- mean over ranks, in place (ncclAvg); then *sqnorm_out (device double, may be NULL) += sum of squares of the reduced
 * values on the same stream: the all-reduce of a finished gradient slice and its share of clip_grad_norm_'s norm in one call */
int vptr_allreduce_grads(void* comm, float* flat, long long n, double* sqnorm_out, vptr_stream_t stream);
/* scratch bytes of the entry points that take a caller-provided workspace: op 0 vptr_norm_act_bwd, 1 vptr_bn_stats,
 * 2 vptr_head_conv7x7_bwd (rows = F, ch = Ci, hw = H = W, mode = Co) */
long long vptr_workspace_bytes(int op, long long rows, int ch, int hw, int mode);

/* ---- stage-1 autoencoder training (train_AutoEncoder.py:44-86; model/ResNetAutoEncoder.py): train-mode BatchNorm2d and the
 * gradients stage 2 never needs.  Convolutions run as vptr_im2col + vptr_gemm_tf32 (forward, weight gradient) and
 * vptr_gemm_tf32 + vptr_col2im (input gradient). ------------------------------------------------------------------------- */
int vptr_stem_conv7x7_raw(const float* x, const float* wpk, float* out, int F, int Ci, int H, int W, int Co, vptr_stream_t stream);
/* z = act(BN(x; batch mean / rstd from vptr_bn_stats)) [+ res]; act 0 none, 1 ReLU, 2 ReLU after the residual add */
int vptr_bn_act_fwd(const float* x, float* z, const float* res, const float* mean, const float* rstd, const float* gamma,
                    const float* beta, long long rows, int ch, int act, int round_tf32, vptr_stream_t stream);
int vptr_bn_act_bwd(const float* dz, const float* x, const float* z, const float* mean, const float* rstd, const float* gamma,
                    float* g0, float* dx, float* dgamma, float* dbeta, long long rows, int ch, int act, float* ws, int round_tf32,
                    vptr_stream_t stream);
int vptr_col2im(const float* dcol, float* dx, int F, int H, int W, int C, int k, int stride, int pad, int pad_mode, vptr_stream_t stream);
int vptr_stem_wgrad(const float* x, const float* dy, float* dw, int F, int Ci, int H, int W, vptr_stream_t stream);
int vptr_act_bwd(const float* dout, const float* out, float* dpre, long long n, int act, vptr_stream_t stream);

/* ---- CUDA-graph support: a whole training step (every entry point above takes borrowed pointers and an explicit stream, allocates
 * nothing and never synchronises) can be captured once and replayed.  Two pieces of state must change between replays: ---------- */
/* the dropout / DropPath epoch mixed into every seed: reset >= 0 sets it (0 = eager default), reset < 0 increments it on the device */
int vptr_rng_advance(long long reset, vptr_stream_t stream);
/* a device-resident counter (the optimizer's step count): *ctr += inc */
int vptr_counter_add(long long* ctr, long long inc, vptr_stream_t stream);
/* vptr_adamw_multi with the update count read from device memory (step_dev) instead of the launch argument */
int vptr_adamw_multi_dev(const long long* table, int n, long long total_units, int vec, float lr, float beta1, float beta2, float eps,
                         float weight_decay, long long step, const long long* step_dev, const double* sqnorm, float max_norm,
                         vptr_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* VPTR_B200_H */
