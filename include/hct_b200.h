/*
 * hct_b200.h -- C ABI of libhct_b200.so: the B200 (sm_100a) kernels behind the HeadCT-Foundation
 * 3D-ViT hot path.
 *
 * The reference (nirvanesque/headCT_foundation) has NO FFI / plugin registry: every GPU
 * instruction it executes is a stock PyTorch library kernel called from Python.  The boundary a
 * maintainer binds is therefore "one C entry point per library call site on the hot path"; each
 * declaration below cites the reference call site(s) it replaces (file:line relative to the
 * reference tree).  INTEGRATION.md shows the ctypes stub the reference side would add.
 *
 * Conventions:
 *   - plain pointers + sizes; all pointers are DEVICE pointers unless stated otherwise;
 *   - every call enqueues work on the given stream (cudaStream_t passed as void*) and never
 *     synchronises the device.  Process-wide state: per-process caches (kernel attributes, the TMA
 *     encoder entry point), the hct_*_set_* switches and the hct_profile_* timing state -- plain
 *     globals, meant for ONE host thread per process (one process per GPU, as the reference runs);
 *     set a switch before the worker threads start if a host is threaded;
 *   - return value: 0 = ok, 1 = invalid argument, 2 = CUDA error, 3 = unsupported shape;
 *     hct_last_error() returns a thread-local message for the last non-zero return;
 *   - "bf16" = __nv_bfloat16 storage, row-major with an explicit leading dimension in ELEMENTS.
 */
#ifndef HCT_B200_H
#define HCT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* hct_stream_t; /* cudaStream_t */

const char* hct_last_error(void);
int hct_abi_version(void);
/* number of kernels launched by this library in this process (bench.py's gpu_launches) */
long long hct_launch_count(void);
/* Measurement aid (bench.py roofline): when enabled every hct_gemm_bf16 launch is bracketed by CUDA
 * events on its own stream; collect() returns the summed kernel time, 2*M*N*K flops and launch count
 * since the previous collect (it waits for those events). */
int hct_profile_enable(int on);     /* bit mask of (1 << class); 1 = GEMM only, 0 = off */
int hct_profile_collect(double* total_ms, double* total_flops, long long* launches);
/* The same per kernel class (bench.py roofline_secondary).  class: 0 GEMM (flops), 1 attention forward (flops = 4 S^2 D per
 * (batch, layer)), 2 attention backward (flops = 2 x forward, delta / tail launches included in the time), 3 LayerNorm
 * forward, 4 LayerNorm backward, 5 masked-MSE loss, 6 clip + AdamW, 7 HU windowing, 8 patchify (3..8: algorithmic bytes). */
int hct_profile_collect_class(int cls, double* total_ms, double* total_work, long long* launches);

/* ---------------------------------------------------------------------------------------------
 * GEMM core (tcgen05 / TMEM / TMA):  C[M,N] = epilogue( A[M,K] * B[N,K]^T )
 * replaces every nn.Linear / Conv3d(k=s) forward + backward on the path:
 *   attentionblock.py:54 (qkv), :64 (proj); monai MLPBlock linear1/linear2 (attentionblock.py:98);
 *   patch_embedding.py:149 (Conv3d as GEMM); mae.py:255 (decoder_embed), :272 (decoder_pred);
 *   dino_head.py:38,40 (head MLP, prototype layer) -- and their autograd dgrad / wgrad.
 * Operand storage:  K-major  : X[row * ld + k]        (row = m for A, n for B)
 *                   MN-major : X[k * ld + row]        (the transposed storage; used by dgrad/wgrad)
 * ------------------------------------------------------------------------------------------- */
enum hct_epilogue {
  HCT_EPI_BF16 = 0,       /* out_bf16 = alpha*acc + bias                                        */
  HCT_EPI_GELU_BF16 = 1,  /* t = acc + bias; out2_bf16 = t (if out2); out_bf16 = gelu_erf(t)     */
  HCT_EPI_RES_F32 = 2,    /* out_f32 = res_f32 + acc + bias          (out may alias res)         */
  HCT_EPI_POS_F32 = 3,    /* out_f32[remap(r)] = acc + bias + pos[pos_idx ? pos_idx[r] : r % pos_period] */
  HCT_EPI_DGELU_BF16 = 4, /* out_bf16 = acc * gelu_erf'(aux_bf16)                                */
  HCT_EPI_F32 = 5,        /* out_f32 = alpha*acc + bias                                          */
  HCT_EPI_ATOMIC_F32 = 6, /* out_f32 += alpha*acc   (split-K reduction; out pre-zeroed by caller) */
  HCT_EPI_GELU_DERIV_BF16 = 7, /* t = acc + bias; out_bf16 = gelu_erf(t); out2_bf16 = gelu_erf'(t): the forward
                                  saves the derivative, so that the backward of the activation is ...          */
  HCT_EPI_MUL_BF16 = 8    /* ... out_bf16 = acc * aux_bf16  (no special-function work in the dgrad epilogue)   */
};

typedef struct hct_gemm_desc {
  int32_t M, N, K;
  const void* A; int64_t lda; int32_t a_mn_major;
  const void* B; int64_t ldb; int32_t b_mn_major;
  int32_t epilogue;
  void* out;  int64_t ldo;
  void* out2; int64_t ldo2;
  const float* bias;                 /* [N] fp32 or NULL */
  const float* res; int64_t ldres;   /* fp32 residual */
  const void* aux;  int64_t ldaux;   /* bf16 pre-activation (DGELU) or multiplicand (MUL) */
  const float* pos; int64_t ldpos;   /* fp32 position table */
  const int32_t* pos_idx;            /* per-row index into pos, or NULL */
  int32_t pos_period;
  /* output row remap: r -> (r / rows_in) * rows_out + (r % rows_in) + row_off; rows that fall
   * outside [0, rows_out) within their group are skipped.  rows_in == 0 disables the remap. */
  int32_t rows_in, rows_out, row_off;
  float alpha;
  int32_t splits;                    /* split-K factor for HCT_EPI_ATOMIC_F32; 0 = auto */
  float* colsum;                     /* optional fp32 [N], += column sums of the bf16 output (bias gradient
                                        of the Linear that consumes it); bf16-output epilogues only */
} hct_gemm_desc;

int hct_gemm_bf16(const hct_gemm_desc* desc, hct_stream_t stream);
/* 1 (default): CTA-pair kernel (tcgen05 cta_group::2, 256x256 tiles); 0: single-CTA kernel (128x256 tiles). */
int hct_gemm_set_cta_pair(int enable);
/* Diagnostics: clock64 timeline of CTA 0 (MMA warp and first epilogue warp, per tile); buf = device buffer of
 * >= 1536 int64 or NULL to switch it off (tools/gemm_dbg.py). */
int hct_gemm_trace(void* buf);

/* ---------------------------------------------------------------------------------------------
 * Row kernels (HBM-bound)
 * ------------------------------------------------------------------------------------------- */
/* nn.LayerNorm forward: attentionblock.py:97-98, mae.py:240,271, vit.py:169.
 * x fp32 [rows, dim] -> y (bf16 if y_bf16 else fp32) ; mean/rstd fp32 [rows] (may be NULL).
 * beta == NULL selects RMSNorm (src/models/layers.py:29-53, the NORM_LAYER: 'rmsnorm' option of
 * main_downstream.py:111-116): y = x * rsqrt(mean(x^2) + eps) * gamma, rstd_out = that rsqrt, mean_out is not written. */
int hct_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* y, int y_bf16,
                      float* mean, float* rstd, int64_t rows, int32_t dim, float eps,
                      hct_stream_t stream);
/* LayerNorm backward (autograd of the call sites above).  dy: bf16 (dy_bf16) or fp32 [rows, dim].
 * dx_out_f32 = (dres_in ? dres_in : 0) + dLN ; optional bf16 copy dx_out_bf16 ;
 * dgamma/dbeta fp32 [dim] are ACCUMULATED (+=) with atomics; dxsum (optional, fp32 [dim], +=) receives
 * the column sums of the bf16 dx output = the bias gradient of the Linear that consumes it.
 * mean == NULL selects the RMSNorm backward (dbeta must then be NULL). */
/* 1 (default): rows are staged with cp.async.bulk (TMA unit, mbarrier completion); 0: per-lane cp.async */
int hct_layernorm_set_bulk(int enable);
int hct_layernorm_bwd(const void* dy, int dy_bf16, const float* x, const float* gamma,
                      const float* mean, const float* rstd, const float* dres_in,
                      float* dx_out_f32, void* dx_out_bf16, float* dgamma, float* dbeta, float* dxsum,
                      int64_t rows, int32_t dim, hct_stream_t stream);
/* fp32 -> bf16 cast of a contiguous buffer (autocast weight casts). */
int hct_cast_f32_to_bf16(const float* src, void* dst, int64_t n, hct_stream_t stream);
int hct_cast_bf16_to_f32(const void* src, float* dst, int64_t n, hct_stream_t stream);
/* out_bf16 = dy_bf16 * gelu_erf'(pre_bf16)  (nn.GELU backward in the DINO head MLP, dino_head.py:16-21) */
int hct_gelu_bwd(const void* dy, const void* pre, void* out, int64_t n, hct_stream_t stream);
/* out[c] += sum_r X[r, c]   (bias gradients; X bf16 or fp32, ld in elements). */
int hct_colsum(const void* x, int x_bf16, int64_t ld, float* out, int64_t rows, int32_t cols,
               hct_stream_t stream);
/* dst[b, row_off + r, :] = src[r, :] for b < batch, r < nrows  (cls / register tokens:
 * mae.py:233-234, vit.py:147-160). */
int hct_broadcast_rows(const float* src, float* dst, int32_t batch, int32_t nrows,
                       int64_t dst_rows_per_batch, int32_t row_off, int32_t dim, hct_stream_t stream);
/* out[r, :] += sum_b src[b, row_off + r, :]  (gradient of the above). */
int hct_reduce_rows(const float* src, float* out, int32_t batch, int32_t nrows,
                    int64_t src_rows_per_batch, int32_t row_off, int32_t dim, hct_stream_t stream);
/* fp32 rows -> bf16 rows with a per-group row window (slices the cls/register prefix off a gradient
 * or casts a GEMM operand):  dst[g*rows_per_group + r, :] = bf16(src[(g*src_rows_per_group + src_row_off + r) * src_ld ...]) */
int hct_copy_rows_f32_to_bf16(const float* src, int64_t src_ld, int64_t src_rows_per_group,
                              int32_t src_row_off, void* dst, int64_t dst_ld, int64_t groups,
                              int32_t rows_per_group, int32_t dim, hct_stream_t stream);
/* out[idx[r], :] += src[r, :]  (fp32 atomics; gradient of the position-embedding gather for the
 * kept patches, patch_embedding.py:156 under mae.py:211-212).  src bf16 [rows, dim]. */
int hct_scatter_add_rows(const void* src_bf16, const int32_t* idx, float* out, int64_t rows,
                         int32_t dim, hct_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * MAE-specific kernels
 * ------------------------------------------------------------------------------------------- */
/* MultipleWindowScaleStack (src/data/transforms.py:22-36, windows :130): HU -> 3 clipped windows.
 * hu: [nvol, 1, vox] fp32 (hu_i16 = 0) or int16 (hu_i16 = 1); out: [nvol, nwin, vox] fp32 or bf16.
 * a_min/a_max: HOST arrays of nwin floats. */
int hct_window_scale_stack(const void* hu, int hu_i16, void* out, int out_bf16, int64_t nvol,
                           int64_t vox, int32_t nwin, const float* a_min, const float* a_max,
                           hct_stream_t stream);
/* im2col for Conv3d(k = s = patch) (patch_embedding.py:149): x fp32 [B,C,H,W,D] -> cols bf16
 * [rows, C*p^3] in (c,ph,pw,pd) order.  patch_ids (int64 [B, rows_per_vol]) selects which patches
 * are materialised (MAE keeps 25 %: mae.py:211-212); NULL = all patches in (gh,gw,gd) order.
 * pos_idx_out (int32 [rows], may be NULL) receives the patch index of each row. */
int hct_patchify(const float* x, void* cols, const int64_t* patch_ids, int32_t* pos_idx_out,
                 int32_t B, int32_t C, int32_t H, int32_t W, int32_t D, int32_t p,
                 int32_t rows_per_vol, hct_stream_t stream);
/* random_masking index part (mae.py:205-216): stable ascending argsort of noise rows.
 * noise fp32 [N, L] -> ids_restore int64 [N, L], ids_keep int64 [N, len_keep], mask fp32 [N, L]. */
int hct_mask_indices(const float* noise, int64_t* ids_restore, int64_t* ids_keep, float* mask,
                     int32_t N, int32_t L, int32_t len_keep, hct_stream_t stream);
/* token gather (mae.py:212): dst[n, row_off + j, :] = src[n, ids[n, j], :]   (fp32 rows). */
int hct_gather_tokens(const float* src, const int64_t* ids, float* dst, int32_t N, int32_t L,
                      int32_t n_ids, int64_t dst_rows_per_batch, int32_t row_off, int32_t dim,
                      hct_stream_t stream);
/* gradient of the gather: dsrc[n, ids[n,j], :] = ddst[n, row_off + j, :], other rows zero. */
int hct_scatter_tokens(const float* ddst, const int64_t* ids, float* dsrc, int32_t N, int32_t L,
                       int32_t n_ids, int64_t ddst_rows_per_batch, int32_t row_off, int32_t dim,
                       hct_stream_t stream);
/* decoder input assembly (mae.py:257-265): y bf16 [N, 1+keep, dim] (decoder_embed output) ->
 * out fp32 [N, 1+L, dim]: row 0 = y[:,0] + dec_cls ; row 1+l = (ids_restore[n,l] < keep ?
 * y[n, 1+ids_restore[n,l]] : mask_token) + dec_pos[l]. */
int hct_decoder_assemble(const void* y, const int64_t* ids_restore, const float* mask_token,
                         const float* dec_cls, const float* dec_pos, float* out, int32_t N,
                         int32_t L, int32_t keep, int32_t dim, hct_stream_t stream);
/* its gradient: dy bf16 [N, 1+keep, dim]; dmask_token / ddec_cls fp32 [dim] accumulated (+=). */
int hct_decoder_assemble_bwd(const float* dout, const int64_t* ids_restore, void* dy,
                             float* dmask_token, float* ddec_cls, int32_t N, int32_t L,
                             int32_t keep, int32_t dim, hct_stream_t stream);
/* patchify + forward_loss (mae.py:150-170, :289-299).  pred bf16 [N, prefix + L, p^3*C] in
 * (ph,pw,pd,c) order -- `pred_prefix_rows` leading rows per sample (the cls row, dropped at
 * mae.py:273) are skipped; imgs fp32 [N,C,H,W,D], mask fp32 [N,L].  loss_out fp32 [4 + N*L]:
 * [0] = sum(mask*mse), [1] = sum(mask), [2] = loss = [0]/[1], [4..) = per-patch MSE workspace.
 * The final reduction is a fixed-order tree (deterministic). */
int hct_mae_loss_fwd(const void* pred, int32_t pred_prefix_rows, const float* imgs,
                     const float* mask, float* loss_out, int32_t N, int32_t C, int32_t H, int32_t W,
                     int32_t D, int32_t p, int32_t norm_pix, hct_stream_t stream);
/* dpred bf16 (same layout as pred, may alias it) = dloss * mask * 2 (pred - target) / (P * sum(mask));
 * rows with mask == 0 and the prefix rows are written as zeros.  dloss, mask_sum: DEVICE scalars. */
int hct_mae_loss_bwd(const void* pred, int32_t pred_prefix_rows, const float* imgs,
                     const float* mask, const float* dloss, const float* mask_sum, void* dpred,
                     int32_t N, int32_t C, int32_t H, int32_t W, int32_t D, int32_t p,
                     int32_t norm_pix, hct_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Attention (F.scaled_dot_product_attention, attentionblock.py:61; no mask, no dropout)
 * qkv bf16 [B, S, 3, H, hd] (the qkv Linear's natural output, attentionblock.py:54);
 * out bf16 [B, S, H*hd]; lse fp32 [B, H, S].  hd in {32, 48, 64}.
 * ------------------------------------------------------------------------------------------- */
int hct_attention_fwd(const void* qkv, void* out, float* lse, int32_t B, int32_t S, int32_t H,
                      int32_t hd, hct_stream_t stream);
/* hd 64/48 kernel choice.  2 (default): tcgen05/TMEM kernels for every forward tile and every full 128-row backward
 * tile, a single row behind the last full tile (S = 128 k + 1 with the cls token) of dQ/dK/dV on a CUDA-core row kernel;
 * 3: tcgen05 for every tile, backward tail tile included; 1: tcgen05, but the forward's S %% 128 <= 32 tail rows on the
 * mma.sync kernel; 0: mma.sync kernels only */
int hct_attention_set_tcgen05(int mode);
/* 1 (default): when the last 64-wide block of the tcgen05 backward kernels holds <= 16 rows (S = 64 k + 1 with the cls
 * token) it is computed together with block 0 -- one MMA -> softmax -> MMA chain step less; 0: as its own step */
int hct_attention_set_merge_tail(int enable);
/* 1: the dK/dV backward kernel works on 32-query blocks with two S^T / dP^T buffer pairs in tensor memory, so the MMAs
 * of block i+1 overlap the softmax of block i (measured: not faster); 0 (default): 64-query blocks,
 * single-buffered (the merge_tail switch applies) */
int hct_attention_set_dkdv32(int enable);
/* 1: backward on the pipelined persistent kernels of hct_attention_bwd3.cu (one CTA per SM, three score-buffer pairs in
 * tensor memory, two softmax warp groups); 0: the two-CTA-per-SM kernels.  Same results either way. */
int hct_attention_set_bwd3(int enable);
/* 1: forward on the pipelined persistent kernel of hct_attention_fwd2.cu (one CTA per SM, two query tiles of a head in
 * flight, two score buffers per tile, sixteen softmax warps); 0: four CTAs per SM, one tile each.  Same results up to the
 * rounding of the probabilities (the two kernels move the running maximum at the same points). */
int hct_attention_set_fwd2(int enable);
/* Bit mask, default 3.  Bit 0: when S = 64 k + 1 (the cls token) the tcgen05 forward does not spend a chain step (S MMA ->
 * softmax -> P V MMA) on the single key behind the last full 64-key block: its score is a dot product per query row and
 * its contribution a rank-1 update, both folded into the epilogue in fp32.  Bit 1: when S = 128 k + 1 the single query row
 * behind the last full 128-row tile does not get a CTA of its own: the producer warp of the last full tile's CTA runs its
 * online softmax on the CUDA cores against the K / V blocks that pass through shared memory.  0: both get their own block /
 * tile (A/B).  Results agree to the bf16 rounding of the probabilities involved. */
int hct_attention_set_tail_key(int fold);
/* Softmax arithmetic of the tcgen05 forward (fwd) and of the pipelined backward (bwd): -1 = scalar fp32; 0 = the packed
 * two-lane fp32 instructions of sm_100 (FFMA2 / FADD2 / FMUL2); n > 0 = packed, and n of every 8 exponential pairs evaluated
 * on the FMA pipe (Cody-Waite split + degree-3 polynomial, relative error 7.5e-5) instead of MUFU.EX2 (fwd: 3, bwd: 2);
 * -2 leaves a setting unchanged.  Defaults are the measured fastest: fwd -1, bwd 0 (profiles/r02_attn_softmax_ab_v2.txt).
 * A throughput knob: probabilities are rounded to bf16 (2^-9) right after, so results agree to that rounding. */
int hct_attention_set_poly(int fwd, int bwd);
/* 1: the GEMM, attention and LayerNorm kernels are launched with programmatic stream serialization -- the next grid's CTAs
 * start on an SM as soon as the previous grid's CTA there has exited and wait (griddepcontrol.wait) for the rest of it
 * after their prologue; 0 (default): plain stream order.  Same results.  Worth 5-7 us per launch where the GPU is not at
 * its power limit (sequences of small GEMMs); neutral on the batch-256 training step. */
int hct_set_pdl(int enable);
/* 1 (default): the pipelined backward hands its dQ / dK / dV tiles to cp.async.bulk.tensor stores (when every 128-row tile
 * is full); 0: per-lane store loop (A/B comparison).  Same results. */
int hct_attention_set_bwd3_drain(int tma);
/* debugging aid: clock64 timeline of CTA 0 of the bwd3 dK/dV kernel into a device buffer of >= 4096 int64 (NULL = off) */
int hct_attention_trace3(void* buf);
/* dqkv bf16 same layout as qkv.  delta_ws: fp32 workspace [B, H, S]. */
/* Diagnostics: clock64 event timeline of one CTA of the dK/dV backward kernel.  buf = device buffer of >= 768 int64
 * (layout: [producer | MMA | softmax warp 0][block][8 events], see tools/attn_dbg.py) or NULL to switch it off. */
int hct_attention_trace(void* buf);
int hct_attention_bwd(const void* qkv, const void* out, const void* dout, const float* lse,
                      void* dqkv, float* delta_ws, int32_t B, int32_t S, int32_t H, int32_t hd,
                      hct_stream_t stream);
/* The same, and the qkv-bias gradient of the block's qkv Linear (attentionblock.py:36, qkv_bias=True in the shipped configs):
 * the column sums of dqkv over all B * S tokens are ADDED to dqkv_colsum (fp32 [3 * H * hd], zeroed by the caller; NULL = none).
 * The pipelined tcgen05 backward sums them from the tiles it stages for its stores (no second pass over dqkv); the other
 * paths run hct_colsum behind the kernels.  Accumulation order differs between the two (fp32 atomics). */
int hct_attention_bwd_bias(const void* qkv, const void* out, const void* dout, const float* lse,
                           void* dqkv, float* delta_ws, float* dqkv_colsum, int32_t B, int32_t S, int32_t H,
                           int32_t hd, hct_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * DINO kernels
 * ------------------------------------------------------------------------------------------- */
/* F.normalize(p=2, dim=-1) (dino_head.py:39): x bf16/fp32 [rows, dim] -> y bf16, inv_norm fp32[rows] */
int hct_l2norm_fwd(const void* x, int x_bf16, void* y_bf16, float* inv_norm, int64_t rows, int32_t dim,
                   hct_stream_t stream);
int hct_l2norm_bwd(const void* dy_bf16, const void* y_bf16, const float* inv_norm, void* dx_bf16,
                   int64_t rows, int32_t dim, hct_stream_t stream);
/* weight_norm (dino_head.py:26-29): w_bf16[r,:] = g[r] * v[r,:] / ||v[r,:]|| ; inv_norm fp32[rows] */
int hct_weightnorm_fwd(const float* v, const float* g, void* w_bf16, float* inv_norm, int64_t rows,
                       int32_t dim, hct_stream_t stream);
/* dv = g*inv*(dw - (dw . vhat) vhat); dg[r] = dw[r,:] . vhat[r,:].  dv / dg may be NULL (g is frozen when
 * norm_last_layer=True, dino_head.py:28-29; with norm_last_layer=False the gain trains). */
int hct_weightnorm_bwd(const float* dw, const float* v, const float* g, const float* inv_norm,
                       float* dv, float* dg, int64_t rows, int32_t dim, hct_stream_t stream);
/* DINOLoss.forward (losses.py:63-89).  student fp32 [ncrops*B, K], teacher fp32 [2*B, K],
 * center fp32 [K].  loss_out fp32[1] (may be NULL) is ACCUMULATED into (zero it first).
 * stats_ws: fp32 [2*(ncrops+2)*B] (row max / log-sum-exp; recomputed on every call).
 * If dstudent (bf16 [ncrops*B, K]) is non-NULL the gradient wrt student is written, scaled by
 * *dloss (DEVICE scalar, NULL = 1). */
int hct_dino_loss(const float* student, const float* teacher, const float* center, float* loss_out,
                  float* stats_ws, void* dstudent_bf16, const float* dloss, int32_t B, int32_t ncrops,
                  int32_t K, float student_temp, float teacher_temp, hct_stream_t stream);
/* update_center (losses.py:91-102) split around the all-reduce: colsum then EMA.
 *   hct_colsum(teacher...) -> batch_center ; [all-reduce] ; center = m*center + (1-m)*bc/denom */
int hct_center_ema(float* center, const float* batch_center_sum, float denom, float momentum,
                   int32_t K, hct_stream_t stream);
/* _update_momentum_encoder (misc.py:386-397): one launch over a device table of n tensors.
 * table: int64 [n, 4] = {teacher_ptr, student_ptr, numel, teacher_bf16_shadow_ptr or 0}. p_k = m*p_k + (1-m)*p_q (fp32);
 * when the shadow pointer is set the bf16 copy of the new p_k (the GEMM operand of the next forward) is written too. */
int hct_ema_multi(const int64_t* table, int32_t n, float m, hct_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Train-time augmentation of volumes that already sit in HBM (MAE / ViT chain):
 * mae3d_transforms, src/data/transforms.py:195-236.  Random draws stay on the host (MONAI draws them with a numpy
 * RandomState); the kernels apply them.
 * ------------------------------------------------------------------------------------------- */
/* CastToTyped(float32) + RandFlipd on spatial axes 0/1/2 + RandShiftIntensityd (transforms.py:195-223):
 * in: fp16 (in_f16, the cached format of cpu_caching.py) or fp32 [nvol, C, D0, D1, D2]; out fp32 same shape;
 * flip_bits[v]: bit k set = reverse spatial axis k (NULL = none); offsets[v]: added to every voxel (NULL = 0). */
int hct_flip_shift(const void* in, int32_t in_f16, float* out, const uint8_t* flip_bits, const float* offsets,
                   int64_t nvol, int32_t C, int32_t D0, int32_t D1, int32_t D2, hct_stream_t stream);
/* One axis of RandGaussianSmoothd (transforms.py:228-236; MONAI GaussianFilter is separable, zero-padded), applied to
 * n samples: sample s is volume in_idx[s] of `in` (NULL: s) and is written to volume out_idx[s] of `out` (NULL: s) --
 * only the smoothed fraction of a batch is touched.  out[.., p, ..] = sum_k taps[s][k + radius] * in[.., p + k, ..];
 * taps fp32 [n, 2*radius+1]; volumes fp32 [C, D0, D1, D2]; in != out. */
int hct_gaussian_smooth_axis(const float* in, const int32_t* in_idx, float* out, const int32_t* out_idx,
                             const float* taps, int32_t radius, int64_t n, int32_t C, int32_t D0, int32_t D1,
                             int32_t D2, int32_t axis, hct_stream_t stream);

/* DINO multi-crop (DataAugmentationDINO3D, src/data/transforms.py:39-105): ResizeWithPadOrCrop + CenterSpatialCrop +
 * RandSpatialCrop + Resize(mode="area") of one crop as a single gather, with the RandFlip x 3 + RandShiftIntensity
 * that follow the resize (transforms.py:61-66) folded in.  boxes: int32 [ncrops, 8] = {sample, start0, start1, start2,
 * size0, size1, size2, flip_bits}; starts are SOURCE voxel coordinates (negative / past the end = the zero padding of
 * ResizeWithPadOrCrop); flip bit k reverses OUTPUT axis k; offsets fp32 [ncrops] or NULL is added last.
 * src fp16 / fp32 [B, C, S0, S1, S2]; out fp32 [ncrops, C, T0, T1, T2].
 * Area resize = adaptive average pooling, windows [floor(i n / T), ceil((i + 1) n / T)). */
/* 1 (default): row-staged kernel when the source rows allow it (S2 %% 8 == 0, S2 <= 256); 0: per-voxel gather only */
int hct_crop_resize_set_rows(int enable);
int hct_crop_resize_area(const void* src, int32_t src_f16, const int32_t* boxes, const float* offsets, float* out, int64_t ncrops,
                         int32_t C, int32_t S0, int32_t S1, int32_t S2, int32_t T0, int32_t T1, int32_t T2,
                         hct_stream_t stream);
/* RandAdjustContrast (transforms.py:92; MONAI AdjustContrast): x <- ((x - min) / (range + 1e-7))^gamma * range + min
 * per sample, in place, for the samples with gamma[s] > 0.  x fp32 [nsamples, per_sample]; minmax_ws int32 [2 nsamples]. */
int hct_adjust_contrast(float* x, const float* gamma, int32_t* minmax_ws, int64_t nsamples, int64_t per_sample,
                        hct_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * fp32 mode (the reference with --use_amp off: engine_pretrain_mae.py:57, engine_pretrain_dino.py:73,
 * engine_downstream.py:84 -- autocast(enabled=use_amp)).  Activations stay fp32 between kernels.
 * GEMMs run on hct_gemm_bf16 with every fp32 operand split into two bf16 terms (hi = bf16(x), lo = bf16(x - hi)) and the
 * three leading products hi*hi + hi*lo + lo*hi obtained by concatenating the terms along the contraction dimension
 * (K' = 3K): operand relative precision 2^-17, fp32 accumulation.
 * ------------------------------------------------------------------------------------------- */
/* dst (bf16) = the 3-term form of src (fp32 rows gathered as in hct_copy_rows_f32_to_bf16).  role_b = 0: terms
 * (hi, hi, lo) for the A operand, 1: (hi, lo, hi) for the B operand.  stack = 0: dst [rows, 3*cols] (K-major operand, terms
 * side by side), 1: dst [3*rows, cols] (MN-major operand, terms as row blocks). */
int hct_split3_bf16(const float* src, int64_t src_ld, int64_t src_rows_per_group, int32_t src_row_off,
                    int32_t rows_per_group, void* dst, int64_t rows, int32_t cols, int32_t role_b, int32_t stack,
                    hct_stream_t stream);
/* exact-erf GELU (monai MLPBlock act, nn.GELU(approximate='none')) and dy * gelu'(x), elementwise fp32 */
int hct_gelu_f32(const float* x, float* y, int64_t n, hct_stream_t stream);
int hct_gelu_bwd_f32(const float* dy, const float* x, float* dx, int64_t n, hct_stream_t stream);
/* fp32 row gather (same row mapping as hct_copy_rows_f32_to_bf16) and fp32 row scatter-add */
int hct_copy_rows_f32(const float* src, int64_t src_ld, int64_t src_rows_per_group, int32_t src_row_off, float* dst,
                      int64_t groups, int32_t rows_per_group, int32_t cols, hct_stream_t stream);
int hct_scatter_add_rows_f32(const float* src, const int32_t* idx, float* out, int64_t rows, int32_t dim, hct_stream_t stream);
/* F.scaled_dot_product_attention (attentionblock.py:61) in fp32: qkv / dqkv fp32 [B,S,3,H,hd], out / dout fp32 [B,S,H*hd],
 * lse / delta_ws fp32 [B,H,S].  hd in {32, 48, 64}. */
int hct_attention_f32_fwd(const float* qkv, float* out, float* lse, int32_t B, int32_t S, int32_t H, int32_t hd,
                          hct_stream_t stream);
int hct_attention_f32_bwd(const float* qkv, const float* out, const float* dout, const float* lse, float* dqkv,
                          float* delta_ws, int32_t B, int32_t S, int32_t H, int32_t hd, hct_stream_t stream);
/* fp32 variants of hct_patchify (cols fp32), hct_decoder_assemble(_bwd) (y / dy fp32) and hct_mae_loss_fwd/_bwd (pred /
 * dpred fp32); argument meaning as for the bf16 entry points. */
int hct_patchify_f32(const float* x, float* cols, const int64_t* patch_ids, int32_t* pos_idx_out, int32_t B, int32_t C,
                     int32_t H, int32_t W, int32_t D, int32_t p, int32_t rows_per_vol, hct_stream_t stream);
int hct_decoder_assemble_f32(const float* y, const int64_t* ids_restore, const float* mask_token, const float* dec_cls,
                             const float* dec_pos, float* out, int32_t N, int32_t L, int32_t keep, int32_t dim,
                             hct_stream_t stream);
int hct_decoder_assemble_bwd_f32(const float* dout, const int64_t* ids_restore, float* dy, float* dmask_token,
                                 float* ddec_cls, int32_t N, int32_t L, int32_t keep, int32_t dim, hct_stream_t stream);
int hct_mae_loss_fwd_f32(const float* pred, int32_t pred_prefix_rows, const float* imgs, const float* mask, float* loss_out,
                         int32_t N, int32_t C, int32_t H, int32_t W, int32_t D, int32_t p, int32_t norm_pix,
                         hct_stream_t stream);
int hct_mae_loss_bwd_f32(const float* pred, int32_t pred_prefix_rows, const float* imgs, const float* mask,
                         const float* dloss, const float* mask_sum, float* dpred, int32_t N, int32_t C, int32_t H,
                         int32_t W, int32_t D, int32_t p, int32_t norm_pix, hct_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Train-step glue: per-parameter clip (misc.py:374-383) + AdamW
 * (optimizers.py:354-360) as one multi-tensor launch each.
 * table: int64 [n, 7] = {param_ptr, grad_ptr, exp_avg_ptr, exp_avg_sq_ptr, numel, param_bf16_shadow_ptr or 0,
 * steps_behind}; the AdamW launch also refreshes the bf16 copy of every updated parameter that has one.
 * steps_behind = (step count of row 0) - (step count of this row): torch.optim.AdamW bias-corrects each parameter with
 * its own state['step'], and a tensor whose gradient was None for some steps (cancel_gradients_last_layer,
 * misc.py:366-371 / engine_pretrain_dino.py:95) lags the others; `step` below is row 0's count for this update.
 * norms_ws: fp32 [n] workspace receiving each gradient's SQUARED L2 norm.
 * ------------------------------------------------------------------------------------------- */
int hct_grad_norms_multi(const int64_t* table, int32_t n, float* norms_ws, hct_stream_t stream);
int hct_adamw_multi(const int64_t* table, int32_t n, const float* norms_ws, float clip, float lr,
                    float beta1, float beta2, float eps, float weight_decay, int32_t step,
                    hct_stream_t stream);
/* Same update with the per-step scalars read from DEVICE memory, so that the launch can be captured in a CUDA graph and
 * replayed while the schedule moves: hyper fp32 [5] = {lr, weight_decay, 1 - beta1^step, sqrt(1 - beta2^step), step}
 * (lr_sched.py:18-55 changes lr every iteration; the bias corrections change with the step count). */
int hct_adamw_multi_dev(const int64_t* table, int32_t n, const float* norms_ws, float clip, const float* hyper,
                        float beta1, float beta2, float eps, hct_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Downstream heads: LoRA q/v adapters, BatchNorm1d over token rows, attentive pooling.
 * ------------------------------------------------------------------------------------------- */
/* LoRA residuals of SelfAttention (src/models/attentionblock.py:57-59).  The reference reshapes -- it does not
 * permute -- lora_q(x) [B,N,C] to [B,H,N,hd] before adding it to q, so flat element f = (h*N + n)*hd + d of a
 * sample's LoRA output lands on q[b,h,n,d]; this call reproduces that on our qkv layout [B,N,3,H,hd] (bf16):
 *   backward == 0:  qkv[b,n,0,h,d] += lq[b].flat[f],  qkv[b,n,2,h,d] += lv[b].flat[f]      (in place)
 *   backward != 0:  lq[b].flat[f] = qkv[b,n,0,h,d],   lv[b].flat[f] = qkv[b,n,2,h,d]       (qkv = dqkv; the adjoint)
 * lq, lv: bf16 [B, N*C]. */
int hct_lora_shuffle(void* qkv, void* lq, void* lv, int64_t batch, int32_t seq, int32_t heads, int32_t head_dim,
                     int32_t backward, hct_stream_t stream);
/* nn.BatchNorm1d(dim, affine=False) applied to [B, dim, N] = per-column statistics of the token matrix
 * x fp32 [rows = B*N, dim] (src/models/classifier.py:64-65,89,96; also :18,31 for the linear probe).
 * stats (training): sums fp32 [2*dim] must be ZERO on entry and receives (sum x, sum x^2); mean / invstd fp32 [dim]
 *   get the batch mean and 1/sqrt(biased var + eps); running_mean / running_var (may be NULL) are updated in place
 *   with `momentum` and the unbiased variance, as torch does.
 * apply: y = (x - mean) * invstd, bf16 or fp32; with is_var != 0 the third argument is a (running) variance and
 *   invstd_scratch fp32 [dim] receives 1/sqrt(var + eps) first (eval mode).
 * bwd: dx fp32 = invstd * (dy - mean_r(dy) - x_hat * mean_r(dy * x_hat)); sums fp32 [2*dim] ZERO on entry (workspace);
 *   sums == NULL: eval-mode backward dx = invstd * dy. */
int hct_colnorm_stats(const float* x, float* sums, int64_t rows, int32_t dim, float eps, float momentum,
                      float* mean, float* invstd, float* running_mean, float* running_var, hct_stream_t stream);
int hct_colnorm_apply(const float* x, const float* mean, const float* invstd_or_var, int32_t is_var, float eps,
                      float* invstd_scratch, void* y, int32_t y_bf16, int64_t rows, int32_t dim, hct_stream_t stream);
int hct_colnorm_bwd(const void* dy, int32_t dy_bf16, const float* x, const float* mean, const float* invstd,
                    float* sums, float* dx, int64_t rows, int32_t dim, hct_stream_t stream);
/* Attentive pooling of AttentionClassifier.forward (classifier.py:84-94): num_queries learned query tokens
 * (cls fp32 [num_queries, H*hd], shared by the batch) attend over the N tokens of each sample.
 * kv bf16 [B, N, 2, H, hd] = the wkv Linear's natural output.  scale_total multiplies q: the reference scales q by
 * self.scale AND F.scaled_dot_product_attention applies 1/sqrt(hd) again, so callers pass self.scale / sqrt(hd).
 * out fp32 [B, num_queries, H*hd]; probs fp32 [B, H, num_queries, N] (saved for backward).
 * bwd: dout fp32 like out -> dkv bf16 like kv (fully written), dcls fp32 [num_queries, H*hd] ACCUMULATED (+=). */
int hct_pool_attention_fwd(const float* cls, const void* kv, float* out, float* probs, int32_t batch, int32_t seq,
                           int32_t heads, int32_t head_dim, int32_t num_queries, float scale_total, hct_stream_t stream);
int hct_pool_attention_bwd(const float* cls, const void* kv, const float* out, const float* probs, const float* dout,
                           float* dcls, void* dkv, int32_t batch, int32_t seq, int32_t heads, int32_t head_dim,
                           int32_t num_queries, float scale_total, hct_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* HCT_B200_H */
