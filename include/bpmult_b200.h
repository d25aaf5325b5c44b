/* bpmult_b200 -- C ABI of the B200-native BPMulT fusion trunk (sm_100a).
 *
 * The reference (Damorgal/Biprojection-Multimodal-Transformer) has no FFI / plugin interface: its hot path is
 * reached through the Python nn.Module API (bpmult/models/__init__.py:12-14 get_model -> train.py:311-321
 * model_forward).  This header is therefore the boundary a maintainer binds instead of the ATen calls the
 * reference makes on that path; each entry point names the reference op (file:line under /root/reference/bpmult)
 * it replaces.  INTEGRATION.md shows the ctypes stub that wires these into the reference's modules.
 *
 * Conventions
 *  - Plain pointers and sizes only.  All pointers are DEVICE pointers unless stated otherwise.
 *  - `stream` is a cudaStream_t passed as void*.  Calls never allocate, never synchronise and are capturable in
 *    CUDA graphs.  Every function returns 0 on success or a negative BPM_E* code; bpm_last_error() (host, per
 *    thread) describes the failure.  There is NO CPU fallback: a missing GPU / wrong arch is an error.
 *  - dtype codes: BPM_F32 = 0, BPM_BF16 = 1 ("storage type T" below).  Accumulation is always fp32.
 *  - Internal activation layout ("rows"): batch-major rows r = b*T + t, row pitch Dp = round_up(D, 64) elements,
 *    pad columns are ZERO.  Per-head tensors (q, k, v, attention output) use pitch H*dhp, dhp = round_up(dh, 32)
 *    (16 for dh <= 16), head h at columns [h*dhp, h*dhp + dh), pad columns zero.
 *  - Dropout: counter-based and stateless.  Element e of the padded row-major tensor a mask applies to (for attention:
 *    ((b*H + h)*T + i)*S + j) takes 16 bits of a multiply-xorshift-multiply hash of its pair counter e/2, keyed by (seed, site):
 *    keep(e) <=> half-word >= round(p * 65536); kept values are scaled by 1/(1-p) (exact definition: csrc/bpm_common.cuh).
 *    `seed_ptr` (device, may be NULL) overrides `seed` when non-NULL so a captured graph can be replayed with a new seed.
 *    Backward kernels regenerate masks from (seed, site); nothing is stored except the optional attention keep bits.
 *
 * Entry-point families named in SURVEY section 8b and where they live here (one forward and one backward entry per family):
 *    embed / projection     bpm_stage_rows + bpm_gemm (Conv1d k=1 as a row GEMM), bpm_embed_fwd / bpm_embed_bwd
 *    layernorm              bpm_layernorm_fwd / bpm_layernorm_bwd, bpm_layernorm_bwd_cast (fused operand emission); the residual add is the
 *                           producing GEMM's epilogue (bpm_gemm_t.residual)
 *    gemm fwd/dgrad/wgrad   bpm_gemm with ta / tb / accumulate / colsum_out and the epilogue fields of bpm_gemm_t
 *    xattn                  bpm_xattn_fwd / bpm_xattn_bwd, workspace query bpm_xattn_bwd_workspace
 *    seq GMU                bpm_gemm x 3 + bpm_gmu_fwd / bpm_gmu_bwd
 *    head GMU + BCE         bpm_pool_fwd / bpm_pool_bwd, bpm_tsgate_fwd / bpm_tsgate_bwd, bpm_bce_fwd_bwd
 *    optimizer              bpm_adam_step
 *    4-modality extras      bpm_timelin_fwd / bpm_timelin_bwd, bpm_conv1d_im2col / col2im, bpm_adaptive_pool_fwd / bwd
 *    parameter staging      bpm_pack_matrix / bpm_unpack_matrix, bpm_remap_batch / bpm_remap_units, bpm_ln_fold_*
 */
#ifndef BPMULT_B200_H
#define BPMULT_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BPM_F32 0
#define BPM_BF16 1

#define BPM_OK 0
#define BPM_EINVAL (-1)      /* bad argument / unsupported shape or alignment */
#define BPM_ELAUNCH (-2)     /* CUDA launch / driver error */
#define BPM_ENOGPU (-3)      /* no sm_100 device */

typedef struct {
  uint64_t seed;             /* used when seed_ptr == NULL */
  const uint64_t* seed_ptr;  /* device pointer to the seed (graph-replay safe), or NULL */
  uint64_t site;             /* unique id of the dropout site (selects the Philox stream) */
  float p;                   /* drop probability; 0 disables */
} bpm_dropout_t;

int bpm_version(void);
const char* bpm_last_error(void);
/* 1 when device `dev` is compute capability 10.x */
int bpm_device_ok(int dev);
/* diagnostic knobs for profiling scripts (slot 0: GEMM stage bypass bits, 1: attention); the product never sets them */
int bpm_debug_set(int slot, int value);
/* device buffer (uint64 [5][4096]) that receives a per-role event trace of CTA 0 of the attention backward kernel; NULL = off */
int bpm_debug_set_ptr(void* p);

/* ---- weight staging ------------------------------------------------------------------------------------------
 * Reference parameters stay fp32 in reference layout (state_dict names of SURVEY 8b); kernels consume zero-padded
 * copies in storage type T.  Optional head remap: row r = h*dh + j -> h*dhp + j (row_dh > 0), same for columns.
 * Replaces nothing in the reference (layout glue); the inverse accumulates padded fp32 gradients back. */
int bpm_pack_matrix(const float* src, int rows, int cols, int ld_src, void* dst, int rows_p, int cols_p, int dst_dtype,
                    int row_dh, int row_dhp, int col_dh, int col_dhp, void* stream);
int bpm_unpack_matrix(const float* src_p, int rows_p, int cols_p, float* dst, int rows, int cols, int ld_dst,
                      int row_dh, int row_dhp, int col_dh, int col_dhp, int accumulate, float scale, void* stream);

/* batched form: one launch for a table of remaps held in DEVICE memory (e.g. all ~115 parameters of an encoder).
 * mode 0 = pack (src fp32 reference layout -> dst padded, dtype dst_dtype), mode 1 = unpack (src padded fp32 -> dst fp32 reference
 * layout, (+)= scale * src). */
typedef struct {
  const void* src; void* dst;
  int32_t rows, cols, ld_src, ld_dst, rows_p, cols_p, row_dh, row_dhp, col_dh, col_dhp, dst_dtype, accumulate;
  float scale; int32_t pad_;
} bpm_remap_desc_t;
int bpm_remap_batch(const bpm_remap_desc_t* descs_dev, int n, int mode, void* stream);
/* the same with an explicit work list: unit = rows [row0, row0 + nrows) of descriptor `desc`, counted in DESTINATION rows (padded rows
 * when packing, reference rows when unpacking); one thread block per unit.  The caller sizes units by element count (~16 K per unit)
 * so that tensors of very different sizes share the machine evenly. */
typedef struct { int32_t desc, row0, nrows, pad_; } bpm_remap_unit_t;
int bpm_remap_units(const bpm_remap_desc_t* descs_dev, const bpm_remap_unit_t* units_dev, int n_units, int mode, void* stream);

/* ---- input staging: models/mmtr.py:741-761 (transpose, embed dropout on text, zero-pad time to n_vec) ----------
 * src fp32 element (b, t, c) at src[b*sb + t*st + c*sc]; dst T rows b*Tp + t, pitch Cp, zero for t >= T or c >= C. */
int bpm_stage_rows(const float* src, int B, int T, int C, int64_t sb, int64_t st, int64_t sc, void* dst, int Tp, int Cp,
                   int dst_dtype, bpm_dropout_t drop, void* stream);
/* inverse gather for input gradients: dsrc(b,t,c) (+)= mask*g[b*Tp+t, c] */
int bpm_unstage_rows(const float* g, int B, int T, int C, int Tp, int Cp, float* dsrc, int64_t sb, int64_t st, int64_t sc,
                     int accumulate, bpm_dropout_t drop, void* stream);

/* ---- embed: models/transformer.py:66-79 + models/position_embedding.py:8-76 -----------------------------------
 * y[r, c] = dropout(scale * x[r, c] + pe[pos(r), c]);  pos = t+1 if x[r, 0] != 0 else 0;  pe is fp32 [T+1, Dp]
 * (row 0 zero).  x is T [B*T, Dp]; y has dtype y_dtype.  bwd: dx[r, c] (+)= scale * mask * dy[r, c] (fp32). */
int bpm_embed_fwd(const void* x, int x_dtype, const float* pe, int B, int T, int D, int Dp, float scale, void* y, int y_dtype,
                  bpm_dropout_t drop, void* stream);
int bpm_embed_bwd(const float* dy, int rows, int D, int Dp, float scale, float* dx, int accumulate, bpm_dropout_t drop,
                  void* stream);

/* ---- LayerNorm: models/transformer.py:197-202,227-229 (nn.LayerNorm, eps 1e-5, statistics over the D real columns)
 * fwd writes zeros to pad columns and saves mean / rstd (fp32 [rows]).
 * bwd: dx (fp32) = [accumulate ? dx : 0] + LN'(dy);  dgamma / dbeta (fp32 [Dp]) are accumulated atomically. */
int bpm_layernorm_fwd(const void* x, int x_dtype, const float* gamma, const float* beta, int rows, int D, int Dp, float eps,
                      void* y, int y_dtype, float* mean, float* rstd, void* stream);
int bpm_layernorm_bwd(const void* dy, int dy_dtype, const void* x, int x_dtype, const float* mean, const float* rstd,
                      const float* gamma, int rows, int D, int Dp, float* dx, int accumulate, float* dgamma, float* dbeta,
                      void* stream);
/* same, and additionally cast_out (cast_dtype, pitch Dp, may be NULL) = dropmask(cast_drop) * dx_new: the GEMM operand of the next
 * backward block (what bpm_cast_drop would produce from dx in a separate pass) */
int bpm_layernorm_bwd_cast(const void* dy, int dy_dtype, const void* x, int x_dtype, const float* mean, const float* rstd,
                           const float* gamma, int rows, int D, int Dp, float* dx, int accumulate, float* dgamma, float* dbeta,
                           void* cast_out, int cast_dtype, bpm_dropout_t cast_drop, void* stream);

/* ---- LayerNorm affine folded into the K / V projection of a crossmodal layer ---------------------------------------
 * transformer.py:83-85 feeds the same x_k / x_v to all L layers; only layer_norms[.] and in_proj differ.  x_hat is computed
 * once per encoder (bpm_layernorm_fwd with a unit affine) and K = x_hat (W diag(gamma))^T + (bias + W beta).
 * fwd: Wp[map(i), j] = W[i,j]*gamma[j] (dtype wp_dtype, pitch ldp), bp[map(i)] = bias[i] + sum_j W[i,j]*beta[j];
 *      map = head remap of rows (row_dh -> row_dhp, 0 = identity); W, bias, gamma, beta are fp32 in reference layout.
 * bwd: gWf / gbf are the padded fp32 gradient accumulators of (Wp, bp); gW += gWf*gamma + gbf (x) beta, gb += gbf (padded
 *      accumulators of W / bias), dgamma[j] += sum_i gWf*W, dbeta[j] += sum_i gbf[i]*W[i,j]. */
int bpm_ln_fold_fwd(const float* W, int ldw, const float* bias, const float* gamma, const float* beta, int rows, int cols,
                    int row_dh, int row_dhp, void* Wp, int wp_dtype, int ldp, float* bp, void* stream);
int bpm_ln_fold_bwd(const float* W, int ldw, const float* gamma, const float* beta, int rows, int cols, int row_dh, int row_dhp,
                    const float* gWf, int ldf, const float* gbf, float* gW, int ldg, float* gb, float* dgamma, float* dbeta,
                    void* stream);

/* batched form: a device table of problems (mode 0 = fwd uses W, bias, gamma, beta -> Wp, bp; mode 1 = bwd uses W, gamma, beta, gWf, gbf
 * -> gW, gb, dgamma, dbeta), one launch; max_rows = largest `rows` in the table */
typedef struct {
  const float* W; const float* bias; const float* gamma; const float* beta; void* Wp; float* bp;
  const float* gWf; const float* gbf; float* gW; float* gb; float* dgamma; float* dbeta;
  int32_t rows, cols, ldw, ldp, ldf, ldg, row_dh, row_dhp, wp_dtype, pad_;
} bpm_fold_desc_t;
int bpm_ln_fold_batch(const bpm_fold_desc_t* descs_dev, int n, int max_rows, int mode, void* stream);

/* ---- GEMM with fused epilogue: F.linear (multihead_attention.py:152-158), fc1/fc2 (transformer.py:186-190),
 *      out_proj (multihead_attention.py:130), Conv1d k=1 (mmtr.py:748-750), GMU linears (mmtr.py:190-194) ---------
 * C[M,N] = epi( op(A)[M,K] * op(B)[K,N] ).  ta = 0: A stored [M,K] (pitch lda); ta = 1: A stored [K,M].
 * tb = 0: B stored [N,K] (nn.Linear weight layout); tb = 1: B stored [K,N].
 * epi(v)(m,n): v = (v + bias[n]) * alpha; relu if act==1; dropout; if gate: v = gate[m,n] > 0 ? v*gate_scale : 0;
 *              v += residual[m,n]; accumulate ? C += v : C = v.
 * bf16 inputs run on tcgen05 tensor cores (TMA -> smem -> tcgen05.mma -> TMEM -> tcgen05.ld epilogue) and need
 * 16-byte aligned pointers and pitches; fp32 inputs run an FFMA kernel (exact-fp32 "precision mode"). */
typedef struct {
  int ab_dtype, ta, tb, M, N, K;
  const void* A; int lda;
  const void* B; int ldb;
  void* C; int ldc; int c_dtype;
  const float* bias; float alpha; int act;
  bpm_dropout_t drop;
  const void* gate; int ldg; int gate_dtype; float gate_scale;
  const void* residual; int ldr; int res_dtype;
  int accumulate;            /* C must be fp32; split-K partial sums are added atomically */
  int split_k;               /* 0 = auto */
  float* colsum_out;         /* optional, ta = 1 only: colsum_out[m] += sum_k op(A)[m, k]  (= the bias gradient when this GEMM is a
                                weight gradient dW = dY^T X); fused into the MMA main loop, dY is not re-read */
} bpm_gemm_t;
int bpm_gemm(const bpm_gemm_t* g, void* stream);

/* column sums for bias gradients: out[n] (+)= sum_m X[m, n]  (fp32 atomics) */
int bpm_colsum(const void* X, int dtype, int M, int N, int ld, float* out, void* stream);

/* ---- crossmodal attention: models/multihead_attention.py:95-127 + mask models/transformer.py:209-216 ----------
 * q [B,T,H*dhp] (already scaled by dh^-0.5), k, v [B,S,H*dhp] in T; out [B,T,H*dhp]; lse fp32 [B,H,T].
 * mask_off >= 0: key j visible to query i iff j <= i + mask_off (reference: mask_off = |S - T|); mask_off < 0: no mask.
 * key_pad (uint8 [B,S], 1 = padded key, may be NULL): superset feature, default off (the reference has none,
 * multihead_attention.py:52); a padded key gets probability 0 and zero dk / dv.  Every query must keep one visible key.
 * bwd writes dq * dq_scale, dk, dv (T) and needs a workspace `delta` of bpm_xattn_bwd_workspace(a) floats, 128-byte aligned: on return
 * [0 .. B*H*T) = rowsum(dO*O), [B*H*T .. 2*B*H*T) = lse*log2e; the tensor-core kernels for head dim 128 keep their fp32 dQ accumulator
 * (TMA reduce-add target, [B, T, H*dhp]) behind that.
 * Kernels: dhp = 32 -> attn_tc.cu, dhp = 64 (forward) / 128 -> attn_tc128.cu (tcgen05 + TMEM + TMA); fp32 storage or any other head
 * dim -> exact-fp32 kernels (attn_simt.cu). */
typedef struct {
  int dtype, B, T, S, H, dh, dhp, mask_off;
  const uint8_t* key_pad;
  bpm_dropout_t drop;
  uint32_t* drop_bits;  /* optional scratch [B*H, T, ceil(S/32)]: keep bits written by the forward (dropout on) and read by the
                           backward instead of regenerating the Philox stream; NULL = regenerate */
  int ld_kv;            /* row pitch (elements) of k and v; 0 = H*dhp.  Lets k / v be column slices of an all-layers projection */
  int ld_dkv;           /* row pitch (elements) of dk and dv; 0 = H*dhp */
} bpm_attn_t;
int bpm_xattn_fwd(const bpm_attn_t* a, const void* q, const void* k, const void* v, void* out, float* lse, void* stream);
int bpm_xattn_bwd(const bpm_attn_t* a, const void* q, const void* k, const void* v, const void* out, const void* dout,
                  const float* lse, float* delta, void* dq, float dq_scale, void* dk, void* dv, void* stream);
/* number of floats bpm_xattn_bwd needs at `delta` for this problem (always >= 2*B*H*T) */
int64_t bpm_xattn_bwd_workspace(const bpm_attn_t* a);
/* head-averaged probabilities (multihead_attention.py:133-135), only on request: w fp32 [B,T,S] */
int bpm_xattn_weights(const bpm_attn_t* a, const void* q, const void* k, const float* lse, float* w, void* stream);

/* ---- sequence GMU: models/mmtr.py:179-195 (GatedMultimodalLayerFeatures), :161-177 (GatedMultimodalLayer) ------
 * pre-activations come from bpm_gemm; this fuses tanh / sigmoid / gating (+ optional addend, mmtr.py:806).
 * features = 1: y = z*h1*a1 + (1-z)*h2*a2 (+ add);  features = 0: y = z*h1 + (1-z)*h2.
 * bwd: dh1pre, dh2pre, dzpre (T) and the direct input grads da1 / da2 (fp32, accumulated). */
int bpm_gmu_fwd(int dtype, int features, const void* a1, const void* a2, const void* h1pre, const void* h2pre,
                const void* zpre, const void* addend, int rows, int Dp, void* y, void* z_out, void* stream);
int bpm_gmu_bwd(int dtype, int features, const void* a1, const void* a2, const void* h1pre, const void* h2pre,
                const void* zpre, const float* dy, int rows, int Dp, void* dh1pre, void* dh2pre, void* dzpre,
                float* da1, float* da2, void* stream);

/* ---- elementwise helpers ------------------------------------------------------------------------------------- */
/* y = a + b (T), mmtr.py:799-800 */
int bpm_add(int dtype, const void* a, const void* b, void* y, int64_t n, void* stream);
/* dst(fp32) (+)= src (T or fp32) */
int bpm_axpy_f32(const void* src, int src_dtype, float* dst, int64_t n, int accumulate, void* stream);
/* y (T) = dropmask * x (fp32): fp32 gradient -> GEMM operand, regenerating a residual-dropout mask */
int bpm_cast_drop(const float* x, void* y, int y_dtype, int rows, int cols, bpm_dropout_t drop, void* stream);
/* pooling mmtr.py:808: out[b, col_off + c] = x[b*T + 0, c] + x[b*T + T-1, c] (fp32 out, pitch ld_out); bwd scatters */
int bpm_pool_fwd(const void* x, int dtype, int B, int T, int Dp, float* out, int ld_out, int col_off, void* stream);
int bpm_pool_bwd(const float* dout, int ld_out, int col_off, int B, int T, int Dp, float* dx, void* stream);

/* ---- final GMU (TextShifting3/4Layer, mmtr.py:197-247): fused = sum_i sigmoid(zpre_i) * tanh(hpre_i) -----------
 * hpre, zpre: fp32 [n_in][B, Dp] contiguous blocks; z_out fp32 [B, n_in*Dp] (gates, mmtr.py:863-866) */
int bpm_tsgate_fwd(const float* hpre, const float* zpre, int n_in, int B, int Dp, float* fused, float* z_out, void* stream);
int bpm_tsgate_bwd(const float* hpre, const float* zpre, const float* dfused, int n_in, int B, int Dp, float* dhpre,
                   float* dzpre, void* stream);

/* ---- time-axis Linear of the 4-modality model: mmtr.py:507-508,530,553  transfm_x2y(h.permute(2,1,0)).permute(2,1,0) ------
 * y[b, t2, d] = bias[t2] + sum_t W[t2, t] x[b, t, d] on batch-major [B, T, ld] rows (W = the nn.Linear weight [Tout, Tin], fp32;
 * columns D..ld are written as zeros).  Backward: dy fp32 -> dx fp32 (+= when accumulate_dx), dW += , db += (NULL pointers skip;
 * dW == NULL with db != NULL computes the bias gradient alone).  These are the exact-fp32 kernels; with bf16 storage the engine runs the
 * three products as per-sample bpm_gemm calls on tensor cores (y_b = W x_b, dW += dy_b x_b^T, dx_b += W^T dy_b) and takes only db here. */
int bpm_timelin_fwd(int dtype, const void* x, const float* W, const float* bias, void* y, int B, int Tin, int Tout, int D, int ld,
                    void* stream);
int bpm_timelin_bwd(int x_dtype, const float* dy, const void* x, const float* W, float* dx, int accumulate_dx, float* dW, float* db,
                    int B, int Tin, int Tout, int D, int ld, void* stream);

/* ---- AudioEncoder of the 4-modality model: mmtr.py:93-108  Conv1d(C, C, k=KW, stride) x 2 + AdaptiveAvgPool1d(Tp) --------------------
 * The strided convolution is an implicit GEMM (K = KW*C, run by bpm_gemm on tensor cores); these are the layout kernels around it,
 * on time-major rows [B*T, C] (C % 8 == 0):
 *   im2col   col[(b*Tout + t), k*C + c] = x[(b*Tin + stride*t + k), c],  Tout = (Tin - KW)/stride + 1  (col has x's dtype, pitch KW*C)
 *   col2im   dx[(b*Tin + u), c] = sum_{k: (u-k) % stride == 0} dcol[(b*Tout + (u-k)/stride), k*C + c]  (fp32 out, a gather: no atomics)
 *   weights  Wp[co, k*Cin + ci] = W[co, ci, k] (nn.Conv1d layout, fp32 -> dst dtype) and the inverse for the weight gradient
 *   pooling  y[(b*Tp + i), c] = mean_{t in [floor(i*T/Tp), ceil((i+1)*T/Tp))} x[(b*T + t), c]  (torch AdaptiveAvgPool1d) and its backward */
int bpm_conv1d_im2col(const void* x, int dtype, int B, int Tin, int C, int ldx, int KW, int stride, void* col, int Tout, void* stream);
int bpm_conv1d_col2im(const void* dcol, int dtype, int B, int Tin, int C, int KW, int stride, int Tout, float* dx, int lddx, void* stream);
int bpm_conv1d_pack_weight(const float* W, int Cout, int Cin, int KW, void* Wp, int dst_dtype, void* stream);
int bpm_conv1d_unpack_wgrad(const float* gWp, int Cout, int Cin, int KW, float* gW, int accumulate, void* stream);
int bpm_adaptive_pool_fwd(const void* x, int x_dtype, int B, int T, int C, int ldx, int Tp, void* y, int y_dtype, int ldy, void* stream);
int bpm_adaptive_pool_bwd(const float* dy, int B, int T, int C, int Tp, int lddy, float* dx, int lddx, void* stream);

/* ---- loss: train.py:99-106,333 nn.BCEWithLogitsLoss(pos_weight), mean over (B, C) -----------------------------
 * loss (fp32 scalar, overwritten) and dlogits = dloss/dlogits * grad_scale (fp32 [B, ldl]) in one launch. */
int bpm_bce_fwd_bwd(const float* logits, int ldl, const float* targets, const float* pos_weight, int B, int C,
                    float grad_scale, float* loss, float* dlogits, void* stream);

/* ---- optimiser: train.py:123-125 optim.Adam (default betas / eps, no weight decay) ----------------------------
 * step_ptr: device int64 step counter (already incremented); grad_scale multiplies the gradient (1/world, 1/accum).
 * lr_ptr (nullable): device fp32 learning rate that overrides `lr` -- lets a ReduceLROnPlateau scheduler (train.py:128-136,
 * 408) change the rate of a step that has been captured into a CUDA graph. */
int bpm_adam_step(float* param, const float* grad, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                  float eps, float grad_scale, const int64_t* step_ptr, const float* lr_ptr, void* stream);

#ifdef __cplusplus
}
#endif
#endif
