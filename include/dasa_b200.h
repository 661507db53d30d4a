/*
 * dasa_b200 — C ABI of the B200-native kernels behind DASA's agent_dg navigation-policy hot path.
 *
 * The reference (sunqiang85/DASA) has no FFI layer: its boundary for this path is the torch.nn.Module surface
 * that r2r_src/agent_dg.py touches. The Python drop-in modules in dasa_b200/modules.py keep
 * that surface; underneath they call this library through ctypes with raw device pointers. Every entry point
 * below names the reference code it replaces (paths relative to r2r_src/).
 *
 * Conventions
 *   - all tensors are device pointers, fp32 row-major unless stated; `ld*` are leading dimensions in ELEMENTS
 *     so strided slices (e.g. feature[..., :2048] of a [.., 2176] row) are read/written in place;
 *   - caller owns every buffer (incl. workspaces); kernels never allocate, free or retain pointers;
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); no host synchronisation inside, so
 *     every call is CUDA-graph capturable;
 *   - return value: 0 = ok, negative = DASA_ERR_* below; asynchronous CUDA faults surface at the next sync;
 *   - dropout is always an explicit, caller-provided keep mask (uint8, 1 = keep) plus the scale 1/(1-p);
 *     a NULL mask means "no dropout" (eval mode).
 */
#ifndef DASA_B200_H
#define DASA_B200_H

#include <stddef.h>
#include <stdint.h>

typedef unsigned short dasa_half_t;                   /* IEEE binary16 bits */

#ifdef __cplusplus
extern "C" {
#endif

#define DASA_OK 0
#define DASA_ERR_BAD_SHAPE (-1)
#define DASA_ERR_BAD_ALIGN (-2)
#define DASA_ERR_WORKSPACE (-3)
#define DASA_ERR_CUDA (-4)
#define DASA_ERR_UNSUPPORTED (-5)

/* library / build identification */
int dasa_version(void);                 /* 100*major + minor */
const char* dasa_build_arch(void);      /* "sm_100a" */
const char* dasa_last_error(void);      /* text of the last CUDA error seen by a launch wrapper */

/* ------------------------------------------------------------------------------------------------ GEMM
 * C[M,N] = epilogue( alpha * opA(A)[M,K] * opB(B)[K,N] + beta * C ), row-major C with leading dim ldc.
 *   a_kmajor=1: A stored [M,K] (k contiguous, lda >= K);  a_kmajor=0: A stored [K,M] (m contiguous, lda >= M)
 *   b_kmajor=1: B stored [N,K] (k contiguous, ldb >= K) — the nn.Linear weight layout;  b_kmajor=0: B stored [K,N]
 * epilogue: DASA_EPI_* applied after alpha/beta; `bias` is [N] (may be NULL); for DASA_EPI_GATE the result is
 *   sigmoid(acc + bias) * gate_src[m*ld_gate + n], optionally multiplied by drop_mask*drop_scale, and the sigmoid
 *   itself is stored to `gate_out` (ld = ld_gate_out) when non-NULL (needed by the backward pass).
 * precision: DASA_PREC_FP32 = FFMA (exact fp32 products); DASA_PREC_TF32 = tcgen05 kind::tf32 tensor cores
 *   (fp32 storage, 10-bit mantissa products, fp32 accumulate in TMEM) — falls back to FP32 when the operands do
 *   not satisfy the TMA alignment rules (16-byte base/stride). Never a CPU fallback.
 * workspace: split-K partial sums; size from dasa_gemm_workspace_bytes (may be 0).
 * Replaces nn.Linear / torch.bmm / torch.matmul call sites: agent_dg.py:1538 (a_fc), model.py:277,315 (linear_in),
 * model.py:292 (linear_out), model.py:501, 514 (LSTMCell gates), vilmodel.py:210-212,233,247,293,306 (BERT), 1088.
 */
enum { DASA_EPI_NONE = 0, DASA_EPI_BIAS = 1, DASA_EPI_BIAS_TANH = 2, DASA_EPI_BIAS_GELU = 3, DASA_EPI_BIAS_RELU = 4,
       DASA_EPI_GATE = 5, DASA_EPI_TANH = 6 };
enum { DASA_PREC_FP32 = 0, DASA_PREC_TF32 = 1 };

typedef struct {
  const float* bias;          /* [N] or NULL */
  const float* gate_src;      /* EPI_GATE: f, [M, ld_gate] */
  int64_t ld_gate;
  float* gate_out;            /* EPI_GATE: sigmoid(acc+bias) saved here when non-NULL, [M, ld_gate_out] */
  int64_t ld_gate_out;
  const uint8_t* drop_mask;   /* optional keep mask [M, N] contiguous, applied to the final value */
  float drop_scale;
} dasa_epilogue_t;

/* Debug / experiment switch: route many-wave 128x128 TF32 GEMMs through the cluster-of-2 TMA-multicast variant (each CTA
 * loads half of the shared A tile and multicasts it). Off by default: measured no gain on B200 (see gemm_tc.cu). */
int dasa_debug_gemm_multicast(int on);
/* Routing of K-major TF32 GEMMs to the persistent CTA-pair kernel (tcgen05 cta_group::2, 256 x 256 tiles, gemm_tc2.cu):
 * 0 = never, 1 = when the tile count fills the 74 TPCs (default; also env DASA_TC_PAIR), 2 = always (tests). */
int dasa_debug_gemm_pair(int mode);
/* Routing of K-major TF32 GEMMs with M <= 32 rows (the decoder's per-action projections) to the weight-streaming mma.sync kernel
 * (gemm_skinny.cu: K slices reduced through shared memory and a thread-block cluster's DSMEM): 0 = off, 1 = on for the shapes where it wins (default; env
 * DASA_SKINNY), 2 = every eligible shape (tests). */
int dasa_debug_gemm_skinny(int on);

/* Which kernel family took each dasa_gemm call since the last reset (host-side counters, one slot per DASA_ROUTE_*):
 * tests assert with them that the benchmarked configuration really runs on the tcgen05 kernels and that no TF32-mode GEMM
 * silently fell back to the FFMA kernel because of a misaligned operand (that case also prints one warning to stderr).
 * out[i] = count of route i for i < min(n, DASA_ROUTE_COUNT); reset != 0 zeroes the counters afterwards.               */
enum { DASA_ROUTE_SKINNY = 0,          /* gemm_skinny.cu: M <= 32 weight-streaming mma.sync TF32                        */
       DASA_ROUTE_PAIR = 1,            /* gemm_tc2.cu: persistent CTA-pair tcgen05 kernel, K-major operands             */
       DASA_ROUTE_PAIR_MN = 2,         /* gemm_tc2.cu: MN-major operand(s) (dX = dY.W, dW = dY^T.X)                     */
       DASA_ROUTE_PAIR_MN_SPLITK = 3,  /* the same, K split + deterministic fold                                        */
       DASA_ROUTE_PAIR_GROUPED = 4,    /* two problems per launch (bi-LSTM recurrence, both directions)                 */
       DASA_ROUTE_TC_SINGLE = 5,       /* gemm_tc.cu: single-CTA tcgen05 kernel (few-tile problems, split-K)            */
       DASA_ROUTE_SIMT_TF32_MODE = 6,  /* FFMA kernel under DASA_PREC_TF32 for a layout / size the tensor kernels do not take (e.g. K < 32) */
       DASA_ROUTE_SIMT_MISALIGNED = 7, /* FFMA kernel under DASA_PREC_TF32 ONLY because of operand alignment: a performance bug */
       DASA_ROUTE_SIMT_FP32 = 8,       /* FFMA kernel, DASA_PREC_FP32 requested                                          */
       DASA_ROUTE_PAIR_F16 = 9,        /* gemm_tc2.cu: fp16 operands, tcgen05 kind::f16 (dasa_gemm_f16)                 */
       DASA_ROUTE_COUNT = 10 };
int dasa_debug_gemm_route_counts(int64_t* out, int n, int reset);

/* C[M, N] = epilogue(A[M, K] B[N, K]^T (+ bias)) with IEEE fp16 operands (both K-major, leading dimensions in elements, multiples of
 * 8) on the persistent CTA-pair tcgen05 kernel with kind::f16 (twice the TF32 rate), fp32 accumulation. c_half = 0: C is float
 * [M, ldc]; 1: C is fp16 [M, ldc] (N % 4 == 0) - the activation of the next fp16 GEMM. epilogue: DASA_EPI_NONE / _BIAS / _BIAS_GELU,
 * or DASA_EPI_GATE with fp32 output (the DGAdaChannel gate over an fp16 copy of the depth features; gate operand, saved gate and
 * keep mask as in dasa_gemm).
 * Used for the forward-only GEMMs of the frozen transformer stack (vilmodel.py:1370-1410: 9 language + 3 cross-modal layers,
 * detached in the train configuration): fp16 keeps TF32's 10 mantissa bits, so inside fp16's normal range the products are the
 * TF32 kernel's. dasa_gemm_f16_supported: 1 when the shape has enough 256 x 256 tiles for this kernel (else use dasa_gemm).     */
int dasa_gemm_f16_supported(int M, int N, int K);
int dasa_gemm_f16(int M, int N, int K, const dasa_half_t* A, int64_t lda, const dasa_half_t* B, int64_t ldb, void* C, int64_t ldc,
                  int c_half, int epilogue, const dasa_epilogue_t* epi, void* stream);
/* C[M, N] = alpha * A^T B + beta * C with A stored [K][lda] and B stored [K][ldb] as IEEE fp16 (both MN-major: the rows of both
 * arrays run over the reduction index - the shape of a weight gradient dW = dY^T X), tcgen05 kind::f16, fp32 accumulate, K split
 * through the workspace (dasa_gemm_workspace_bytes(M, N, K, DASA_PREC_TF32) bytes) with a deterministic fold. lda / ldb in
 * elements, multiples of 8; K >= 64. The packed bi-LSTM's dW_ih / dW_hh (r2rmodel.py:2339-2357 backward) use it with the fp16
 * operand copies its recurrence keeps (alpha = 2^-8 undoes the gradient copy's scale).                                        */
int dasa_gemm_f16_mn(int M, int N, int K, float alpha, const dasa_half_t* A, int64_t lda, const dasa_half_t* B, int64_t ldb,
                     float beta, float* C, int64_t ldc, void* workspace, size_t workspace_bytes, void* stream);
size_t dasa_gemm_workspace_bytes(int M, int N, int K, int precision);
/* 1 when dasa_gemm(DASA_PREC_TF32) runs this operand-layout combination on the tensor cores directly: always for two K-major
 * operands; for an MN-major A ([K][M] in memory) and / or B ([K][N]) when the problem is large enough for the persistent CTA-pair
 * kernel, which feeds tcgen05 MN-major TF32 operands (TMA boxes of 32 columns x 32 k rows, 128B swizzle) — the backward GEMMs
 * dX = dY.W and dW += dY^T.X then need no transposed copy. Otherwise such layouts run on the FFMA kernel.                */
int dasa_gemm_layout_on_tensor_cores(int a_kmajor, int b_kmajor, int M, int N, int K);
int dasa_gemm(int a_kmajor, int b_kmajor, int M, int N, int K, float alpha, const float* A, int64_t lda,
              const float* B, int64_t ldb, float beta, float* C, int64_t ldc, int epilogue,
              const dasa_epilogue_t* epi, int precision, void* workspace, size_t workspace_bytes, void* stream);

/* column sums: out[n] (+)= sum_m X[m*ldx + n]   (bias gradients) */
int dasa_colsum(const float* X, int64_t ldx, int M, int N, float* out, int accumulate, void* stream);
/* the same over an fp16 matrix, out[n] (+)= scale * sum_m X[m, n]: bias gradients from the scaled fp16 gradient copies the
 * fp16-operand weight-gradient path keeps (scale = the inverse of the copy's factor).                                  */
int dasa_colsum_h(const dasa_half_t* X, int64_t ldx, int M, int N, float scale, float* out, int accumulate, void* stream);

/* ------------------------------------------------------------------------------------------ AdaIN family
 * a1: DGAdaChannel ab_type=a, a_type=sigmoid (agent_dg.py:1534-1547) = dasa_gemm(..., DASA_EPI_GATE) for the fused
 *     path, or this elementwise epilogue when the pre-activation g is already in memory:
 *     out[r, c] = sigmoid(g[r, c]) * f[r, c] (* mask * scale);  r < R, c < C.                                   */
int dasa_gate_modulate(const float* g, int64_t ldg, const float* f, int64_t ldf, float* out, int64_t ldo,
                       const uint8_t* drop_mask, float drop_scale, int R, int C, void* stream);
/* backward of the gate: dg[r,c] = dout[r,c] * mask*scale * f[r,c] * s(1-s), s = saved sigmoid (Appendix A, K1) */
int dasa_gate_backward(const float* dout, int64_t lddo, const float* f, int64_t ldf, const float* s, int64_t lds,
                       const uint8_t* drop_mask, float drop_scale, float* dg, int64_t lddg, int R, int C, void* stream);
/* the same, also writing dg16[r, c] = fp16(dg[r, c] * scale16) (saturating; [R, C] contiguous, C % 4 == 0): the dY operand of
 * the gate's weight gradient on dasa_gemm_f16_mn. dg may be NULL (only the fp16 copy is written).                      */
int dasa_gate_backward_h(const float* dout, int64_t lddo, const float* f, int64_t ldf, const float* s, int64_t lds,
                         const uint8_t* drop_mask, float drop_scale, float* dg, int64_t lddg, dasa_half_t* dg16, float scale16,
                         int R, int C, void* stream);

/* a2: per-channel statistics over the views of one panorama (agent_dg.py:1651-1656):
 *     stats[n, 0:C]=mean, [C:2C]=unbiased std, [2C:3C]=max, [3C:4C]=min of d[n, :, c] over V views.            */
int dasa_view_stats(const float* d, int64_t ld_row, int64_t ld_sample, int N, int V, int C, float* stats, void* stream);
/* a2: out[n,v,c] = a[n,c] * f[n,v,c] + b[n,c]  (agent_dg.py:1636, 1661); b may be NULL                          */
int dasa_channel_modulate(const float* f, int64_t ldf_row, int64_t ldf_sample, const float* a, const float* b,
                          float* out, int64_t ldo_row, int64_t ldo_sample, int N, int V, int C, void* stream);
/* a2 backward of channel_modulate (autograd of agent_dg.py:1636, 1661): da[n,c] = sum_v dout*f, db[n,c] = sum_v dout
 *     (db may be NULL), df[n,v,c] = a[n,c]*dout (df may be NULL: f is environment data in the agent). One pass.    */
int dasa_channel_modulate_bwd(const float* dout, int64_t ldo_row, int64_t ldo_sample, const float* f, int64_t ldf_row,
                              int64_t ldf_sample, const float* a, float* da, float* db, float* df, int64_t lddf_row,
                              int64_t lddf_sample, int N, int V, int C, void* stream);
/* a2: model.adaptive_instance_normalization (model.py:1822-1840): per (n, view) row statistics over the C channels,
 *     unbiased variance + eps; out = (f - mu_f) / sd_f * sd_d + mu_d. One pass: read f, read d, write out.      */
int dasa_adain_rows(const float* f, int64_t ldf, const float* d, int64_t ldd, float* out, int64_t ldo,
                    int R, int C, float eps, void* stream);

/* ---------------------------------------------------------------------------------- decoder attention (a3-a5)
 * Single-pass "row attention": for each sample b, rows r < nrows[b] (or `rows` if nrows==NULL) of ctx[b] (row
 * stride ld_row, sample stride ld_sample, D channels) are staged ONCE in shared memory by bulk-async (TMA) copies,
 * split across the CTAs of a thread-block cluster by channel; partial logits z_r = ctx_r . t are reduced across the
 * cluster through distributed shared memory, then
 *   shift_k == 0 : alpha = softmax(z masked to -inf where mask[b,r] != 0)            (model.py:276-288, SoftDotAttention)
 *   shift_k  > 0 : p = softmax(z); kappa = softmax(W_s h + b_s);
 *                  q[e,l] = sum_j kappa_j p[e,(l+j-k/2) mod headings]                 (model.py:327-345, ShiftSoftDotAttention)
 * and wc = sum_r w_r ctx_r is produced from the same shared-memory tile. attn_out receives the PRE-shift softmax
 * (model.py:336,351-353); q_out (optional) the shifted weights (needed by backward).
 * wc is written with leading dim ld_wc so it can land directly in the [wc ; h] concat buffer of linear_out.
 */
int dasa_row_attention_fwd(const float* ctx, int64_t ld_row, int64_t ld_sample, int B, int rows, int D,
                           const float* t, int64_t ld_t, const uint8_t* mask, int64_t ld_mask,
                           int shift_k, int headings, const float* kappa_logits, int64_t ld_kappa,
                           float* wc, int64_t ld_wc, float* attn_out, float* q_out, float* kappa_out, void* stream);
/* Debug hook (scripts/ only): when buf != NULL the large-batch pipelined forward writes SM-clock stamps of its hand-offs
 * (producer issue, dots start/end, softmax start/end, weighted-sum start/end) to buf[CTA][ceil(B/clusters)][8] (int64).
 * Pass NULL to switch it off. Not used by the product path. */
int dasa_debug_row_attention_trace(void* buf);

/* Fused K1 -> K3: DGAdaChannel's gate epilogue (agent_dg.py:1544-1547: df = sigmoid(g) * f on the first
 * gate_C channels, given the pre-activations g = a_fc(d); the remaining D - gate_C angle channels pass through; chan_scale
 * (optional, [gate_C]) = the drop_env noise shared by batch and views, agent_dg.py:656) feeding ShiftSoftDotAttention
 * (model.py:327-345) WITHOUT materialising df_t: the raw feature slice is staged in shared memory by TMA, modulated in place by
 * the gate streamed from HBM, and the attention (logits, softmax over the views, circular heading shift, weighted sum) runs on
 * the resident tile. Algorithmic bytes 4*B*rows*(D + gate_C) + 4*(2*B*D + B*rows) instead of an extra write + read of df_t.
 * Forward only (the training rollout needs df_t for the candidate logits and the backward pass). Outputs as dasa_row_attention_fwd. */
int dasa_gate_shift_attention_fwd(const float* f, int64_t ld_row, int64_t ld_sample, int B, int rows, int D,
                                  const float* gate_pre, int64_t ld_grow, int64_t ld_gsample, int gate_C,
                                  const float* chan_scale, const float* t, int64_t ld_t, int shift_k, int headings,
                                  const float* kappa_logits, int64_t ld_kappa, float* wc, int64_t ld_wc,
                                  float* attn_out, float* q_out, float* kappa_out, void* stream);
/* backward (Appendix A, K3 / K4): given dwc -> dctx (may be NULL), dt, dkappa_logits (shift only).
 * dctx_accumulate != 0 adds into dctx instead of overwriting it.                                                  */
int dasa_row_attention_bwd(const float* ctx, int64_t ld_row, int64_t ld_sample, int B, int rows, int D,
                           const float* t, int64_t ld_t, const float* attn, const float* q, const float* kappa,
                           int shift_k, int headings, const float* dwc, int64_t ld_dwc,
                           float* dctx, int64_t ldd_row, int64_t ldd_sample, int dctx_accumulate,
                           float* dt, int64_t ld_dt, float* dkappa_logits, int64_t ld_dkappa, void* stream);

/* a5: candidate logits (model.py:559 with output_prob=False; masking agent_dg.py:832-841):
 *     logit[b,c] = cand[b,c,:] . t[b,:] for c < cand_leng[b], -inf otherwise.                                     */
int dasa_cand_logits_fwd(const float* cand, int64_t ld_row, int64_t ld_sample, int B, int Nc, int D,
                         const float* t, int64_t ld_t, const int32_t* cand_leng, float* logit, void* stream);
/* backward: dcand[b,c,:] = dlogit[b,c] * t[b,:] (zero rows for c >= leng; only the first Dc channels are written),
 *           dt[b,:] = sum_c dlogit[b,c] * cand[b,c,:]                                                              */
int dasa_cand_logits_bwd(const float* cand, int64_t ld_row, int64_t ld_sample, int B, int Nc, int D,
                         const float* t, int64_t ld_t, const int32_t* cand_leng, const float* dlogit,
                         float* dcand, int64_t ldd_row, int64_t ldd_sample, int Dc, float* dt, int64_t ld_dt, void* stream);

/* --------------------------------------------------------------------------------------------- LSTM (a6, a9)
 * Pointwise part of an LSTM step, gate order i,f,g,o (nn.LSTMCell model.py:437,514; nn.LSTM r2rmodel.py:2342).
 * gates = ga (+ gb) (+ bias_a + bias_b), each [B, 4H] with its own leading dim; gb / biases may be NULL.
 * `active` (optional, [B] int32 lengths) with `pos`: rows with pos >= active[b] keep (h_prev, c_prev) and write zeros
 * to `seq_out` (packed-sequence semantics). acts_out (optional, [B,4H]) receives the post-nonlinearity gates for
 * the backward pass.                                                                                               */
int dasa_lstm_pointwise_fwd(const float* ga, int64_t ld_ga, const float* gb, int64_t ld_gb, const float* bias_a,
                            const float* bias_b, const float* c_prev, int64_t ld_cp, const float* h_prev, int64_t ld_hp,
                            float* h_out, int64_t ld_h, float* c_out, int64_t ld_c, float* seq_out, int64_t ld_seq,
                            float* acts_out, int64_t ld_acts, const int32_t* active, int pos, int B, int H,
                            const uint8_t* seq_mask, float seq_scale, void* stream);
/* seq_mask (optional, contiguous [B,H] keep flags) with seq_scale: seq_out = dropout(h') — the decoder's drop(h_1) fused
 * into the cell (model.py:515-516).
 * backward: (dh, dc_in, saved acts, c_prev, c_new) -> dgates [B,4H], dc_prev. dh2 (optional) is added to dh, through
 * dh2_mask * dh2_scale when given (the gradient of the dropped copy).
 * Inactive rows (pos >= active[b]) pass dh/dc straight through into dh_pass/dc_prev and produce zero dgates.       */
int dasa_lstm_pointwise_bwd(const float* dh, int64_t ld_dh, const float* dh2, int64_t ld_dh2, const float* dc,
                            int64_t ld_dc, const float* acts, int64_t ld_acts, const float* c_prev, int64_t ld_cp,
                            const float* c_new, int64_t ld_cn, float* dgates, int64_t ld_dg, float* dc_prev,
                            int64_t ld_dcp, float* dh_pass, int64_t ld_dhp, const int32_t* active, int pos,
                            int B, int H, const uint8_t* dh2_mask, float dh2_scale, void* stream);

/* Fused recurrence of the packed bidirectional encoder LSTM (r2rmodel.py:2339-2357) for small batches (B <= 20,
 * H % 64 == 0): ONE call issues the whole time loop (one launch per step covering both directions; each CTA owns 16
 * hidden units, streams its recurrent-weight rows from L2 and applies the pointwise cell in the same kernel).
 * Index d = 0 is the forward direction, d = 1 the reverse direction; step s processes position l = s (d=0) or L-1-s (d=1).
 *   fwd: xp[d] = x W_ih^T [B, L, 4H] (no bias); hs/cs[d] = [L+1, B, H] state BEFORE step s at index s (index 0 zero-filled
 *        by the caller); acts[d] = [L, B, 4H] post-nonlinearity gates; out = [B, L, 2H] (zero rows past each length).
 *   bwd: w_hh_t[d] = W_hh^T [H, 4H]; dout = grad of `out`; dh_fin/dc_fin[d] = grad of the final states [B, H] or NULL;
 *        dgates[d] = [L, B, 4H] (output, step order); dh_pass/dc_work[d] = [2, B, H] scratch.                          */
typedef struct {
  const float* xp[2]; const float* w_hh[2]; const float* b_ih[2]; const float* b_hh[2];
  float* hs[2]; float* cs[2]; float* acts[2]; float* out; const int32_t* lengths;
  int B, L, H;
} dasa_bilstm_fwd_t;
typedef struct {
  const float* w_hh_t[2]; const float* acts[2]; const float* cs[2]; const float* dout;
  const float* dh_fin[2]; const float* dc_fin[2];
  float* dgates[2]; float* dh_pass[2]; float* dc_work[2];
  const int32_t* lengths;
  int B, L, H;
} dasa_bilstm_bwd_t;
int dasa_bilstm_max_batch(void);
/* precision: DASA_PREC_FP32 = FFMA kernels (exact fp32); DASA_PREC_TF32 = mma.sync TF32 tensor-core kernels (H % 128 == 0) */
/* In the TF32 mode with H % 256 == 0 and H <= 1024 both calls run ONE cooperative launch for the whole time loop (recurrent weights
 * resident in shared memory as fp16, fp16 state / scaled gate-gradient exchange through L2, device-wide barrier per step) instead
 * of one launch per step; dasa_debug_bilstm_persist(0) selects the per-step kernels (1 = default, also env DASA_BILSTM_PERSIST).
 * The exchange rows and the barrier word live in a static device scratch: calls on ONE stream at a time per device (like
 * dasa_colsum's slab scratch); the launch is cooperative, so it fails instead of deadlocking if the grid cannot be co-resident. */
int dasa_debug_bilstm_persist(int mode);
int dasa_bilstm_seq_fwd(const dasa_bilstm_fwd_t* args, int precision, void* stream);
int dasa_bilstm_seq_bwd(const dasa_bilstm_bwd_t* args, int precision, void* stream);
/* Large-batch form of the same recurrence (the batched teacher-forced schedule runs all T x B instruction copies at once):
 * per time step ONE grouped launch of the persistent CTA-pair tcgen05 GEMM for both directions (TF32 products) + ONE pointwise
 * launch; backward splits K three ways and the next step's pointwise kernel sums the partials (deterministic order).
 * Same argument structs; w_hh_t[d] = W_hh[d]^T as [H, 4H] rows. workspace: dasa_bilstm_seq_gemm_workspace(B, H, backward) bytes. */
size_t dasa_bilstm_seq_gemm_workspace(int B, int H, int backward);
int dasa_bilstm_seq_gemm_fwd(const dasa_bilstm_fwd_t* args, void* workspace, size_t workspace_bytes, void* stream);
int dasa_bilstm_seq_gemm_bwd(const dasa_bilstm_bwd_t* args, void* workspace, size_t workspace_bytes, void* stream);

/* Padding-free form of the large-batch recurrence (nn.LSTM over a PackedSequence touches only valid tokens, r2rmodel.py:2339-2357).
 * The R sequences are RANKED by length (descending; perm[rank] = original sequence index) and every per-token array is stored in
 * position-block order: block p = the tokens at (reversed-sequence) position p of the n_rows[p] sequences longer than p, rank-major,
 * at rows off[p] .. off[p] + n_rows[p] of a compact [N = off[L], .] array. The forward direction's step s works on block s, the
 * reverse direction's on block L-1-s: live rows are a prefix of the ranks, so a step is one grouped tcgen05 GEMM with its own row
 * count per direction on contiguous operand blocks, and x W_ih^T / the weight gradients run over N rows instead of R x L.
 *   n_rows, off: HOST arrays ([L] non-increasing with n_rows[0] == R; [L+1] running sum). perm: DEVICE int32 [R].
 *   xp[d] [N, 4H] = x W_ih[d]^T without bias; hprev[d] [N, H] receives the state BEFORE every token (the K-major operand of the
 *   recurrent GEMM and of dW_hh); cs[d] [L+1, R, H] cell state before step s at index s, rank-major (live rows only);
 *   acts[d] [N, 4H]; out [R, L, 2H] in ORIGINAL sequence order, zero-initialised by the caller (only valid tokens are written);
 *   h_fin / c_fin [R, H] per direction in RANK order. workspace: dasa_bilstm_packed_workspace(R, H, backward) bytes.            */
typedef struct {
  int R, L, H;
  const int32_t* n_rows; const int64_t* off; const int32_t* perm;
  const float* xp[2]; const float* w_hh[2]; const float* b_ih[2]; const float* b_hh[2];
  float* hprev[2]; float* cs[2]; float* acts[2];
  float* out;
  float* h_fin[2]; float* c_fin[2];
  /* dropout on the layer output (r2rmodel.py:2357), fused into the write of `out`: out_mask = uint8 keep mask [R, L, 2H] (original
   * order), or NULL with drop_p > 0: flags drawn in place, out[seq, l, c] = stream byte (seq*L + l)*2H + c of (seed, drop_base)
   * (see dasa_mha_fwd_h16); NULL and drop_p == 0: no dropout. The backward struct takes the same fields and masks dout.          */
  const uint8_t* out_mask; const uint64_t* drop_seed_dev; uint64_t drop_seed; uint64_t drop_base; float drop_p; float drop_scale;
  /* Fused recurrence (all four non-NULL and H % 64 == 0; else the two-launch form): w_hh16[d] = dasa_lstm_whh_interleave_f16 of
   * W_hh[d], h16[d] [N, H] fp16 scratch (the state rows as the fp16 A operand). One launch per time step: tcgen05 kind::f16
   * GEMM of both directions with the LSTM cell applied to the accumulator (gates never reach memory as pre-activations).       */
  const dasa_half_t* w_hh16[2]; dasa_half_t* h16[2];
} dasa_bilstm_packed_fwd_t;
/* out [4H, H] fp16 = the rows of W_hh ([4H, H] fp32, gate-major i, f, g, o) interleaved so that every 128 consecutive output
 * columns of h W_hh^T are the four gates of 32 consecutive hidden units: row nt*256 + ch*128 + gate*32 + ul <- gate*H + nt*64 +
 * ch*32 + ul. H % 64 == 0.                                                                                                       */
int dasa_lstm_whh_interleave_f16(const float* w_hh, dasa_half_t* out, int H, void* stream);
/* Backward: dout [R, L, 2H] (original order), dh_fin / dc_fin [R, H] rank order (may be NULL), w_hh_t[d] = W_hh[d]^T [H, 4H].
 * Writes dgates[d] [N, 4H] (compact token order: the dY of dW_ih = dgates^T x, dW_hh = dgates^T hprev, dX = dgates W_ih).
 * dc_work[d]: scratch [2, R, H].                                                                                               */
typedef struct {
  int R, L, H;
  const int32_t* n_rows; const int64_t* off; const int32_t* perm;
  const float* w_hh_t[2]; const float* acts[2]; const float* cs[2];
  const float* dout; const float* dh_fin[2]; const float* dc_fin[2];
  float* dgates[2]; float* dc_work[2];
  const uint8_t* out_mask; const uint64_t* drop_seed_dev; uint64_t drop_seed; uint64_t drop_base; float drop_p; float drop_scale;
  /* fp16 recurrence (all four non-NULL and H % 64 == 0; else TF32): w_hh_t16[d] = fp16 copy of w_hh_t[d], dg16[d] [N, 4H] fp16
   * scratch receiving dgates * 2^8 (saturating) = the A operand of dh = dgates W_hh on tcgen05 kind::f16 (the sums are rescaled
   * by 2^-8; fp16 keeps TF32's 11 significant bits for |dgate| in [2.4e-7, 256)).                                               */
  /* With the fp16 recurrence dgates[d] may be NULL: only the scaled fp16 copy is written (the weight / bias gradients read it).  */
  const dasa_half_t* w_hh_t16[2]; dasa_half_t* dg16[2];
} dasa_bilstm_packed_bwd_t;
size_t dasa_bilstm_packed_workspace(int R, int H, int backward);
int dasa_bilstm_packed_fwd(const dasa_bilstm_packed_fwd_t* args, void* workspace, size_t workspace_bytes, void* stream);
int dasa_bilstm_packed_bwd(const dasa_bilstm_packed_bwd_t* args, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------- persistent decoder rollout
 * BAttnDecoderLSTM.forward (model.py:472-574) for T consecutive actions of B <= 32 episodes in ONE cooperative launch: one CTA
 * per SM stays resident for the whole rollout, the phases of an action are separated by a device-wide barrier instead of
 * kernel boundaries (8 barriers instead of ~28 dependent launches per action). Per action t:
 *   P1  tk = [linear_in ; linear_shift] drop(h~_{t-1}) + b           weight-streaming mma.sync TF32, 16 W rows per CTA
 *   P2  ShiftSoftDotAttention over feat[t] (model.py:327-345)        channel slices resident in shared memory (bulk-async
 *                                                                     prefetched one action ahead), partial dots through L2
 *   P3  gates = [W_ih | W_hh] [emb ; attn_feat ; h~_{t-1}] + b, LSTM cell fused in the epilogue (CTA = 8 units x 4 gates)
 *   P4  t2 = attention_layer.linear_in drop(h_1)
 *   P5  SoftDotAttention over ctx[t] with the padding mask (model.py:276-288)
 *   P6  h~_t = tanh(linear_out [wc ; drop(h_1)])  (+ drop(h~_t) for the next action)
 * T = 1 is the per-action form the sampled / greedy rollouts use. All [T, B, .] buffers are contiguous unless a stride is
 * given. Everything the backward pass re-reads is written to caller-owned buffers (no recomputation).
 * Dropout: m_hprev / m_h1 are [T, B, H] keep masks (NULL = eval), scale = 1/(1-p); emb and feat arrive already dropped.
 * Requirements: H % 16 == 0, (E+F+H) % 32 == 0, NK % 32 == 0, D % 4 == 0, F % 4 == 0, V, L <= 128, B <= 32, shift_k <= 15;
 * dasa_decoder_rollout_supported() says whether the shared-memory plan fits (else use the per-op entry points).      */
/* out[i] = fp16(in[i]) (round to nearest even), n elements; in 16-byte aligned when n >= 8. Used for the decoder kernel's weights. */
int dasa_f32_to_f16(const float* in, dasa_half_t* out, int64_t n, void* stream);
/* out[r, c] = fp16(in[r * ld_in + c]) for c < C (C % 8 == 0, ld_in % 4 == 0, ld_out % 8 == 0): a column slice of strided rows
 * (the RGB part of the depth features, the K-major A operand of the AdaIN gate GEMM on dasa_gemm_f16).                    */
int dasa_f32_to_f16_rows(const float* in, int64_t ld_in, dasa_half_t* out, int64_t ld_out, int R, int C, void* stream);
typedef struct {
  int T, B, H, E, F, V, L, D, headings, shift_k, NK;
  const float* emb;                                   /* [T, B, E] drop(tanh(embedding(action)))  (model.py:504-505)        */
  const float* feat; int64_t feat_ld_row, feat_ld_b, feat_ld_t;   /* [T, B, V, F] AdaIN'd (and dropped) views            */
  const float* ctx; int64_t ctx_ld_row, ctx_ld_b, ctx_ld_t;       /* [T, B, L, D] encoder context per action              */
  const uint8_t* ctx_mask; int64_t ctx_mask_ld;       /* [B, L] 1 = padding (NULL = none)                                  */
  const float* h0; const float* c0;                   /* [B, H] recurrent state before action 0 (h~_{-1}, c_{-1})          */
  const uint8_t* m_hprev; const uint8_t* m_h1; float drop_scale;
  /* Weights are FP16 copies of the fp32 parameters (dasa_f32_to_f16, refreshed after every optimizer step): fp16's 10-bit mantissa
   * is TF32's, so the products equal the TF32 tensor-core products of the fp32 weights, at half the bytes per action.         */
  const dasa_half_t* w_feat; const float* b_feat;     /* [NK, H] = [feat_att.linear_in ; linear_shift ; 0], bias [NK] fp32 */
  const dasa_half_t* w_lstm; const float* b_ih; const float* b_hh;   /* [4H, E+F+H] = [W_ih | W_hh]                       */
  const dasa_half_t* w_att_in;                        /* [D, H]   attention_layer.linear_in                                */
  const dasa_half_t* w_att_out;                       /* [H, D+H] attention_layer.linear_out                               */
  float* hprev_drop;                                  /* [T, B, H]   drop(h~_{t-1})                                        */
  float* tk;                                          /* [T, B, NK]  attention target | shift logits                      */
  float* p; float* q; float* kappa;                   /* [T, B, V] pre-shift softmax, [T, B, V] shifted, [T, B, shift_k]   */
  float* xh;                                          /* [T, B, E+F+H] = [emb ; attn_feat ; h~_{t-1}]                      */
  float* acts;                                        /* [T, B, 4H]  post-nonlinearity gates i,f,g,o                       */
  float* c;                                           /* [T+1, B, H] cell state, slot 0 = c0                               */
  float* h1;                                          /* [T, B, H]                                                         */
  float* cat;                                         /* [T, B, D+H] = [wc ; drop(h_1)]                                    */
  float* t2;                                          /* [T, B, D]                                                         */
  float* alpha;                                       /* [T, B, L]                                                         */
  float* htilde;                                      /* [T, B, H]                                                         */
  float* zpart;                                       /* scratch, dasa_decoder_rollout_scratch_floats(B) floats            */
  unsigned int* barrier;                              /* scratch, 4 bytes, zeroed by the call                              */
  /* scratch, dasa_decoder_rollout_x16_halves(T, B, H, E, F, D) halves, 16-byte aligned: fp16 copies of the GEMM operands
   * (drop(h~), [emb ; attn ; h~], [wc ; drop(h_1)]). Every CTA re-reads the whole operand in every GEMM phase, so their bytes - not
   * the weights' - dominate the L2 traffic of an action; the operands are O(1), so fp16 keeps the 11 significant bits TF32 uses.  */
  dasa_half_t* x16;
} dasa_decoder_fwd_t;
size_t dasa_decoder_rollout_x16_halves(int T, int B, int H, int E, int F, int D);
/* Backward of the same T actions (reverse order) in one cooperative launch. Weight operands are the TRANSPOSED weights
 * ([in, out] rows with leading dimension ld_*): dX = dY.W then streams K-major rows like the forward. Weight gradients are NOT
 * formed here: du, dt2, dgates, dtk (the dY of the four projections) are left in [T, B, .] buffers for one long-K GEMM each.
 *   d_htilde [T, B, H]: gradient arriving at every h~_t from outside the recurrence (candidate logits); d_h1 (optional): at h_1
 *   (the critic reads it in the sampled rollout); d_c_last (optional): at the last cell state (per-action use, T = 1). Outputs: demb [T, B, E], dfeat [T, B, V, F] (strides given), dctx [T, B, L, D]
 *   contiguous (zero rows where masked), dh0, dc0 [B, H].                                                                    */
typedef struct {
  int T, B, H, E, F, V, L, D, headings, shift_k, NK;
  const float* feat; int64_t feat_ld_row, feat_ld_b, feat_ld_t;
  const float* ctx; int64_t ctx_ld_row, ctx_ld_b, ctx_ld_t;
  const uint8_t* ctx_mask; int64_t ctx_mask_ld;
  const uint8_t* m_hprev; const uint8_t* m_h1; float drop_scale;
  const dasa_half_t* w_feat_t; int64_t ld_w_feat_t;         /* [H, NK]     fp16, leading dimensions in elements                  */
  const dasa_half_t* w_lstm_t; int64_t ld_w_lstm_t;         /* [E+F+H, 4H]                                                       */
  const dasa_half_t* w_att_in_t; int64_t ld_w_att_in_t;     /* [H, D]                                                            */
  const dasa_half_t* w_att_out_t; int64_t ld_w_att_out_t;   /* [D+H, H]                                                          */
  const float* tk; const float* p; const float* q; const float* kappa; const float* acts; const float* c;
  const float* cat; const float* t2; const float* alpha; const float* htilde;
  const float* d_htilde; const float* d_h1; const float* d_c_last;   /* d_c_last (optional) [B, H]: gradient at c_{T-1}      */
  float* du; float* dt2; float* dgates; float* dtk;   /* [T, B, H], [T, B, D], [T, B, 4H], [T, B, NK]                      */
  float* demb;                                        /* [T, B, E]                                                         */
  float* dfeat; int64_t dfeat_ld_row, dfeat_ld_b, dfeat_ld_t;
  float* dctx;                                        /* [T, B, L, D] contiguous                                           */
  float* dh0; float* dc0;                             /* [B, H]                                                            */
  float* dcat; float* dattn; float* dhdir; float* dc_carry;   /* scratch [B, D+H], [B, F], [B, H], [B, H]                  */
  float* zpart; unsigned int* barrier;
  /* scratch, dasa_decoder_rollout_g16_halves(T, B, H, D, NK) halves, 16-byte aligned: fp16 copies of the gradient operands du,
   * dt2, dgates, dtk scaled by 2^8 (saturating), the X operand of the four dX projections; the sums are rescaled by 2^-8.       */
  dasa_half_t* g16;
} dasa_decoder_bwd_t;
size_t dasa_decoder_rollout_g16_halves(int T, int B, int H, int D, int NK);
int dasa_decoder_rollout_supported(int B, int H, int E, int F, int V, int L, int D, int NK, int shift_k);
size_t dasa_decoder_rollout_scratch_floats(int B);
/* Profiling aid: SM-clock timestamps of CTA 0 at the phase barriers of the first 4 actions of the LAST rollout launch (forward or
 * backward): out[0] = after the prologue, out[1 + 8*i + ph] = after phase ph of the i-th processed action. Returns the count. */
int dasa_debug_decoder_phase_clocks(long long* out, int n);
/* Profiling aid: `iters` device-wide barriers of the persistent kernels (one CTA per SM, cooperative launch) with one global write
 * per thread in between; out[0] (device) = SM clocks CTA 0 spent. variant 0 = the barrier the rollout kernels use, 1 = release /
 * acquire only, 2 = fence + relaxed atomic + volatile poll. junk: >= 256 * SMs floats, ctr: 4 bytes.                          */
int dasa_debug_barrier_bench(int variant, int iters, unsigned int* ctr, long long* out, float* junk, void* stream);
int dasa_decoder_rollout_fwd(const dasa_decoder_fwd_t* args, void* stream);
int dasa_decoder_rollout_bwd(const dasa_decoder_bwd_t* args, void* stream);

/* ------------------------------------------------------------------------------------------- encoder pieces (a9)
 * BertEmbeddings (vilmodel.py:161-176): out[b,l,:] = LN(word[ids[b,l]] + pos[l] + type[0]) (* mask*scale).          */
int dasa_embed_layernorm(const int64_t* ids, int64_t ld_ids, int B, int L, int Hd, const float* word, const float* pos,
                         const float* type0, const float* gamma, const float* beta, float eps,
                         const uint8_t* drop_mask, float drop_scale, float* out, void* stream);
/* BertSelfOutput / BertOutput / VisionEncoder tail (vilmodel.py:246-250, 305-309, 1089-1094):
 *   z = x (* mask*scale) (+ resid);  out = LN(z) * gamma + beta (* post_mask*post_scale); rows R, width Hd <= 1024.
 *   stats_out (optional, [R,2]) receives (mean, rstd) and z_out (optional, [R,Hd] contiguous) the LN input, both
 *   only needed by the backward pass (finetune config).                                                             */
int dasa_dropout_residual_layernorm(const float* x, int64_t ldx, const uint8_t* drop_mask, float drop_scale,
                                    const float* resid, int64_t ldr, const float* gamma, const float* beta, float eps,
                                    const uint8_t* post_mask, float post_scale, float* out, int64_t ldo,
                                    float* stats_out, float* z_out, dasa_half_t* out_half, int R, int Hd, void* stream);
/* out_half (optional): an fp16 copy of `out`, [R, Hd] contiguous - the A operand of the next dasa_gemm_f16.            */
/* backward of the above (finetune config): given dout -> dresid (gradient w.r.t. z, i.e. w.r.t. the residual input) and
 * dx (gradient w.r.t. x, = dresid * mask*scale); ACCUMULATES dgamma/dbeta (atomics).                                */
int dasa_layernorm_bwd(const float* dout, int64_t lddo, const float* z, const float* gamma, const float* stats,
                       const uint8_t* drop_mask, float drop_scale, const uint8_t* post_mask, float post_scale,
                       float* dx, int64_t lddx, float* dresid, int64_t lddr, float* dgamma, float* dbeta,
                       int R, int Hd, void* stream);
/* Multi-head attention for short sequences (vilmodel.py:203-236, 479-506): one CTA per (batch, head).
 *   q [B,Lq,*], k/v [B,Lk,*] with row strides ldq/ldk/ldv and sample strides sq/sk/sv (so fused QKV buffers are
 *   addressed in place); scores = q k^T / sqrt(dh) + (key_pad[b,j] ? -10000 : 0); softmax; optional keep mask on the
 *   probabilities [B,heads,Lq,Lk]; out[b,i,head*dh:(head+1)*dh] = P v. probs_out optional (for backward).
 *   precision DASA_PREC_TF32 runs Q K^T and P V on mma.sync TF32 tensor cores (softmax in fp32).                   */
int dasa_mha_fwd(const float* q, int64_t ldq, int64_t sq, const float* k, int64_t ldk, int64_t sk, const float* v,
                 int64_t ldv, int64_t sv, const uint8_t* key_pad, int64_t ld_pad, const uint8_t* drop_mask,
                 float drop_scale, float* out, int64_t ldo, int64_t so, float* probs_out,
                 int B, int heads, int Lq, int Lk, int dh, int precision, int out_half, void* stream);
/* out_half != 0 (forward-only, probs_out NULL): `out` receives IEEE halves (ldo / so in halves) for dasa_gemm_f16.      */
int dasa_mha_bwd(const float* q, int64_t ldq, int64_t sq, const float* k, int64_t ldk, int64_t sk, const float* v,
                 int64_t ldv, int64_t sv, const float* probs, const uint8_t* drop_mask, float drop_scale,
                 const float* dout, int64_t ldo, int64_t so, float* dq, int64_t lddq, int64_t sdq, float* dk,
                 int64_t lddk, int64_t sdk, float* dv, int64_t lddv, int64_t sdv,
                 int B, int heads, int Lq, int Lk, int dh, void* stream);
/* Packed (variable-length) forward of the same attention for the frozen language / cross-modal stack: the valid tokens of
 * all sequences are stored back to back (no padding rows are computed at all - padded keys carry an additive -10000 in the
 * reference, i.e. an exact 0 after the fp32 softmax, so the valid outputs are unchanged). For a packed operand, sample b
 * owns rows [off[b], off[b]+len[b]); pass NULL off/len for a dense operand ([B, max_L, *], sample stride dense_*_stride
 * elements, e.g. the 36 panorama views). out is laid out like q. drop_mask keeps the padded shape [B,heads,max_Lq,max_Lk]. */
int dasa_mha_fwd_varlen(const float* q, int64_t ldq, const int32_t* q_off, const int32_t* q_len, const float* k,
                        int64_t ldk, const float* v, int64_t ldv, const int32_t* k_off, const int32_t* k_len,
                        int64_t dense_q_stride, int64_t dense_kv_stride, const uint8_t* drop_mask, float drop_scale,
                        float* out, int64_t ldo, int B, int heads, int max_Lq, int max_Lk, int dh, int precision,
                        int out_half, void* stream);
/* The same attention on fp16 operands for the frozen (forward-only) stack, dh = 64 (vilmodel.py:203-236, 479-506, outputs
 * detached at :1377-1410): q / k / v are IEEE halves as written by dasa_gemm_f16 with c_half = 1 (leading dimensions and sample
 * strides in halves, multiples of 8; fused QKV buffers are addressed in place), packed (off / len) or dense (NULL) per operand
 * like dasa_mha_fwd_varlen. Persistent CTAs, 2-stage cp.async ring over the (sample, head) units, mma.sync m16n8k16, fp32
 * softmax. Probability dropout: `drop_mask` (uint8 keep mask [B, heads, max_Lq, max_Lk], tests) or, when it is NULL and
 * drop_p > 0, keep flags drawn in place from the counter-hash stream (seed read from drop_seed_dev when non-NULL, else
 * drop_seed; hash index base drop_base): the flags of probability [b, h, r, j] are stream byte
 *   ((((b*heads + h)*nMT + r/16)*nNT + j/8)*32 + (r%8)*4 + (j%8)/2)*4 + 2*(j%2) + (r%16)/8,  nMT = ceil(max_Lq/16),
 *   nNT = 2*ceil(max_Lk/16), i.e. byte i of dasa_dropout_mask(mask, n, p, seed, drop_base). out: fp16 (out_half) or fp32.        */
int dasa_mha_fwd_h16(const dasa_half_t* q, int64_t ldq, int64_t sq, const int32_t* q_off, const int32_t* q_len,
                     const dasa_half_t* k, int64_t ldk, const dasa_half_t* v, int64_t ldv, int64_t skv,
                     const int32_t* k_off, const int32_t* k_len, const uint8_t* key_pad, int64_t ld_pad,
                     const uint8_t* drop_mask, const uint64_t* drop_seed_dev, uint64_t drop_seed, uint64_t drop_base,
                     float drop_p, float drop_scale, void* out, int64_t ldo, int64_t so, int out_half, int B, int heads,
                     int max_Lq, int max_Lk, int dh, void* stream);
/* Forward-only dropout -> + resid -> LayerNorm of the frozen stack (vilmodel.py:246-250, 305-309): like
 * dasa_dropout_residual_layernorm without the backward outputs, x optionally fp16 (x_half: halves, ldx in halves - the
 * c_half output of dasa_gemm_f16) and the keep flags either from `drop_mask` or, when it is NULL and drop_p > 0, drawn in
 * place: element [row, c] = stream byte row*Hd + c of (seed, drop_base) (see dasa_mha_fwd_h16). */
int dasa_dropout_residual_layernorm_fwd(const void* x, int x_half, int64_t ldx, const uint8_t* drop_mask,
                                        const uint64_t* drop_seed_dev, uint64_t drop_seed, uint64_t drop_base,
                                        float drop_p, float drop_scale, const float* resid, int64_t ldr,
                                        const float* gamma, const float* beta, float eps, float* out, int64_t ldo,
                                        dasa_half_t* out_half, int R, int Hd, void* stream);
/* dst[r,:] = src[idx[r],:] for r < R (C % 4 == 0): packs the valid tokens of a padded [B*L, C] activation             */
int dasa_gather_rows(const float* src, int64_t ld_src, const int32_t* idx, float* dst, int64_t ld_dst, int R, int C,
                     void* stream);
/* token reversal straight from the packed layout into the padded [B, L, Hd] input of the bi-LSTM                        */
int dasa_reverse_tokens_packed(const float* x, const int32_t* offsets, const int32_t* lengths, float* out, int B, int L,
                               int Hd, void* stream);
/* token reversal (r2rmodel.py:2326-2330): out[b,i,:] = x[b, len_b-1-i, :] for i < len_b else 0. Self-inverse, so
 * the same call maps gradients back.                                                                                */
int dasa_reverse_tokens(const float* x, float* out, const int32_t* lengths, int B, int L, int Hd, void* stream);

/* ----------------------------------------------------------------------------------------- elementwise helpers */
/* y = x * mask * scale (mask may be NULL -> copy); strided rows                                                     */
int dasa_dropout_apply(const float* x, int64_t ldx, const uint8_t* mask, float scale, float* y, int64_t ldy,
                       int R, int C, void* stream);
/* dx = dy * (1 - y*y) (tanh), dx = dy * (y > 0) (relu), dx = dy * gelu'(y) with y = the PRE-activation (gelu, finetune
 * config), optional keep mask folded in front (dy*mask*scale)                                                       */
int dasa_act_backward(int act /*0 tanh, 1 relu, 2 gelu*/, const float* dy, int64_t lddy, const float* y, int64_t ldy,
                      const uint8_t* mask, float scale, float* dx, int64_t lddx, int R, int C, void* stream);
/* y = gelu_erf(x) over n contiguous elements (vilmodel.py:125-131); the forward-only path fuses it in the GEMM epilogue */
int dasa_gelu_fwd(const float* x, float* y, int64_t n, void* stream);
/* y (+)= a*x   elementwise over [R,C] strided                                                                      */
int dasa_axpy2d(float a, const float* x, int64_t ldx, float* y, int64_t ldy, int accumulate, int R, int C, void* stream);
/* keep-mask generator (counter-based hash RNG): mask[i] = u(seed, offset+i) >= p                                   */
int dasa_dropout_mask(uint8_t* mask, int64_t n, float p, uint64_t seed, uint64_t offset, void* stream);
/* same, with the seed read from device memory at run time, so a captured CUDA graph draws fresh masks on every replay
 * (dasa_bump_counter advances that seed inside the graph)                                                           */
int dasa_dropout_mask_dev(uint8_t* mask, int64_t n, float p, const uint64_t* seed_dev, uint64_t offset, void* stream);
int dasa_bump_counter(uint64_t* counter, uint64_t inc, void* stream);
/* out[c*ld_out + r] = in[r*ld_in + c]  (operand re-layout for the K-major tensor-core GEMM in the backward pass)    */
int dasa_transpose(const float* in, int64_t ld_in, int rows, int cols, float* out, int64_t ld_out, void* stream);

/* --------------------------------------------------------------------------------------- loss / action (a10, a13)
 * nn.CrossEntropyLoss(ignore_index, sum) over masked logits + argmax (agent_dg.py:850, 870-873):
 *   loss_acc[0] += sum_b CE(logit[b,:], target[b]) ; dlogit = (softmax - onehot) * grad_scale for valid targets, 0
 *   for ignored rows; action[b] = argmax (first maximal index, like torch.max).                                    */
int dasa_masked_ce(const float* logit, int64_t ld, const int64_t* target, int ignore_index, int B, int Nc,
                   float grad_scale, float* loss_acc, float* dlogit, int64_t* action, float* logprob_action,
                   float* entropy, void* stream);

/* ------------------------------------------------------------------------------- sampled feedback / A2C (a10, a11)
 * feedback='sample' (agent_dg.py:876-882): probs = softmax(logit); entropy = Categorical(probs).entropy();
 * action = injected (action_in != NULL) | inverse-CDF sample with the uniform u[b] (u != NULL) | argmax;
 * logprob = log probs[action]. probs [B,Nc] is kept for the backward:
 *   dlogit_j = dlogp (delta_aj - p_j) - dent p_j (log p_j + H).                                                      */
int dasa_policy_sample_fwd(const float* logit, int64_t ld, int B, int Nc, const float* u, const int64_t* action_in,
                           int64_t* action, float* logprob, float* entropy, float* probs, void* stream);
int dasa_policy_sample_bwd(const float* probs, const int64_t* action, const float* dlogp, const float* dent,
                           const float* entropy, int B, int Nc, float* dlogit, int64_t ld, void* stream);
/* reward / mask / ended bookkeeping of one action (agent_dg.py:890-930): END = last candidate or ignore id; reward +-2 on
 * END by dist < 3, else sign of the distance reduction; mask = 0 for episodes that had already ended; ended |= END.   */
int dasa_nav_reward(const int64_t* action, const int32_t* cand_leng, int ignore_id, const float* dist,
                    const float* last_dist, uint8_t* ended, float* reward, float* mask, int B, void* stream);
/* Speaker greedy decode (speaker.py:318-343), word selection of one step: logits[:, unk] = -inf; word = argmax (first index on
 * ties); emitted[b * ld_emitted] = ended[b] ? pad : word; next_word[b] = word (fed to the next step); ended |= emitted == eos. */
int dasa_speaker_select(const float* logit, int64_t ld, int B, int V, int unk, int pad, int eos, uint8_t* ended,
                        int64_t* next_word, int64_t* emitted, int64_t ld_emitted, void* stream);
/* A2C epilogue (agent_dg.py:943-999) over [T,B] stacks: R_b = ended_b ? 0 : last_value_b; for t = T-1..0:
 *   R = R*gamma + reward_t; a = R - value_t; loss += sum_b (-logp_t a m_t + 0.5 a^2 m_t - ent_coef ent_t m_t); total += m_t;
 * loss /= total (normalize 1) | B (2) | 1 (0). Also writes dloss/dlogp, dloss/dent, dloss/dvalue (the advantage in the
 * policy term is detached, as in the reference). ent / dent may be NULL (feedback='argmax').                          */
int dasa_a2c_loss(const float* logp, const float* ent, const float* value, const float* last_value, const float* reward,
                  const float* mask, const uint8_t* ended, float gamma, float ent_coef, int normalize, int T, int B,
                  float* loss, float* total, float* dlogp, float* dent, float* dvalue, void* stream);

/* ----------------------------------------------------------------------------------------------- optimizer (a12)
 * torch.optim.RMSprop step (alpha=0.99, eps=1e-8, no momentum, not centered; agent_dg.py:214-241) fused over one
 * flat parameter group, with the clip coefficient of clip_grad_norm (agent_dg.py:1392-1393) folded in:
 *   g = grad * clip_coef[0] (clip_coef on device, may be NULL); sq = alpha*sq + (1-alpha) g^2; p -= lr * g/(sqrt(sq)+eps) */
int dasa_rmsprop_step(float* param, const float* grad, float* square_avg, int64_t n, float lr, float alpha, float eps,
                      float weight_decay, const float* clip_coef, const float* lr_scale, void* stream);
/* lr_scale (device, may be NULL): the step runs with lr * lr_scale[0] - the LambdaLR multiplier of the decoder / critic / adaIn
 * optimizers (agent_dg.py:219-241) kept on the device so that a captured CUDA graph follows the schedule.
 * dasa_lr_lambda: mult[0] = lr_lambda(*iter) (linear warm-up over warm_steps, 1 until decay_start, then lr_decay ^
 * ((iter - decay_start) / decay_intervals)), then *iter += advance.                                                            */
int dasa_lr_lambda(int* iter, int warm_steps, int decay_start, int decay_intervals, float lr_decay, float* mult, int advance,
                   void* stream);
/* sum of squares of a flat buffer, accumulated into out[0] (for the global grad norm)                              */
int dasa_sumsq(const float* x, int64_t n, float* out, void* stream);
/* clip_coef[0] = min(1, max_norm / (sqrt(sumsq[0]) + 1e-6))                                                         */
int dasa_clip_coef(const float* sumsq, float max_norm, float* clip_coef, void* stream);

/* --------------------------------------------------------------------- device-resident environment
 * Tables (dasa_b200/navgraph.py; replace R2RBatch.paths / .distances / buffered_state_dict, env.py:182-198, 291-298):
 *   rgb_bank, dep_bank [n_vp, V, C]; nbr, nbr_point [n_vp, dmax]; deg [n_vp]; cand_angle [n_vp, dmax, 12, 4];
 *   view_angle [12, V, 4]; agent_angle [V, 4]; dist_tab [n_vp, n_vp]; next_hop [n_vp, n_vp] (candidate slot, -1 at goal).
 * Episode state [B]: vp (viewpoint), view (viewIndex 0..35), goal, ended.
 * dasa_env_observe = env.py:_get_obs + make_candidate (buffered branch, :299-311) + agent_dg.py get_input_feat (:313-323),
 *   _candidate_variable (:300-311), _teacher_action (:325-344): writes f_t/d_t [B, V, C+A] (sample stride ld_f_sample),
 *   cand/cand_d [B, nc, C+A] (sample stride ld_c_sample; slot deg = END row = zeros, later slots zero), input_a_t [B, A],
 *   cand_leng [B] (= deg + 1), target [B] (teacher candidate index, deg = STOP at the goal, ignore_id once ended; may be
 *   NULL), dist [B] (distance to the goal; may be NULL). ended may be NULL.                                           */
int dasa_env_observe(const float* rgb_bank, const float* dep_bank, const int32_t* nbr, const int32_t* nbr_point,
                     const int32_t* deg, const float* cand_angle, const float* view_angle, const float* agent_angle,
                     const float* dist_tab, const int32_t* next_hop, int n_vp, int dmax, const int32_t* vp,
                     const int32_t* view, const int32_t* goal, const uint8_t* ended, int B, int V, int C, int A, int nc,
                     int headings, int ignore_id, float* f_t, float* d_t, int64_t ld_f_sample, float* cand,
                     float* cand_d, int64_t ld_c_sample, float* input_a_t, int32_t* cand_leng, int64_t* target,
                     float* dist, void* stream);
/* dasa_env_step = agent_dg.py:890-935: END = (action == deg) or ignore_id; otherwise make_equiv_action (:358-391): view =
 *   the candidate's pointId, vp = the candidate's viewpoint (ended episodes keep moving under sampled feedback, as in the
 *   reference); d = dist_tab[vp, goal]; reward = +-2 on END by d < 3, else sign(last_dist - d), 0 and mask 0 if already
 *   ended; ended |= END; last_dist = d. traj_vp / traj_view [B] (may be NULL) record the state after the action.
 *   err[0] |= 1 for an action outside the candidate list, |= 2 where the reference raises "The action doesn't change
 *   the move" (agent_dg.py:925).                                                                                       */
int dasa_env_step(const int64_t* action, int ignore_id, const int32_t* nbr, const int32_t* nbr_point, const int32_t* deg,
                  int dmax, const float* dist_tab, int n_vp, int32_t* vp, int32_t* view, const int32_t* goal,
                  uint8_t* ended, float* last_dist, float* reward, float* mask, int32_t* traj_vp, int32_t* traj_view,
                  int32_t* err, int B, void* stream);

/* --submit "avoiding cyclic path" (agent_dg.py:834-840): add the current viewpoint of every episode to its visited set
 * (`visited`: [B, words] uint32 bitmap over the viewpoints, words*32 >= n_vp, zeroed at reset), then mask every candidate whose
 * viewpoint was visited: blocked[b, k] = 1 (uint8 [B, nc], may be NULL) and logit[b, k] = -inf (fp32 [B, ld_logit], may be NULL).
 * The END slot is never masked (it is not an entry of ob['candidate']). */
int dasa_env_visited_mask(const int32_t* vp, const int32_t* nbr, const int32_t* deg, int dmax, int n_vp, uint32_t* visited,
                          int words, uint8_t* blocked, int nc, float* logit, int64_t ld_logit, int B, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DASA_B200_H */
