// Epilogue shared by the FFMA and tcgen05 GEMM kernels.
#pragma once
#include "common.cuh"

struct EpiParams {
  const float* bias;
  const float* gate_src;
  int64_t ld_gate;
  float* gate_out;
  int64_t ld_gate_out;
  const uint8_t* drop_mask;
  float drop_scale;
};

static inline EpiParams make_epi(const dasa_epilogue_t* e) {
  EpiParams p{};
  if (e) {
    p.bias = e->bias; p.gate_src = e->gate_src; p.ld_gate = e->ld_gate; p.gate_out = e->gate_out;
    p.ld_gate_out = e->ld_gate_out; p.drop_mask = e->drop_mask; p.drop_scale = e->drop_scale;
  }
  return p;
}

// gelu(x) = x * 0.5 * (1 + erf(x / sqrt(2)))  (vilmodel.gelu, vilmodel.py:125-131) with erf from Abramowitz & Stegun 7.1.26:
// erf(z) = 1 - (a1 t + .. + a5 t^5) exp(-z^2), t = 1 / (1 + p z), |error| <= 1.5e-7 (fp32 rounding level). Branch-free, two
// MUFU ops (rcp, ex2) + ~12 FMA-pipe instructions instead of erff's ~35 with a divergent branch: the erf epilogue of the
// 256 x 256 GELU tiles took three times as long as their fp16 main loop (166 us for 20300 x 3072 x 768; profiles/r02_*).
// With q = 0.5 * poly * exp(-z^2): x >= 0 -> x - x q, x < 0 -> x q (no cancellation in either tail) = max(x, 0) - |x| q.
__device__ __forceinline__ float gelu_erf(float x) {
#ifdef DASA_GELU_EXACT                                   // the exact-fp32 FFMA kernels (DASA_PREC_FP32) keep libm's erff
  return x * 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
#endif
  // raw MUFU approximations (rcp: the denominator is >= 1; ex2: the argument is <= 0 and a flushed denormal is an exact-enough
  // zero) - the libm-conforming exp2f / division wrap each in a denormal rescue (3 + 4 more instructions per element)
  const float ax = fabsf(x);
  float t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.23164190f, ax, 1.0f)));          // p / sqrt(2)
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"((x * x) * -0.72134752044448170f));      // exp(-x^2 / 2)
  float poly = fmaf(t, 1.061405429f, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float q = (0.5f * t) * poly * e;
  return fmaf(-ax, q, fmaxf(x, 0.f));                       // x >= 0: x - x q;  x < 0: x q
}

// v = alpha*acc + beta*C already applied by the caller. Compile-time epilogue kind: each kernel instantiation carries
// exactly one variant (a runtime switch inside the fully unrolled accumulator loops made the kernels instruction-fetch bound).
template <int EPI>
__device__ __forceinline__ constexpr bool epi_has_bias() { return EPI != DASA_EPI_NONE && EPI != DASA_EPI_TANH; }

// everything after the bias add (callers that loop over rows hoist the per-column bias load out of the loop)
template <int EPI>
__device__ __forceinline__ float apply_activation_t(float v, int m, int n, int N, const EpiParams& ep) {
  if constexpr (EPI == DASA_EPI_BIAS_TANH || EPI == DASA_EPI_TANH) v = tanhf(v);
  if constexpr (EPI == DASA_EPI_BIAS_GELU) v = gelu_erf(v);
  if constexpr (EPI == DASA_EPI_BIAS_RELU) v = fmaxf(v, 0.f);
  if constexpr (EPI == DASA_EPI_GATE) {
    const float s = sigmoidf_(v);
    if (ep.gate_out != nullptr) ep.gate_out[(int64_t)m * ep.ld_gate_out + n] = s;
    v = s * __ldg(ep.gate_src + (int64_t)m * ep.ld_gate + n);
  }
  if (ep.drop_mask != nullptr) v *= ep.drop_mask[(int64_t)m * N + n] ? ep.drop_scale : 0.f;
  return v;
}

template <int EPI>
__device__ __forceinline__ float apply_epilogue_t(float v, int m, int n, int N, const EpiParams& ep) {
  if constexpr (EPI != DASA_EPI_NONE && EPI != DASA_EPI_TANH) {
    if (ep.bias != nullptr) v += __ldg(ep.bias + n);
  }
  if constexpr (EPI == DASA_EPI_BIAS_TANH || EPI == DASA_EPI_TANH) v = tanhf(v);
  if constexpr (EPI == DASA_EPI_BIAS_GELU) v = gelu_erf(v);
  if constexpr (EPI == DASA_EPI_BIAS_RELU) v = fmaxf(v, 0.f);
  if constexpr (EPI == DASA_EPI_GATE) {
    const float s = sigmoidf_(v);
    if (ep.gate_out != nullptr) ep.gate_out[(int64_t)m * ep.ld_gate_out + n] = s;
    v = s * __ldg(ep.gate_src + (int64_t)m * ep.ld_gate + n);
  }
  if (ep.drop_mask != nullptr) v *= ep.drop_mask[(int64_t)m * N + n] ? ep.drop_scale : 0.f;
  return v;
}

// runtime-dispatched form for the small (non-unrolled) reduction kernels
static __device__ __noinline__ float apply_epilogue(float v, int m, int n, int N, int epilogue, const EpiParams& ep) {
  switch (epilogue) {
    case DASA_EPI_BIAS: return apply_epilogue_t<DASA_EPI_BIAS>(v, m, n, N, ep);
    case DASA_EPI_BIAS_TANH: return apply_epilogue_t<DASA_EPI_BIAS_TANH>(v, m, n, N, ep);
    case DASA_EPI_BIAS_GELU: return apply_epilogue_t<DASA_EPI_BIAS_GELU>(v, m, n, N, ep);
    case DASA_EPI_BIAS_RELU: return apply_epilogue_t<DASA_EPI_BIAS_RELU>(v, m, n, N, ep);
    case DASA_EPI_GATE: return apply_epilogue_t<DASA_EPI_GATE>(v, m, n, N, ep);
    case DASA_EPI_TANH: return apply_epilogue_t<DASA_EPI_TANH>(v, m, n, N, ep);
    default: return apply_epilogue_t<DASA_EPI_NONE>(v, m, n, N, ep);
  }
}

extern int64_t g_gemm_routes[DASA_ROUTE_COUNT];     // api.cu: per-route call counters (dasa_debug_gemm_route_counts)

// implemented in gemm_simt.cu / gemm_tc.cu
size_t dasa_gemm_simt_workspace(int M, int N, int K);
int dasa_gemm_simt(int a_kmajor, int b_kmajor, int M, int N, int K, float alpha, const float* A, int64_t lda,
                   const float* B, int64_t ldb, float beta, float* C, int64_t ldc, int epilogue, const EpiParams& ep,
                   void* workspace, size_t workspace_bytes, cudaStream_t st);
bool dasa_gemm_tc_supported(int a_kmajor, int b_kmajor, int M, int N, int K, const float* A, int64_t lda, const float* B,
                            int64_t ldb, const float* C, int64_t ldc);
size_t dasa_gemm_tc_workspace(int M, int N, int K);
int dasa_gemm_tc(int a_kmajor, int b_kmajor, int M, int N, int K, float alpha, const float* A, int64_t lda,
                 const float* B, int64_t ldb, float beta, float* C, int64_t ldc, int epilogue, const EpiParams& ep,
                 void* workspace, size_t workspace_bytes, cudaStream_t st);
// persistent CTA-pair (cta_group::2) kernel, gemm_tc2.cu: tile width (256 / 128) or 0 = keep the single-CTA kernel
int dasa_gemm_pair_plan(int M, int N, int K);
int dasa_gemm_tc_pair(int bn, int M, int N, int K, float alpha, const float* A, int64_t lda, const float* B, int64_t ldb, float beta,
                      float* C, int64_t ldc, int epilogue, const EpiParams& ep, cudaStream_t st);
// MN-major operand layouts on the pair kernel (no transposed copies for dX = dY.W and dW = dY^T.X), gemm_tc2.cu
bool dasa_gemm_pair_mn_supported(int a_kmajor, int b_kmajor, int M, int N, int K, const float* A, int64_t lda, const float* B,
                                 int64_t ldb, int epilogue);
int dasa_gemm_pair_mn_splits(int M, int N, int K);     // K splits (1 = none) the MN-major path uses for this shape
int dasa_gemm_tc_pair_mn(int a_kmajor, int b_kmajor, int M, int N, int K, float alpha, const float* A, int64_t lda, const float* B,
                         int64_t ldb, float beta, float* C, int64_t ldc, void* workspace, size_t workspace_bytes, cudaStream_t st);
// grouped (2 problems) / split-K launch of the pair kernel; returns the number of K splits actually used (>= 1) or a negative error
int dasa_gemm_tc_pair_grouped(int M, int N, int K, const float* const A[2], int64_t lda, const float* const B[2], int64_t ldb,
                              float* const C[2], int64_t ldc, int splits, int64_t split_stride, cudaStream_t st);
bool dasa_gemm_f16_pair_supported(int M, int N, int K);
// C = alpha * A^T B + beta * C, A [K][lda], B [K][ldb] fp16 (both MN-major), split-K through the workspace like dasa_gemm_tc_pair_mn
int dasa_gemm_tc_pair_mn_f16(int M, int N, int K, float alpha, const void* A, int64_t lda, const void* B, int64_t ldb, float beta,
                             float* C, int64_t ldc, void* workspace, size_t workspace_bytes, cudaStream_t st);
int dasa_gemm_tc_pair_f16(int M, int N, int K, const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc, int c_half,
                          int epilogue, const EpiParams& ep, cudaStream_t st);
int dasa_gemm_tc_pair_grouped2(int M0, int M1, int N, int K, const float* const A[2], int64_t lda, const float* const B[2], int64_t ldb,
                               float* const C[2], int64_t ldc, int splits, int64_t split_stride, cudaStream_t st);
// skinny (M <= 32) weight-streaming mma.sync TF32 kernel, gemm_skinny.cu
bool dasa_gemm_skinny_supported(int a_kmajor, int b_kmajor, int M, int N, int K, const float* A, int64_t lda, const float* B,
                                int64_t ldb);
int dasa_gemm_skinny(int M, int N, int K, float alpha, const float* A, int64_t lda, const float* B, int64_t ldb, float beta,
                     float* C, int64_t ldc, int epilogue, const EpiParams& ep, cudaStream_t st);
