// Fused LSTM-cell epilogue of the packed bi-LSTM recurrence (bilstm_packed.cu issues it, gemm_tc2.cu runs it inside the CTA-pair
// GEMM): parameters of one time step for both directions, and the output-dropout helper shared with the pointwise kernels.
#pragma once
#include <cuda_fp16.h>
#include "common.cuh"
#include "rng.cuh"

// Dropout on the layer output (r2rmodel.py:2357 `ctx = self.drop(ctx)`), fused into the only kernel that writes / reads `out`:
// keep flag of out[seq, l, c] from a mask tensor [R, L, 2H] or drawn in place (rng.cuh: stream byte (seq * L + l) * 2H + c).
// Saves a 292 MB read + write pass in each direction (and the 73 MB mask) at the benchmark geometry.
struct PkDrop { const uint8_t* mask; const unsigned long long* seed_dev; unsigned long long seed, base; uint32_t thr; int stream; float scale; };

__device__ __forceinline__ float4 pk_drop4(const PkDrop& dr, int64_t e, float4 v) {   // e: element index of v.x (multiple of 4)
  if (dr.mask != nullptr) {
    const uint32_t m = *reinterpret_cast<const uint32_t*>(dr.mask + e);
    v.x *= (m & 0xFFu) ? dr.scale : 0.f; v.y *= (m & 0xFF00u) ? dr.scale : 0.f;
    v.z *= (m & 0xFF0000u) ? dr.scale : 0.f; v.w *= (m & 0xFF000000u) ? dr.scale : 0.f;
  } else if (dr.stream) {
    DropStream ds;
    ds.mixed = mix_seed(dr.seed_dev ? dr.seed_dev[0] : dr.seed); ds.base = dr.base; ds.thr = dr.thr;
    const uint32_t k4 = stream_keep4(ds, (uint64_t)e >> 2);
    v.x *= (k4 & 1u) ? dr.scale : 0.f; v.y *= (k4 & 2u) ? dr.scale : 0.f;
    v.z *= (k4 & 4u) ? dr.scale : 0.f; v.w *= (k4 & 8u) ? dr.scale : 0.f;
  }
  return v;
}

// Gate nonlinearities of the fused cell on raw MUFU ex2 / rcp (the epilogue is 2 warps per scheduler: with libm's expf / tanhf
// - ~25 dependent instructions each, 20 evaluations per 4 hidden units - its instruction latency chain was 17 us per tile against a
// 6 us main loop). |error| <= 3e-7 absolute for both (ex2.approx: 2 ulp, rcp.approx: 1 ulp; tanh through exp(-2|x|), no overflow).
__device__ __forceinline__ float lstm_sigmoid(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + e));           // e = +inf -> 0
  return r;
}
__device__ __forceinline__ float lstm_tanh(float x) {
  float t, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fabsf(x) * -2.8853900817779268f));   // exp(-2|x|) in (0, 1]
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + t));
  return copysignf((1.f - t) * r, x);
}

// One recurrence step, both directions (index d). The GEMM computes gh = h_prev W_hh^T with the rows of W_hh INTERLEAVED
// (dasa_lstm_whh_interleave_f16): output column c of a 256-wide tile nt holds gate (c % 128) / 32 of hidden unit
// nt * 64 + (c / 128) * 32 + c % 32, so the 128 accumulator columns an epilogue warp drains are i, f, g, o of the same 32 units and
// the cell update happens on the accumulator, in registers: gh never reaches memory and the pointwise launch disappears.
// Row r of direction d: r < n[d] live (full update); n[d] <= r < n_next[d] joins at the next step with a zero state.
struct LstmEpi {
  const float* xp[2];        // [n, 4H] x W_ih^T rows of this step's block (standard gate-major columns)
  const float* b_ih[2]; const float* b_hh[2];
  const float* c_prev[2]; float* c_out[2];
  float* acts[2];            // [n, 4H] i, f, g, o after the nonlinearity (standard layout; the backward pass reads it)
  float* h_next[2];          // fp32 rows of the block the next step reads (nullptr: last step of this direction)
  __half* h16_next[2];       // the same rows as fp16: the A operand of the next step's GEMM
  float* h_fin[2]; float* c_fin[2];
  float* out; const int32_t* perm;
  int pos[2], n[2], n_next[2];
  int L, H;
  PkDrop drop;
};

// gemm_tc2.cu: one launch = the recurrent GEMMs of both directions (fp16 operands, tcgen05 kind::f16) + the cell update.
// A16[d]: [n[d], H] fp16 state rows, W16[d]: [4H, H] fp16 interleaved recurrent weights. H % 64 == 0.
int dasa_gemm_tc_pair_lstm(const __half* const A16[2], const __half* const W16[2], const LstmEpi& le, cudaStream_t st);
// gemm_tc2.cu: grouped (two directions) split-K GEMM on fp16 operands, fp32 partial sums; returns the number of K splits used.
int dasa_gemm_tc_pair_grouped2_f16(int M0, int M1, int N, int K, const __half* const A[2], int64_t lda, const __half* const B[2],
                                   int64_t ldb, float* const C[2], int64_t ldc, int splits, int64_t split_stride, cudaStream_t st);
