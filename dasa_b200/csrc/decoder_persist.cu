// Persistent decoder rollout (include/dasa_b200.h: dasa_decoder_rollout_fwd / _bwd): BAttnDecoderLSTM.forward
// (model.py:472-574) and its backward for T consecutive actions in ONE cooperative launch.
//
// Why: at B = 20 episodes every projection of the decoder is a weight-streaming problem (84 MB of fp32 weights per action
// for 0.8 GFLOP) and the per-op path spends ~250 us per action in ~28 dependent launches against a ~20 us streaming floor.
// Here one CTA per SM stays resident for the whole rollout; the phases of an action are separated by a device-wide barrier
// (one release-add + acquire-poll on an L2 word, ~1 us) instead of kernel boundaries.
//
//   GEMM phases   Y[b, n] = sum_k X[b, k] W[n, k]: a CTA owns 16 (or 32) rows of W and ALL of K; its 8 warps split K in
//                 32-wide chunks, each lane streams its W rows straight from HBM with 128-bit no-allocate loads into
//                 mma.sync.m16n8k8 TF32 A fragments (W rows = MMA M, the <= 32 episodes = MMA N), two chunks in flight per
//                 warp; X (<= 32 x K, produced by the previous phase on other SMs) is read through L2 (ld.global.cg). The
//                 K slices are folded across the warps in shared memory in a fixed order (deterministic), the epilogue
//                 (bias / tanh / the whole LSTM cell / dropout of the next operand) runs on the folded tile.
//   attention     a CTA owns (episode, channel slice): the [rows x slice] tile of the panorama / instruction context is
//                 bulk-async copied (TMA engine, mbarrier completion) into shared memory ONE ACTION AHEAD, so the 19 MB of
//                 per-action context stream under the GEMM phases; partial row dots go through an L2 scratch, the softmax
//                 (+ circular heading shift) is recomputed by every slice owner, the weighted sum comes from the resident tile.
//
// The k-index trick of gemm_skinny.cu is used throughout: a lane loads 4 consecutive k of its rows and uses component j in
// MMA step j for BOTH operands, so every global access is 128-bit and any bijection of the reduction index is a valid GEMM.
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>
#include "common.cuh"

namespace {

constexpr int DP_THREADS = 256;
constexpr int DP_WARPS = 8;
constexpr int DP_CHUNK = 32;         // k per chunk
constexpr int DP_MAXROWS = 128;      // attention rows (views / tokens)
constexpr int DP_MAXK = 15;          // shift kernel taps
constexpr int DP_NST2_H = 3;         // the same for the fp16-X block (its X fragments are half the registers)
constexpr int DP_NST2 = 2;           // chunks in flight per warp in the 32-row GEMM phases (register budget: 3 spill at 24 episodes); 3 in the 16-row ones

// ------------------------------------------------------------------------------------------------ device-wide barrier
struct GridBar {
  unsigned int* ctr;
  unsigned int target;
  unsigned int nblk;
};

// All CTAs are co-resident (cooperative launch). Thread 0 publishes the CTA's writes (bar.sync orders the other threads'
// writes before its gpu-scope release) and polls with acquire loads. Bounded: a protocol bug traps instead of hanging the box.
__device__ __forceinline__ void grid_sync(GridBar& gb) {
  gb.target += gb.nblk;
  __syncthreads();
  if (threadIdx.x == 0) {
    // bar.sync orders every thread's writes before thread 0's gpu-scope release (cumulativity); the acquire poll + the closing
    // bar.sync order them before every thread's later reads on the other SMs. Measured 1.26 us per barrier over 148 CTAs
    // against 1.68 us with __threadfence() on both sides (scripts/barrier_bench.py).
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(gb.ctr) : "memory");
    unsigned int v, it = 0;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(gb.ctr) : "memory");
      if (++it > (1u << 22)) __trap();
    } while ((int)(v - gb.target) < 0);
  }
  __syncthreads();
}

// The barrier in two halves: stores issued BEFORE grid_arrive are visible to every CTA after its grid_wait; stores issued between the
// halves (output rows nothing in this launch reads: dctx, dfeat) drain under the wait instead of in front of the arrive.
__device__ __forceinline__ void grid_arrive(GridBar& gb) {
  gb.target += gb.nblk;
  __syncthreads();
  if (threadIdx.x == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(gb.ctr) : "memory");
}
__device__ __forceinline__ void grid_wait(GridBar& gb) {
  if (threadIdx.x == 0) {
    unsigned int v, it = 0;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(gb.ctr) : "memory");
      if (++it > (1u << 22)) __trap();
    } while ((int)(v - gb.target) < 0);
  }
  __syncthreads();
}

// Phase timestamps of CTA 0 (SM clock) for the LAST launch: [0] = after the prologue barrier, then one per phase barrier of the
// first DP_TIMED_ACTIONS actions. Read with dasa_debug_decoder_phase_clocks (profiling only; ~20 clock reads per action).
constexpr int DP_TIMED_ACTIONS = 4;
__device__ long long g_dp_clock[1 + 8 * DP_TIMED_ACTIONS];
__device__ __forceinline__ void dp_stamp(int idx) {
  if (blockIdx.x == 0 && threadIdx.x == 0 && idx < 1 + 8 * DP_TIMED_ACTIONS) g_dp_clock[idx] = clock64();
}

// ------------------------------------------------------------------------------------------------------------ helpers
__device__ __forceinline__ uint32_t dp_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void dp_mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// data produced by OTHER SMs earlier in this launch: L2 is the point of coherence, never the (non-coherent) L1 / texture path
__device__ __forceinline__ float4 ld_cg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float ld_cg(const float* p) { return __ldcg(p); }
// Launders a pointer through an empty volatile asm: everything derived from it is computed AFTER this point. Without it the
// compiler hoists the per-lane row addresses of all eight GEMM phases out of the rollout loop and keeps them live across every
// phase (~100 registers): the K loops then spill.
template <typename T>
__device__ __forceinline__ T* dp_opaque(T* p) {
  asm volatile("" : "+l"(p));
  return p;
}
// component j of a float4 with j a compile-time constant after unrolling (no address-of: the fragments must stay in registers)
__device__ __forceinline__ float f4_get(const float4& v, int j) { return j == 0 ? v.x : (j == 1 ? v.y : (j == 2 ? v.z : v.w)); }

// fp16 with saturation to +-65504 instead of inf (the backward pass's scaled gradient copies: a clipped value, never a NaN)
__device__ __forceinline__ __half dp_half_sat(float x) {
  unsigned short r;
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(r) : "f"(x));
  return __ushort_as_half(r);
}
__device__ __forceinline__ __half2 dp_half2_sat(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return *reinterpret_cast<__half2*>(&r);
}

struct Slice { int c0, cn; };
// channel slice s of S over D channels, float4-granular
__device__ __host__ __forceinline__ Slice dp_slice(int D, int S, int s) {
  const int n4 = D >> 2;
  const int a = (int)(((int64_t)n4 * s) / S), b = (int)(((int64_t)n4 * (s + 1)) / S);
  Slice r; r.c0 = 4 * a; r.cn = 4 * (b - a);
  return r;
}
__host__ __device__ __forceinline__ int dp_chunk_pitch(int D, int S) { return (int)(((D >> 2) + S - 1) / S) * 4; }

// ---------------------------------------------------------------------------------------------------- GEMM building block
// One CTA-wide product: res[m][nl] = sum_k X[m, k] * W_row(nl)[k] for m < 8*MT (rows >= M read as zero), nl < 16*RT.
// W is stored as fp16: its 10-bit mantissa is exactly TF32's, so for |w| in [2^-14, 65504] the value the tensor core multiplies is
// bit-identical to cvt.rna.tf32(w) of the fp32 master weight (smaller |w| carry an absolute error <= 3e-8) while the weight
// stream per action halves to 42 MB, which stays resident in the 126 MB L2 from one action to the next.
// wrow[i] (i < 2*RT) = this lane's W row for local row g + 8*i, already offset by 4*t halves. K % 32 == 0.
// red: DP_WARPS*RT*MT*128 floats, res: 16*RT*8*MT floats (shared memory). Ends with a __syncthreads: res is complete.
typedef __half wt_t;
// 4 consecutive fp16 weights, read-only path, no L1 allocation, L2 evict_last: the 46 MB weight stream is re-read by every action
// of the rollout and should stay resident in the 126 MB L2 against the read-once panorama / instruction tiles (evict_first) -
// without the hints the weights were fetched from DRAM again in every action (ncu: 59 MB of DRAM reads per action, L2 hit 52 %)
__device__ __forceinline__ uint2 ldg_w4(const wt_t* p) {
  uint2 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.u32 {%0,%1}, [%2], %3;"
               : "=r"(r.x), "=r"(r.y) : "l"(p), "l"(l2_policy_evict_last()));
  return r;
}
// component j of 4 packed halves as a TF32 operand (the fp16 -> fp32 conversion is exact and already TF32-representable)
__device__ __forceinline__ uint32_t h4_tf32(const uint2& v, int j) {
  const uint32_t w = (j < 2) ? v.x : v.y;
  const unsigned short h = (j & 1) ? (unsigned short)(w >> 16) : (unsigned short)(w & 0xffffu);
  return __float_as_uint(__half2float(__ushort_as_half(h)));
}

template <int MT, int RT>
struct GemmFrag {
  uint2 w[2 * RT][2];
  float4 x[MT][2];
};

// half_last: the block has 16 * RT - 8 weight rows - the upper 8 rows of the last 16-row tile are neither loaded nor multiplied
// (their A fragments are zero). Lets a phase with N outputs use row blocks of 8 / 24 rows, so that all 148 CTAs stream weights
// (N = 1024: 64 blocks of 16 rows left 84 SMs idle; N = 3072: 192 blocks of 16 rows took two rounds on 44 of them).
template <int MT, int RT>
__device__ __forceinline__ void dp_load(GemmFrag<MT, RT>& f, const wt_t* (&wrow)[2 * RT], const float* (&xrow)[MT],
                                        const bool (&xok)[MT], int kc, const bool half_last) {
#pragma unroll
  for (int i = 0; i < 2 * RT; ++i) {
    if (half_last && i == 2 * RT - 1) continue;
    f.w[i][0] = ldg_w4(wrow[i] + kc);
    f.w[i][1] = ldg_w4(wrow[i] + kc + 16);
  }
#pragma unroll
  for (int i = 0; i < MT; ++i) {
    if (xok[i]) {
      f.x[i][0] = ld_cg4(xrow[i] + kc);
      f.x[i][1] = ld_cg4(xrow[i] + kc + 16);
    } else {
      f.x[i][0] = make_float4(0.f, 0.f, 0.f, 0.f);
      f.x[i][1] = f.x[i][0];
    }
  }
}

template <int MT, int RT>
__device__ __forceinline__ void dp_compute(float (&acc)[RT][MT][4], const GemmFrag<MT, RT>& f, const bool half_last) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint32_t bf[MT][2];
#pragma unroll
    for (int i = 0; i < MT; ++i) {
      bf[i][0] = dp_tf32(f4_get(f.x[i][0], j));
      bf[i][1] = dp_tf32(f4_get(f.x[i][1], j));
    }
#pragma unroll
    for (int h = 0; h < RT; ++h) {
      uint32_t af[4];
      const bool off = half_last && h == RT - 1;
      af[0] = h4_tf32(f.w[2 * h][0], j);                     // (row g,     k = t)
      af[1] = off ? 0u : h4_tf32(f.w[2 * h + 1][0], j);      // (row g + 8, k = t)
      af[2] = h4_tf32(f.w[2 * h][1], j);                     // (row g,     k = t + 4)
      af[3] = off ? 0u : h4_tf32(f.w[2 * h + 1][1], j);      // (row g + 8, k = t + 4)
#pragma unroll
      for (int i = 0; i < MT; ++i) dp_mma(acc[h][i], af, bf[i][0], bf[i][1]);
    }
  }
}

// fold the K slices of the 8 warps in warp order (deterministic): red[warp][h][i][e][lane] -> res[m * 16 RT + nl]. Ends with a
// __syncthreads: res is complete.
template <int MT, int RT>
__device__ __forceinline__ void dp_fold(const float (&acc)[RT][MT][4], float* red, float* res) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int h = 0; h < RT; ++h)
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) red[(((warp * RT + h) * MT + i) * 4 + e) * 32 + lane] = acc[h][i][e];
  __syncthreads();
  constexpr int NL = 16 * RT, OUTS = NL * 8 * MT;
  for (int o = threadIdx.x; o < OUTS; o += DP_THREADS) {
    const int nl = o % NL, m = o / NL;
    const int h = nl >> 4, r16 = nl & 15, i = m >> 3, mm = m & 7;
    const int e = (mm & 1) + (r16 >= 8 ? 2 : 0), ln = (r16 & 7) * 4 + (mm >> 1);
    float s = 0.f;
#pragma unroll
    for (int w2 = 0; w2 < DP_WARPS; ++w2) s += red[(((w2 * RT + h) * MT + i) * 4 + e) * 32 + ln];
    res[o] = s;                      // res[m * NL + nl]
  }
  __syncthreads();
}

// NST chunks of W / X in flight per warp (registers; the ring is fully unrolled so every fragment index is a constant)
template <int MT, int RT, int NST>
__device__ __forceinline__ void dp_gemm_block(const wt_t* (&wrow)[2 * RT], const float* X, int64_t ldx, int M, int K,
                                              float* red, float* res, const bool half_last = false) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const float* xrow[MT];
  bool xok[MT];
  X = dp_opaque(X);
#pragma unroll
  for (int i = 0; i < MT; ++i) {
    const int m = 8 * i + g;
    xok[i] = m < M;
    xrow[i] = X + (int64_t)(xok[i] ? m : 0) * ldx + 4 * t;
  }
  float acc[RT][MT][4];
#pragma unroll
  for (int h = 0; h < RT; ++h)
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[h][i][e] = 0.f;

  const int nchunks = K / DP_CHUNK;
  GemmFrag<MT, RT> f[NST];
  int c = warp;
#pragma unroll
  for (int s = 0; s < NST - 1; ++s)
    if (c + s * DP_WARPS < nchunks) dp_load<MT, RT>(f[s], wrow, xrow, xok, (c + s * DP_WARPS) * DP_CHUNK, half_last);
#pragma unroll 1
  for (; c < nchunks; c += NST * DP_WARPS) {
#pragma unroll
    for (int s = 0; s < NST; ++s) {
      const int cl = c + (s + NST - 1) * DP_WARPS;
      if (cl < nchunks) dp_load<MT, RT>(f[(s + NST - 1) % NST], wrow, xrow, xok, cl * DP_CHUNK, half_last);
      if (c + s * DP_WARPS < nchunks) dp_compute<MT, RT>(acc, f[s], half_last);
    }
  }

  dp_fold<MT, RT>(acc, red, res);
}

// ---- the same product with X stored as fp16 (forward pass: every GEMM operand of the decoder is O(1) - tanh / sigmoid outputs,
// dropout-scaled states, attention-weighted features - so its fp16 copy carries exactly the 11 significant bits the TF32 tensor
// core would have kept, at half the bytes). X is read by EVERY CTA in every GEMM phase (148 x 20 x K x 4 B: 99 MB per action against
// 42 MB of weights), so halving it is the larger part of the phase's L2 traffic. mma.sync.m16n8k16 f16 x f16 -> f32: a lane
// loads 8 consecutive k (one 128-bit access) of each of its W rows and of its X row per 32-wide chunk and feeds halves
// (4j, 4j+1) / (4j+2, 4j+3) to k slots (2t, 2t+1) / (2t+8, 2t+9) of MMA step j for BOTH operands (any bijection of the reduction
// index is a valid GEMM). wrow[i] / X rows are offset by 8 * t halves.
typedef __half xh_t;
__device__ __forceinline__ uint4 ldg_w8(const wt_t* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p), "l"(l2_policy_evict_last()));
  return r;
}
__device__ __forceinline__ uint32_t u4_get(const uint4& v, int j) { return j == 0 ? v.x : (j == 1 ? v.y : (j == 2 ? v.z : v.w)); }
__device__ __forceinline__ void dp_mma_h(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <int MT, int RT>
struct GemmFragH {
  uint4 w[2 * RT];
  uint4 x[MT];
};
template <int MT, int RT>
__device__ __forceinline__ void dp_load_h(GemmFragH<MT, RT>& f, const wt_t* (&wrow)[2 * RT], const xh_t* (&xrow)[MT],
                                          const bool (&xok)[MT], int kc, const bool half_last) {
#pragma unroll
  for (int i = 0; i < 2 * RT; ++i) {
    if (half_last && i == 2 * RT - 1) continue;
    f.w[i] = ldg_w8(wrow[i] + kc);
  }
#pragma unroll
  for (int i = 0; i < MT; ++i)
    f.x[i] = xok[i] ? __ldcg(reinterpret_cast<const uint4*>(xrow[i] + kc)) : make_uint4(0u, 0u, 0u, 0u);
}
template <int MT, int RT>
__device__ __forceinline__ void dp_compute_h(float (&acc)[RT][MT][4], const GemmFragH<MT, RT>& f, const bool half_last) {
#pragma unroll
  for (int j = 0; j < 2; ++j) {
#pragma unroll
    for (int h = 0; h < RT; ++h) {
      uint32_t af[4];
      const bool off = half_last && h == RT - 1;
      af[0] = u4_get(f.w[2 * h], 2 * j);
      af[1] = off ? 0u : u4_get(f.w[2 * h + 1], 2 * j);
      af[2] = u4_get(f.w[2 * h], 2 * j + 1);
      af[3] = off ? 0u : u4_get(f.w[2 * h + 1], 2 * j + 1);
#pragma unroll
      for (int i = 0; i < MT; ++i) dp_mma_h(acc[h][i], af, u4_get(f.x[i], 2 * j), u4_get(f.x[i], 2 * j + 1));
    }
  }
}
template <int MT, int RT, int NST>
__device__ __forceinline__ void dp_gemm_block_h(const wt_t* (&wrow)[2 * RT], const xh_t* X, int64_t ldx, int M, int K,
                                                float* red, float* res, const bool half_last = false) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const xh_t* xrow[MT];
  bool xok[MT];
  X = dp_opaque(X);
#pragma unroll
  for (int i = 0; i < MT; ++i) {
    const int m = 8 * i + g;
    xok[i] = m < M;
    xrow[i] = X + (int64_t)(xok[i] ? m : 0) * ldx + 8 * t;
  }
  float acc[RT][MT][4];
#pragma unroll
  for (int h = 0; h < RT; ++h)
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[h][i][e] = 0.f;

  const int nchunks = K / DP_CHUNK;
  // Every CTA reads the WHOLE X operand; started at chunk 0 everywhere, all 148 CTAs asked the same few L2 lines for the same
  // chunk at the same moment. Each CTA walks the reduction index from its own rotation instead (any bijection of k is the same
  // sum up to fp32 ordering, and the order is a fixed function of the CTA index: deterministic). Staging X through shared memory
  // with bulk copies (2 x 512-k pieces in flight, all that fits next to the attention tiles) was tried and is slower: 10 -> 17 us
  // for P3, the copies' latency is exposed once per piece.
  const int rot = (int)((blockIdx.x * 11u) % (unsigned)nchunks);
  auto kc_of = [&](int ci) { int q = ci + rot; q = q >= nchunks ? q - nchunks : q; return q * DP_CHUNK; };
  GemmFragH<MT, RT> f[NST];
  int c = warp;
#pragma unroll
  for (int s = 0; s < NST - 1; ++s)
    if (c + s * DP_WARPS < nchunks) dp_load_h<MT, RT>(f[s], wrow, xrow, xok, kc_of(c + s * DP_WARPS), half_last);
#pragma unroll 1
  for (; c < nchunks; c += NST * DP_WARPS) {
#pragma unroll
    for (int s = 0; s < NST; ++s) {
      const int cl = c + (s + NST - 1) * DP_WARPS;
      if (cl < nchunks) dp_load_h<MT, RT>(f[(s + NST - 1) % NST], wrow, xrow, xok, kc_of(cl), half_last);
      if (c + s * DP_WARPS < nchunks) dp_compute_h<MT, RT>(acc, f[s], half_last);
    }
  }
  dp_fold<MT, RT>(acc, red, res);
}

// plain row block: rows n0 .. n0+15 (n0 .. n0+7 with rows8: the upper half is never loaded) of W, clamped to N-1; the caller
// never stores rows >= N
// kper: consecutive k a lane owns per chunk (4: fp32-X block, two 64-bit accesses per row and chunk; 8: fp16-X block, one 128-bit)
__device__ __forceinline__ void dp_rows16(const wt_t* (&wrow)[2], const wt_t* W, int64_t ldw, int n0, int N, const bool rows8 = false,
                                          const int kper = 4) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    int n = n0 + g + ((rows8 || i == 0) ? 0 : 8);
    n = n < N ? n : N - 1;
    wrow[i] = W + (int64_t)n * ldw + kper * t;
  }
}

// --------------------------------------------------------------------------------------------- attention building blocks
struct AttnSmem {
  float* tile;       // [rows][pitch]
  float* tv;         // [pitch] target slice
  float* dv;         // [pitch] (bwd) upstream gradient slice
  float* zrow;       // [DP_MAXROWS] logits / dq
  float* prow;       // [DP_MAXROWS] softmax
  float* wrow;       // [DP_MAXROWS] final weights (shifted q or alpha)
  float* aux;        // [DP_MAXROWS] dz (bwd)
  float* aux2;       // [DP_MAXROWS] dp (bwd)
  float* kap;        // [16]
  float* colred;     // [8 * 128 * 4] column partial sums of the weighted sum
  uint8_t* mk;       // [DP_MAXROWS] context padding flags of this CTA's episode
  uint64_t* bar;
};

// stage the unmasked rows of this CTA's channel slice with bulk-async copies (warp 0), completion on s.bar
__device__ __forceinline__ void dp_issue_tile(float* tile, int pitch, uint64_t* bar, const float* src_b, int64_t ld_row, int rows,
                                              int c0, int cn, const uint8_t* mask_b) {
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    int mine = 0;
    for (int r = lane; r < rows; r += 32) mine += (mask_b == nullptr || mask_b[r] == 0) ? 1 : 0;
    const int nvalid = (int)__reduce_add_sync(0xffffffffu, (unsigned)mine);
    if (lane == 0) mbar_expect_tx(bar, (uint32_t)nvalid * (uint32_t)cn * 4u);
    __syncwarp();
    if (cn > 0)
      for (int r = lane; r < rows; r += 32)
        if (mask_b == nullptr || mask_b[r] == 0)
          bulk_g2s_hint(tile + (size_t)r * pitch, src_b + (int64_t)r * ld_row + c0, (uint32_t)cn * 4u, bar, l2_policy_evict_first());
  }
}

// partial row dots of the resident slice against vec (shared memory): zp[r] = tile[r, :cn] . vec ; masked rows -> 0.
// A warp works on 4 rows at once (8 lanes per row, 3 shuffle steps): with one row per warp iteration the 5-step shuffle chains of
// the 10 rows a warp owns (80 tokens) ran back to back, ~1.5 us of pure latency per attention.
__device__ __forceinline__ void dp_partial_dots(const float* tile, int pitch, const float* vec, int rows, int cn,
                                                const uint8_t* mask_b, float* zp) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int rq = lane >> 3, cl = lane & 7;
  const int n4 = cn >> 2;
  const float4* v4 = reinterpret_cast<const float4*>(vec);
  for (int r0 = wid * 4; r0 < rows; r0 += DP_WARPS * 4) {
    const int r = r0 + rq;
    float acc = 0.f;
    if (r < rows && (mask_b == nullptr || mask_b[r] == 0)) {
      const float4* row = reinterpret_cast<const float4*>(tile + (size_t)r * pitch);
      for (int j = cl; j < n4; j += 8) {
        const float4 a = row[j], b = v4[j];
        acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
      }
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    if (cl == 0 && r < rows) zp[r] = acc;
  }
}

// out[c0 + 4*col ..] = sum_r w[r] * tile[r, col]  over unmasked rows (w in shared memory); out may be any global row
// out16 (optional): the same values as fp16 at the same element offsets (the forward pass's GEMM operand copies)
__device__ __forceinline__ void dp_weighted_sum(const AttnSmem& s, int pitch, int rows, int cn, const uint8_t* mask_b, const float* w,
                                                float* out, __half* out16 = nullptr, const float s16 = 1.f) {
  const int n4 = cn >> 2;
  if (n4 <= 0) return;
  const int G = n4 >= DP_THREADS ? 1 : min(DP_THREADS / n4, 8);
  const int passes = (n4 + DP_THREADS - 1) / DP_THREADS;
  for (int pass = 0; pass < passes; ++pass) {
    const int col = (G > 1) ? (int)(threadIdx.x % n4) : (int)threadIdx.x + pass * DP_THREADS;
    const int grp = (G > 1) ? (int)(threadIdx.x / n4) : 0;
    const bool act = (col < n4) && (grp < G);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (act) {
      const float4* tp = reinterpret_cast<const float4*>(s.tile) + col;
      const int p4 = pitch >> 2;
      for (int r = grp; r < rows; r += G) {
        if (mask_b != nullptr && mask_b[r] != 0) continue;
        const float wr = w[r];
        const float4 v = tp[(size_t)r * p4];
        acc.x = fmaf(wr, v.x, acc.x); acc.y = fmaf(wr, v.y, acc.y); acc.z = fmaf(wr, v.z, acc.z); acc.w = fmaf(wr, v.w, acc.w);
      }
    }
    if (G > 1) {
      __syncthreads();
      if (act) reinterpret_cast<float4*>(s.colred)[grp * n4 + col] = acc;
      __syncthreads();
      if (act && grp == 0)
        for (int g2 = 1; g2 < G; ++g2) {
          const float4 o = reinterpret_cast<const float4*>(s.colred)[g2 * n4 + col];
          acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
        }
    }
    if (act && grp == 0) {
      *reinterpret_cast<float4*>(out + 4 * col) = acc;
      if (out16 != nullptr) {
        __half2 h[2] = {dp_half2_sat(acc.x * s16, acc.y * s16), dp_half2_sat(acc.z * s16, acc.w * s16)};
        *reinterpret_cast<uint2*>(out16 + 4 * col) = *reinterpret_cast<uint2*>(h);
      }
    }
  }
}

// Sum of the S (<= 8) slice partials of row r in slice order (deterministic). All S loads are issued before the first add: written
// as `v += ld_cg(..)` in a loop over the run-time S the compiler serialised them - 7 dependent L2 round trips (~5 us) per row, which
// was most of the softmax phases and 10 of the 16 us of B2b.
__device__ __forceinline__ float dp_sum_partials(const float* zp, int S, int r) {
  float t[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) t[k] = (k < S) ? ld_cg(zp + (size_t)k * DP_MAXROWS + r) : 0.f;
  float v = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k)
    if (k < S) v += t[k];
  return v;
}

// z[r] = sum over the S slice partials (fixed order; one row per thread), then warp 0: masked softmax -> s.prow (0 for masked
// rows). Ends with the softmax visible to warp 0 only: callers __syncthreads() before other warps read s.prow.
__device__ __forceinline__ void dp_softmax_rows(const AttnSmem& s, const float* zpart_b, int S, int rows, const uint8_t* mask_b) {
  for (int r = threadIdx.x; r < DP_MAXROWS; r += DP_THREADS) {
    float v = -INFINITY;
    if (r < rows && (mask_b == nullptr || mask_b[r] == 0)) v = dp_sum_partials(zpart_b, S, r);
    s.zrow[r] = v;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    float z[DP_MAXROWS / 32];
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < DP_MAXROWS / 32; ++i) {
      z[i] = s.zrow[lane + 32 * i];
      mx = fmaxf(mx, z[i]);
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < DP_MAXROWS / 32; ++i) {
      z[i] = (z[i] == -INFINITY) ? 0.f : expf(z[i] - mx);
      sum += z[i];
    }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
#pragma unroll
    for (int i = 0; i < DP_MAXROWS / 32; ++i) {
      const int r = lane + 32 * i;
      if (r < rows) s.prow[r] = z[i] * inv;
    }
  }
}

struct SmemPlan {
  int pitchF, pitchC;          // floats per tile row
  size_t off_red, off_res, off_tileF, off_tileC, off_tv, off_dv, off_small, total;
};

__host__ __device__ inline SmemPlan dp_plan(int MT, int V, int L, int F, int D, int S) {
  SmemPlan p;
  p.pitchF = dp_chunk_pitch(F, S);
  p.pitchC = dp_chunk_pitch(D, S);
  size_t o = 0;
  p.off_red = o; o += sizeof(float) * (size_t)DP_WARPS * 2 * MT * 128;
  p.off_res = o; o += sizeof(float) * (size_t)32 * 8 * MT;
  p.off_tileF = o; o += sizeof(float) * (size_t)V * p.pitchF;
  p.off_tileC = o; o += sizeof(float) * (size_t)L * p.pitchC;
  const int pm = p.pitchF > p.pitchC ? p.pitchF : p.pitchC;
  p.off_tv = o; o += sizeof(float) * (size_t)pm;
  p.off_dv = o; o += sizeof(float) * (size_t)pm;
  p.off_small = o; o += sizeof(float) * (5 * DP_MAXROWS + 16 + 8 * 128 * 4) + 32 + DP_MAXROWS;
  p.total = (o + 127) & ~(size_t)127;
  return p;
}

__device__ __forceinline__ AttnSmem dp_attn_smem(unsigned char* raw, const SmemPlan& pl, bool feat) {
  float* small = reinterpret_cast<float*>(raw + pl.off_small);
  AttnSmem s;
  s.tv = reinterpret_cast<float*>(raw + pl.off_tv);
  s.dv = reinterpret_cast<float*>(raw + pl.off_dv);
  s.zrow = small; s.prow = small + DP_MAXROWS; s.wrow = small + 2 * DP_MAXROWS; s.aux = small + 3 * DP_MAXROWS;
  s.aux2 = small + 4 * DP_MAXROWS;
  s.kap = small + 5 * DP_MAXROWS;
  s.colred = s.kap + 16;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s.colred + 8 * 128 * 4);
  s.mk = reinterpret_cast<uint8_t*>(bars + 4);          // padding flags of this CTA's episode (constant over the rollout)
  s.tile = reinterpret_cast<float*>(raw + (feat ? pl.off_tileF : pl.off_tileC));
  s.bar = bars + (feat ? 0 : 1);
  return s;
}

// which (episode, channel slice) this CTA owns in the attention phases
struct AttnOwner {
  bool on;
  int b, s;
  Slice sl;
  const uint8_t* mask_b;
};
// mask_s: the shared-memory copy of this CTA's episode padding flags (dp_stage_mask), or nullptr for "no mask"
__device__ __forceinline__ AttnOwner dp_owner(int B, int S, int width, const uint8_t* mask_s) {
  AttnOwner o;
  o.on = (int)blockIdx.x < B * S;
  o.b = o.on ? (int)blockIdx.x / S : 0;
  o.s = o.on ? (int)blockIdx.x % S : 0;
  o.sl = dp_slice(width, S, o.s);
  o.mask_b = o.on ? mask_s : nullptr;
  return o;
}
// once per launch: the padding flags of this CTA's episode into shared memory (they were re-read from global memory, one
// dependent L2 round trip per row, in every attention phase of every action)
__device__ __forceinline__ void dp_stage_mask(unsigned char* raw, const SmemPlan& pl, int B, int S, int L, const uint8_t* mask,
                                              int64_t mask_ld) {
  const AttnSmem sc = dp_attn_smem(raw, pl, false);
  const bool on = (int)blockIdx.x < B * S;
  const int b = on ? (int)blockIdx.x / S : 0;
  for (int r = threadIdx.x; r < DP_MAXROWS; r += DP_THREADS) sc.mk[r] = (on && mask != nullptr && r < L) ? mask[(int64_t)b * mask_ld + r] : 0;
}
__device__ __forceinline__ const uint8_t* dp_mask_smem(unsigned char* raw, const SmemPlan& pl, const uint8_t* mask) {
  return mask != nullptr ? dp_attn_smem(raw, pl, false).mk : nullptr;
}

// fp16 copies of the forward GEMM operands inside the caller's x16 scratch: drop(h~_{t-1}) [T,B,H] | [emb ; attn ; h~] [T,B,KX] |
// [wc ; drop(h_1)] [T,B,DC]
struct FwdX16 { __half* hp; __half* xh; __half* cat; };
__device__ __forceinline__ FwdX16 fwd_x16(const dasa_decoder_fwd_t& a) {
  FwdX16 x;
  const int64_t TB = (int64_t)a.T * a.B;
  x.hp = reinterpret_cast<__half*>(a.x16);
  x.xh = x.hp + TB * a.H;
  x.cat = x.xh + TB * (a.E + a.F + a.H);
  return x;
}

// ---- the GEMM phases (one function each)

// The three 16-row GEMM phases of an action share ONE instance of the K loop (the phase picks operands and epilogue at run
// time): with one inlined copy per phase the kernel grew to ~90k instructions and the compiler stopped keeping the operand
// fragments in registers.
template <int MT>
__device__ __forceinline__ void fwd_gemm16(const dasa_decoder_fwd_t& a, const int t, const int ph, float* red, float* res) {
  const int T = a.T, B = a.B, H = a.H, E = a.E, F = a.F, D = a.D, NK = a.NK;
  const int KX = E + F + H, DC = D + H;
  const int tid = threadIdx.x;
  const int64_t tb = (int64_t)t * B;
  const float scale = a.drop_scale;
  const wt_t* W; const xh_t* X;
  int64_t ldw, ldx;
  int N, K;
  const FwdX16 x16 = fwd_x16(a);
  if (ph == 0)      { W = reinterpret_cast<const wt_t*>(a.w_feat);    ldw = H;  N = NK; K = H;  X = x16.hp + tb * H;        ldx = H; }    // P1: tk = W_feat drop(h~) + b
  else if (ph == 4) { W = reinterpret_cast<const wt_t*>(a.w_att_in);  ldw = H;  N = D;  K = H;  X = x16.cat + tb * DC + D;  ldx = DC; }   // P4: t2 = W_att_in drop(h_1)
  else              { W = reinterpret_cast<const wt_t*>(a.w_att_out); ldw = DC; N = H;  K = DC; X = x16.cat + tb * DC;      ldx = DC; }   // P6: h~ = tanh(W_att_out [wc ; drop(h_1)])
  W = dp_opaque(W);
  const bool rows8 = N <= 8 * (int)gridDim.x;                  // P6 (N = H = 1024): 128 blocks of 8 rows, one per CTA
  const int rb = rows8 ? 8 : 16;
  const int nitems = (N + rb - 1) / rb;
  for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
    const wt_t* wrow[2];
    dp_rows16(wrow, W, ldw, item * rb, N, rows8, 8);
    dp_gemm_block_h<MT, 1, 3>(wrow, X, ldx, B, K, red, res, rows8);
    for (int o = tid; o < 16 * 8 * MT; o += DP_THREADS) {
      const int m = o >> 4, n = item * rb + (o & 15);
      if (m >= B || n >= N || (o & 15) >= rb) continue;
      if (ph == 0) {
        a.tk[(tb + m) * NK + n] = res[o] + __ldg(a.b_feat + n);
      } else if (ph == 4) {
        a.t2[(tb + m) * D + n] = res[o];
      } else {
        const float v = tanhf(res[o]);
        a.htilde[(tb + m) * H + n] = v;
        if (t + 1 < T) {
          a.xh[(tb + B + m) * KX + E + F + n] = v;
          x16.xh[(tb + B + m) * KX + E + F + n] = __float2half_rn(v);
          const int64_t mi = (tb + B + m) * H + n;
          const float vd = a.m_hprev ? (a.m_hprev[mi] ? v * scale : 0.f) : v;
          a.hprev_drop[mi] = vd;
          x16.hp[mi] = __float2half_rn(vd);
        }
      }
    }
  }
}

template <int MT>
__device__ __forceinline__ void fwd_p3(const dasa_decoder_fwd_t& a, const int t, float* red, float* res) {
  const int T = a.T, B = a.B, H = a.H, E = a.E, F = a.F, D = a.D, NK = a.NK;
  const int KX = E + F + H, DC = D + H;
  const int tid = threadIdx.x;
  const int64_t tb = (int64_t)t * B;
  const float scale = a.drop_scale;
  (void)T; (void)NK; (void)KX; (void)DC; (void)scale; (void)tb; (void)tid;
    {
      const FwdX16 x16 = fwd_x16(a);
      const xh_t* X = x16.xh + tb * KX;
      const int nitems = H / 8;
      const int lane = tid & 31, g = lane >> 2, tq = lane & 3;
      for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
        const wt_t* wrow[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) wrow[i] = dp_opaque(reinterpret_cast<const wt_t*>(a.w_lstm)) + (int64_t)(i * H + item * 8 + g) * KX + 8 * tq;
        dp_gemm_block_h<MT, 2, DP_NST2_H>(wrow, X, KX, B, KX, red, res);
        for (int o = tid; o < 8 * MT * 8; o += DP_THREADS) {
          const int m = o >> 3, j = o & 7, u = item * 8 + j;
          if (m >= B) continue;
          float gt[4];
#pragma unroll
          for (int qg = 0; qg < 4; ++qg) gt[qg] = res[m * 32 + qg * 8 + j] + __ldg(a.b_ih + qg * H + u) + __ldg(a.b_hh + qg * H + u);
          const float ig = sigmoidf_(gt[0]), fg = sigmoidf_(gt[1]), gg = tanhf(gt[2]), og = sigmoidf_(gt[3]);
          const float cp = ld_cg(a.c + (tb + m) * H + u);
          const float c1 = fg * cp + ig * gg;
          const float h1 = og * tanhf(c1);
          float* ac = a.acts + (tb + m) * 4 * H;
          ac[u] = ig; ac[H + u] = fg; ac[2 * H + u] = gg; ac[3 * H + u] = og;
          a.c[(tb + B + m) * H + u] = c1;
          a.h1[(tb + m) * H + u] = h1;
          const int64_t mi = (tb + m) * H + u;
          const float h1d = a.m_h1 ? (a.m_h1[mi] ? h1 * scale : 0.f) : h1;
          a.cat[(tb + m) * DC + D + u] = h1d;
          x16.cat[(tb + m) * DC + D + u] = __float2half_rn(h1d);
        }
      }
    }
}

// ---- attention phases: NOT inlined, so that none of their state is live across the register-hungry GEMM phases
__device__ __noinline__ void fwd_issue_feat(const dasa_decoder_fwd_t& a, const SmemPlan& pl, unsigned char* raw, int S, int t) {
  const AttnOwner o = dp_owner(a.B, S, a.F, nullptr);
  if (!o.on) return;
  const AttnSmem sf = dp_attn_smem(raw, pl, true);
  dp_issue_tile(sf.tile, pl.pitchF, sf.bar, a.feat + (int64_t)t * a.feat_ld_t + (int64_t)o.b * a.feat_ld_b, a.feat_ld_row, a.V,
                o.sl.c0, o.sl.cn, nullptr);
}
__device__ __noinline__ void fwd_issue_ctx(const dasa_decoder_fwd_t& a, const SmemPlan& pl, unsigned char* raw, int S, int t) {
  const AttnOwner o = dp_owner(a.B, S, a.D, dp_mask_smem(raw, pl, a.ctx_mask));
  if (!o.on) return;
  const AttnSmem sc = dp_attn_smem(raw, pl, false);
  dp_issue_tile(sc.tile, pl.pitchC, sc.bar, a.ctx + (int64_t)t * a.ctx_ld_t + (int64_t)o.b * a.ctx_ld_b, a.ctx_ld_row, a.L,
                o.sl.c0, o.sl.cn, o.mask_b);
}

// P2a: partial view logits of this CTA's channel slice
__device__ __noinline__ void fwd_p2a(const dasa_decoder_fwd_t& a, const SmemPlan& pl, unsigned char* raw, int S, int t) {
  const AttnOwner o = dp_owner(a.B, S, a.F, nullptr);
  if (!o.on) return;
  const AttnSmem sf = dp_attn_smem(raw, pl, true);
  const int tid = threadIdx.x;
  const int64_t tb = (int64_t)t * a.B;
  const float* tk_b = a.tk + (tb + o.b) * a.NK;
  for (int c = tid; c < o.sl.cn; c += DP_THREADS) sf.tv[c] = ld_cg(tk_b + o.sl.c0 + c);
  __syncthreads();
  mbar_wait(sf.bar, (uint32_t)(t & 1));
  dp_partial_dots(sf.tile, pl.pitchF, sf.tv, a.V, o.sl.cn, nullptr, a.zpart + ((size_t)o.b * S + o.s) * DP_MAXROWS);
}

// P2b: softmax over the views, circular heading shift, weighted sum -> xh[:, E + slice]; then prefetch the next action's tile
__device__ __noinline__ void fwd_p2b(const dasa_decoder_fwd_t& a, const SmemPlan& pl, unsigned char* raw, int S, int t) {
  const AttnOwner o = dp_owner(a.B, S, a.F, nullptr);
  if (!o.on) return;
  const AttnSmem sf = dp_attn_smem(raw, pl, true);
  const int tid = threadIdx.x, V = a.V, F = a.F, E = a.E, NK = a.NK, KX = a.E + a.F + a.H;
  const int64_t tb = (int64_t)t * a.B;
  dp_softmax_rows(sf, a.zpart + (size_t)o.b * S * DP_MAXROWS, S, V, nullptr);
  if (tid < 32) {
    const int k = a.shift_k, lane = tid;
    const float kl = (lane < k) ? ld_cg(a.tk + (tb + o.b) * NK + F + lane) : -INFINITY;
    const float kmx = warp_max(kl);
    const float e = (lane < k) ? expf(kl - kmx) : 0.f;
    const float ksum = warp_sum(e);
    if (lane < k) {
      sf.kap[lane] = e / ksum;
      if (o.s == 0) a.kappa[(tb + o.b) * k + lane] = e / ksum;
    }
  }
  __syncthreads();
  {
    const int k = a.shift_k, half = k / 2, Hn = a.headings;
    for (int r = tid; r < V; r += DP_THREADS) {
      const int e = r / Hn, l = r % Hn;
      float qv = 0.f;
      for (int j = 0; j < k; ++j) {
        int src = (l + j - half) % Hn;
        if (src < 0) src += Hn;
        qv = fmaf(sf.kap[j], sf.prow[e * Hn + src], qv);
      }
      sf.wrow[r] = qv;
      if (o.s == 0) {
        a.p[(tb + o.b) * V + r] = sf.prow[r];
        a.q[(tb + o.b) * V + r] = qv;
      }
    }
  }
  __syncthreads();
  dp_weighted_sum(sf, pl.pitchF, V, o.sl.cn, nullptr, sf.wrow, a.xh + (tb + o.b) * KX + E + o.sl.c0,
                  fwd_x16(a).xh + (tb + o.b) * KX + E + o.sl.c0);
  // the next action's tile is prefetched at the top of the NEXT phase (after the barrier): every thread is done with this one
}

// P5a: partial token logits
__device__ __noinline__ void fwd_p5a(const dasa_decoder_fwd_t& a, const SmemPlan& pl, unsigned char* raw, int S, int t) {
  const AttnOwner o = dp_owner(a.B, S, a.D, dp_mask_smem(raw, pl, a.ctx_mask));
  if (!o.on) return;
  const AttnSmem sc = dp_attn_smem(raw, pl, false);
  const int tid = threadIdx.x;
  const int64_t tb = (int64_t)t * a.B;
  const float* t2_b = a.t2 + (tb + o.b) * a.D;
  for (int c = tid; c < o.sl.cn; c += DP_THREADS) sc.tv[c] = ld_cg(t2_b + o.sl.c0 + c);
  __syncthreads();
  mbar_wait(sc.bar, (uint32_t)(t & 1));
  dp_partial_dots(sc.tile, pl.pitchC, sc.tv, a.L, o.sl.cn, o.mask_b, a.zpart + ((size_t)o.b * S + o.s) * DP_MAXROWS);
}

// P5b: masked softmax over the tokens, weighted context -> cat[:, slice]; then prefetch the next action's tile
__device__ __noinline__ void fwd_p5b(const dasa_decoder_fwd_t& a, const SmemPlan& pl, unsigned char* raw, int S, int t) {
  const AttnOwner o = dp_owner(a.B, S, a.D, dp_mask_smem(raw, pl, a.ctx_mask));
  if (!o.on) return;
  const AttnSmem sc = dp_attn_smem(raw, pl, false);
  const int tid = threadIdx.x, L = a.L, DC = a.D + a.H;
  const int64_t tb = (int64_t)t * a.B;
  dp_softmax_rows(sc, a.zpart + (size_t)o.b * S * DP_MAXROWS, S, L, o.mask_b);
  __syncthreads();
  if (o.s == 0)
    for (int r = tid; r < L; r += DP_THREADS) a.alpha[(tb + o.b) * L + r] = sc.prow[r];
  dp_weighted_sum(sc, pl.pitchC, L, o.sl.cn, o.mask_b, sc.prow, a.cat + (tb + o.b) * DC + o.sl.c0,
                  fwd_x16(a).cat + (tb + o.b) * DC + o.sl.c0);
}

__device__ __noinline__ void fwd_prologue(const dasa_decoder_fwd_t& a) {
  const int B = a.B, H = a.H, E = a.E, KX = a.E + a.F + a.H;
  const int gtid = blockIdx.x * DP_THREADS + threadIdx.x, gthreads = gridDim.x * DP_THREADS;
  const float scale = a.drop_scale;
  const FwdX16 x16 = fwd_x16(a);
  // recurrent state of action 0, action embeddings of every action into the [x ; h] rows (fp32 for the backward pass, fp16 for
  // the GEMM phases)
  for (int i = gtid; i < B * H; i += gthreads) {
    const int b = i / H, n = i % H;
    const float h = __ldg(a.h0 + i);
    a.xh[(int64_t)b * KX + E + a.F + n] = h;
    x16.xh[(int64_t)b * KX + E + a.F + n] = __float2half_rn(h);
    const float hd = a.m_hprev ? (a.m_hprev[i] ? h * scale : 0.f) : h;
    a.hprev_drop[i] = hd;
    x16.hp[i] = __float2half_rn(hd);
    a.c[i] = __ldg(a.c0 + i);
  }
  for (int i = gtid; i < a.T * B * E; i += gthreads) {
    const int r = i / E, e = i % E;
    const float v = __ldg(a.emb + i);
    a.xh[(int64_t)r * KX + e] = v;
    x16.xh[(int64_t)r * KX + e] = __float2half_rn(v);
  }
}

// ============================================================================================================== FORWARD
template <int MT>
__global__ void __launch_bounds__(DP_THREADS, 1) decoder_rollout_fwd_kernel(const __grid_constant__ dasa_decoder_fwd_t a, const int S,
                                                                            const __grid_constant__ SmemPlan pl) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* red = reinterpret_cast<float*>(smem_raw + pl.off_red);
  float* res = reinterpret_cast<float*>(smem_raw + pl.off_res);
  GridBar gb{a.barrier, 0u, gridDim.x};
  if (threadIdx.x == 0) {
    const AttnSmem sf = dp_attn_smem(smem_raw, pl, true), sc = dp_attn_smem(smem_raw, pl, false);
    mbar_init(sf.bar, 1); mbar_init(sc.bar, 1); mbar_fence_init();
  }
  dp_stage_mask(smem_raw, pl, a.B, S, a.L, a.ctx_mask, a.ctx_mask_ld);
  __syncthreads();
  fwd_issue_feat(a, pl, smem_raw, S, 0);          // the tiles of action 0 stream in under the prologue and P1
  fwd_issue_ctx(a, pl, smem_raw, S, 0);
  fwd_prologue(a);
  grid_sync(gb);
  dp_stamp(0);
  for (int t = 0; t < a.T; ++t) {
#pragma unroll 1
    for (int ph = 0; ph < 8; ++ph) {
      // prefetch the next action's tiles at the top of the phase AFTER their last reader (a bulk copy in flight at a barrier
      // delayed it): the view tile is free since the barrier that closed P2b, the context tile since the one that closed P5b
      if (t + 1 < a.T) {
        if (ph == 3) fwd_issue_feat(a, pl, smem_raw, S, t + 1);
        if (ph == 7) fwd_issue_ctx(a, pl, smem_raw, S, t + 1);
      }
      switch (ph) {
        case 1: fwd_p2a(a, pl, smem_raw, S, t); break;
        case 2: fwd_p2b(a, pl, smem_raw, S, t); break;
        case 3: fwd_p3<MT>(a, t, red, res); break;          // gates + LSTM cell (CTA = 8 units x 4 gates)
        case 5: fwd_p5a(a, pl, smem_raw, S, t); break;
        case 6: fwd_p5b(a, pl, smem_raw, S, t); break;
        default: fwd_gemm16<MT>(a, t, ph, red, res); break;   // P1 (ph 0), P4 (ph 4), P6 (ph 7): ONE inlined copy
      }
      grid_sync(gb);
      dp_stamp(1 + 8 * t + ph);
    }
  }
}

// fp16 copies of the backward GEMM operands (the gradients du, dt2, dgates, dtk) inside the caller's g16 scratch, multiplied by
// DP_GSCALE = 2^8 so that fp16's normal range covers |g| in [2.4e-7, 256) with TF32's 11 significant bits (smaller values keep an
// absolute step of 2.3e-10, larger ones saturate instead of overflowing); the folded sums are multiplied by 2^-8 (exact).
constexpr float DP_GSCALE = 256.f, DP_GUNSCALE = 1.f / 256.f;
struct BwdX16 { __half* du; __half* dt2; __half* dgates; __half* dtk; };
__device__ __forceinline__ BwdX16 bwd_x16(const dasa_decoder_bwd_t& a) {
  BwdX16 x;
  const int64_t TB = (int64_t)a.T * a.B;
  x.du = reinterpret_cast<__half*>(a.g16);
  x.dt2 = x.du + TB * a.H;
  x.dgates = x.dt2 + TB * a.D;
  x.dtk = x.dgates + TB * 4 * a.H;
  return x;
}

template <int MT>
__device__ __forceinline__ void bwd_gemm16(const dasa_decoder_bwd_t& a, const int t, const int ph, float* red, float* res) {
  const int B = a.B, H = a.H, D = a.D, NK = a.NK;
  const int DC = D + H;
  const int tid = threadIdx.x;
  const int64_t tb = (int64_t)t * B;
  const float scale = a.drop_scale;
  const wt_t* W; const xh_t* X;
  int64_t ldw, ldx;
  int N, K;
  const BwdX16 x16 = bwd_x16(a);
  if (ph == 0)      { W = reinterpret_cast<const wt_t*>(a.w_att_out_t); ldw = a.ld_w_att_out_t; N = DC; K = H;  X = x16.du + tb * H;   ldx = H; }    // B6: dcat = du W_att_out
  else if (ph == 3) { W = reinterpret_cast<const wt_t*>(a.w_att_in_t);  ldw = a.ld_w_att_in_t;  N = H;  K = D;  X = x16.dt2 + tb * D;  ldx = D; }    // B4: dh1d, LSTM cell backward
  else              { W = reinterpret_cast<const wt_t*>(a.w_feat_t);    ldw = a.ld_w_feat_t;    N = H;  K = NK; X = x16.dtk + tb * NK; ldx = NK; }   // B1: dh~_{t-1}, du_{t-1}
  W = dp_opaque(W);
  const bool rows8 = N <= 8 * (int)gridDim.x;                  // B4 / B1 (N = H = 1024): 128 blocks of 8 rows, one per CTA
  const int rb = rows8 ? 8 : 16;
  const int nitems = (N + rb - 1) / rb;
  for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
    const wt_t* wrow[2];
    dp_rows16(wrow, W, ldw, item * rb, N, rows8, 8);
    dp_gemm_block_h<MT, 1, 3>(wrow, X, ldx, B, K, red, res, rows8);
    for (int o = tid; o < 16 * 8 * MT; o += DP_THREADS) {
      const int m = o >> 4, n = item * rb + (o & 15);
      if (m >= B || n >= N || (o & 15) >= rb) continue;
      if (ph == 0) {
        a.dcat[(int64_t)m * DC + n] = res[o] * DP_GUNSCALE;
      } else if (ph == 3) {
        const int u = n;
        const int64_t mi = (tb + m) * H + u;
        const float dh1d = res[o] * DP_GUNSCALE + ld_cg(a.dcat + (int64_t)m * DC + D + u);
        float dhv = a.m_h1 ? (a.m_h1[mi] ? dh1d * scale : 0.f) : dh1d;
        if (a.d_h1) dhv += __ldg(a.d_h1 + mi);
        const float* ac = a.acts + (tb + m) * 4 * H;
        const float ig = __ldg(ac + u), fg = __ldg(ac + H + u), gg = __ldg(ac + 2 * H + u), og = __ldg(ac + 3 * H + u);
        const float cp = __ldg(a.c + mi);
        const float tc = tanhf(__ldg(a.c + (tb + B + m) * H + u));
        const float dcv = ld_cg(a.dc_carry + (int64_t)m * H + u);
        const float dct = dcv + dhv * og * (1.f - tc * tc);
        float* dg = a.dgates + (tb + m) * 4 * H;
        __half* dg16 = x16.dgates + (tb + m) * 4 * H;
        const float d0 = dct * gg * ig * (1.f - ig), d1 = dct * cp * fg * (1.f - fg), d2 = dct * ig * (1.f - gg * gg),
                    d3 = dhv * tc * og * (1.f - og);
        dg[u] = d0; dg[H + u] = d1; dg[2 * H + u] = d2; dg[3 * H + u] = d3;
        dg16[u] = dp_half_sat(d0 * DP_GSCALE); dg16[H + u] = dp_half_sat(d1 * DP_GSCALE);
        dg16[2 * H + u] = dp_half_sat(d2 * DP_GSCALE); dg16[3 * H + u] = dp_half_sat(d3 * DP_GSCALE);
        a.dc_carry[(int64_t)m * H + u] = dct * fg;
      } else {
        const int64_t mi = (tb + m) * H + n;
        float v = res[o] * DP_GUNSCALE;
        v = a.m_hprev ? (a.m_hprev[mi] ? v * scale : 0.f) : v;
        v += ld_cg(a.dhdir + (int64_t)m * H + n);
        if (t > 0) {
          const int64_t pj = mi - (int64_t)B * H;
          const float ht = __ldg(a.htilde + pj);
          const float duv = (__ldg(a.d_htilde + pj) + v) * (1.f - ht * ht);
          a.du[pj] = duv;
          x16.du[pj] = dp_half_sat(duv * DP_GSCALE);
        } else {
          a.dh0[(int64_t)m * H + n] = v;
        }
      }
    }
  }
}

// The 32-row GEMM phases of the backward pass share ONE inlined K loop: B3 (ph 4, d[x ; h] = dgates [W_ih | W_hh], 134 blocks of
// 32 rows) and, when DC outputs need more than one round of 16-row blocks but fit one round of 24-row blocks (DC = 3072 on 148
// CTAs: 128 blocks), B6 (ph 0, dcat = du W_att_out).
__device__ __forceinline__ bool bwd_b6_rows24(const dasa_decoder_bwd_t& a) {
  const int DC = a.D + a.H;
  return DC > 16 * (int)gridDim.x && DC <= 24 * (int)gridDim.x;
}
template <int MT>
__device__ __forceinline__ void bwd_gemm32(const dasa_decoder_bwd_t& a, const int t, const int ph, float* red, float* res) {
  const int B = a.B, H = a.H, E = a.E, F = a.F, D = a.D;
  const int KX = E + F + H, DC = D + H;
  const int tid = threadIdx.x;
  const int64_t tb = (int64_t)t * B;
  const bool b6 = ph == 0;
  const wt_t* W = b6 ? reinterpret_cast<const wt_t*>(a.w_att_out_t) : reinterpret_cast<const wt_t*>(a.w_lstm_t);
  const int64_t ldw = b6 ? a.ld_w_att_out_t : a.ld_w_lstm_t;
  const BwdX16 x16 = bwd_x16(a);
  const xh_t* X = b6 ? x16.du + tb * H : x16.dgates + tb * 4 * H;
  const int K = b6 ? H : 4 * H, N = b6 ? DC : KX;
  const int rb = b6 ? 24 : 32;
  W = dp_opaque(W);
  const int nitems = (N + rb - 1) / rb;
  const int lane = tid & 31, g = lane >> 2, tq = lane & 3;
  for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
    const wt_t* wrow[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int n = item * rb + g + 8 * ((b6 && i == 3) ? 2 : i);
      n = n < N ? n : N - 1;
      wrow[i] = W + (int64_t)n * ldw + 8 * tq;
    }
    dp_gemm_block_h<MT, 2, DP_NST2_H>(wrow, X, K, B, K, red, res, b6);
    for (int o = tid; o < 32 * 8 * MT; o += DP_THREADS) {
      const int m = o >> 5, nl = o & 31, n = item * rb + nl;
      if (m >= B || n >= N || nl >= rb) continue;
      const float v = res[o] * DP_GUNSCALE;
      if (b6) a.dcat[(int64_t)m * DC + n] = v;
      else if (n < E) a.demb[(tb + m) * E + n] = v;
      else if (n < E + F) a.dattn[(int64_t)m * F + (n - E)] = v;
      else a.dhdir[(int64_t)m * H + (n - E - F)] = v;
    }
  }
}

__device__ __noinline__ void bwd_issue_feat(const dasa_decoder_bwd_t& a, const SmemPlan& pl, unsigned char* raw, int S, int t) {
  const AttnOwner o = dp_owner(a.B, S, a.F, nullptr);
  if (!o.on) return;
  const AttnSmem sf = dp_attn_smem(raw, pl, true);
  dp_issue_tile(sf.tile, pl.pitchF, sf.bar, a.feat + (int64_t)t * a.feat_ld_t + (int64_t)o.b * a.feat_ld_b, a.feat_ld_row, a.V,
                o.sl.c0, o.sl.cn, nullptr);
}
__device__ __noinline__ void bwd_issue_ctx(const dasa_decoder_bwd_t& a, const SmemPlan& pl, unsigned char* raw, int S, int t) {
  const AttnOwner o = dp_owner(a.B, S, a.D, dp_mask_smem(raw, pl, a.ctx_mask));
  if (!o.on) return;
  const AttnSmem sc = dp_attn_smem(raw, pl, false);
  dp_issue_tile(sc.tile, pl.pitchC, sc.bar, a.ctx + (int64_t)t * a.ctx_ld_t + (int64_t)o.b * a.ctx_ld_b, a.ctx_ld_row, a.L,
                o.sl.c0, o.sl.cn, o.mask_b);
}

// B5a: partial dalpha_l = ctx_l . dwc   (par = parity of the tile's mbarrier phase)
__device__ __noinline__ void bwd_b5a(const dasa_decoder_bwd_t& a, const SmemPlan& pl, unsigned char* raw, int S, int t, uint32_t par) {
  const AttnOwner o = dp_owner(a.B, S, a.D, dp_mask_smem(raw, pl, a.ctx_mask));
  if (!o.on) return;
  const AttnSmem sc = dp_attn_smem(raw, pl, false);
  const int tid = threadIdx.x, D = a.D, DC = a.D + a.H;
  const int64_t tb = (int64_t)t * a.B;
  for (int c = tid; c < o.sl.cn; c += DP_THREADS) {
    sc.dv[c] = ld_cg(a.dcat + (int64_t)o.b * DC + o.sl.c0 + c);
    sc.tv[c] = __ldg(a.t2 + (tb + o.b) * D + o.sl.c0 + c);
  }
  __syncthreads();
  mbar_wait(sc.bar, par);
  dp_partial_dots(sc.tile, pl.pitchC, sc.dv, a.L, o.sl.cn, o.mask_b, a.zpart + ((size_t)o.b * S + o.s) * DP_MAXROWS);
}

// B5b: dz = alpha (dalpha - sum alpha dalpha); dt2 slice; dctx slice; then prefetch the previous action's tile
__device__ __noinline__ void bwd_b5b(const dasa_decoder_bwd_t& a, const SmemPlan& pl, unsigned char* raw, int S, int t) {
  const AttnOwner o = dp_owner(a.B, S, a.D, dp_mask_smem(raw, pl, a.ctx_mask));
  if (!o.on) return;
  const AttnSmem sc = dp_attn_smem(raw, pl, false);
  const int tid = threadIdx.x, L = a.L, D = a.D;
  const int64_t tb = (int64_t)t * a.B;
  const float* zp = a.zpart + (size_t)o.b * S * DP_MAXROWS;
  for (int r = tid; r < L; r += DP_THREADS) {               // one row per thread: partial sums + saved alpha in one round trip
    const bool ok = o.mask_b == nullptr || o.mask_b[r] == 0;
    sc.zrow[r] = ok ? dp_sum_partials(zp, S, r) : 0.f;
    sc.prow[r] = ok ? __ldg(a.alpha + (tb + o.b) * L + r) : 0.f;
  }
  __syncthreads();
  if (tid < 32) {
    const int lane = tid;
    float pd = 0.f;
    for (int r = lane; r < L; r += 32) pd = fmaf(sc.prow[r], sc.zrow[r], pd);
    pd = warp_sum(pd);
    for (int r = lane; r < L; r += 32) sc.aux[r] = sc.prow[r] * (sc.zrow[r] - pd);
  }
  __syncthreads();
  dp_weighted_sum(sc, pl.pitchC, L, o.sl.cn, o.mask_b, sc.aux, a.dt2 + (tb + o.b) * D + o.sl.c0,      // dt2[c] = sum_l dz_l ctx[l, c]
                  bwd_x16(a).dt2 + (tb + o.b) * D + o.sl.c0, DP_GSCALE);
  // the dctx rows are written by bwd_b5c between the two halves of this phase's barrier
}

// dctx[l, c] = alpha_l dwc[c] + dz_l t2[c] (zero rows where masked) from the vectors B5b left in shared memory; nothing in this
// launch reads dctx, so the 95 KB of stores per CTA go out AFTER the CTA has arrived at the barrier
__device__ __noinline__ void bwd_b5c(const dasa_decoder_bwd_t& a, const SmemPlan& pl, unsigned char* raw, int S, int t) {
  const AttnOwner o = dp_owner(a.B, S, a.D, dp_mask_smem(raw, pl, a.ctx_mask));
  if (!o.on) return;
  const AttnSmem sc = dp_attn_smem(raw, pl, false);
  const int tid = threadIdx.x, L = a.L, D = a.D;
  const int64_t tb = (int64_t)t * a.B;
  const int n4 = o.sl.cn >> 2;
  float* dst_b = a.dctx + ((tb + o.b) * L) * (int64_t)D + o.sl.c0;
  for (int i = tid; i < L * n4; i += DP_THREADS) {
    const int r = i / n4, col = i % n4;
    const float wr = sc.prow[r], dz = sc.aux[r];
    const float4 dw = reinterpret_cast<const float4*>(sc.dv)[col];
    const float4 tv = reinterpret_cast<const float4*>(sc.tv)[col];
    const float4 ov = make_float4(fmaf(wr, dw.x, dz * tv.x), fmaf(wr, dw.y, dz * tv.y), fmaf(wr, dw.z, dz * tv.z),
                                  fmaf(wr, dw.w, dz * tv.w));
    stg_stream4(dst_b + (int64_t)r * D + 4 * col, ov);
  }
}

// B2a: partial dq_v = feat_v . dattn
__device__ __noinline__ void bwd_b2a(const dasa_decoder_bwd_t& a, const SmemPlan& pl, unsigned char* raw, int S, int t, uint32_t par) {
  const AttnOwner o = dp_owner(a.B, S, a.F, nullptr);
  if (!o.on) return;
  const AttnSmem sf = dp_attn_smem(raw, pl, true);
  const int tid = threadIdx.x, F = a.F;
  const int64_t tb = (int64_t)t * a.B;
  for (int c = tid; c < o.sl.cn; c += DP_THREADS) {
    sf.dv[c] = ld_cg(a.dattn + (int64_t)o.b * F + o.sl.c0 + c);
    sf.tv[c] = __ldg(a.tk + (tb + o.b) * a.NK + o.sl.c0 + c);
  }
  __syncthreads();
  mbar_wait(sf.bar, par);
  dp_partial_dots(sf.tile, pl.pitchF, sf.dv, a.V, o.sl.cn, nullptr, a.zpart + ((size_t)o.b * S + o.s) * DP_MAXROWS);
}

// B2b: undo the shift, dz, dt slice, dfeat slice, dkappa logits; then prefetch the previous action's tile
__device__ __noinline__ void bwd_b2b(const dasa_decoder_bwd_t& a, const SmemPlan& pl, unsigned char* raw, int S, int t) {
  const AttnOwner o = dp_owner(a.B, S, a.F, nullptr);
  if (!o.on) return;
  const AttnSmem sf = dp_attn_smem(raw, pl, true);
  const int tid = threadIdx.x, V = a.V, F = a.F, NK = a.NK;
  const int64_t tb = (int64_t)t * a.B;
  const float* zp = a.zpart + (size_t)o.b * S * DP_MAXROWS;
  float* dtk_b = a.dtk + (tb + o.b) * NK;
  __half* dtk16_b = bwd_x16(a).dtk + (tb + o.b) * NK;
  for (int r = tid; r < V; r += DP_THREADS) {               // one row per thread: partial sums + saved p, q in one round trip
    sf.zrow[r] = dp_sum_partials(zp, S, r);
    sf.prow[r] = __ldg(a.p + (tb + o.b) * V + r);
    sf.wrow[r] = __ldg(a.q + (tb + o.b) * V + r);
  }
  if (tid >= DP_THREADS - 32 && (tid & 31) < a.shift_k) sf.kap[tid & 31] = __ldg(a.kappa + (tb + o.b) * a.shift_k + (tid & 31));
  __syncthreads();
  if (tid < 32) {
    const int lane = tid, k = a.shift_k, half = k / 2, Hn = a.headings;
    float* dq = sf.zrow;
    float* dp = sf.aux2;
    for (int r = lane; r < V; r += 32) {
      const int e = r / Hn, m = r % Hn;
      float v = 0.f;
      for (int j = 0; j < k; ++j) {
        int src = (m - j + half) % Hn;
        if (src < 0) src += Hn;
        v = fmaf(sf.kap[j], dq[e * Hn + src], v);
      }
      dp[r] = v;
    }
    // dkappa_j = sum_{e,l} dq[e,l] p[e,(l+j-half) mod Hn];  dkappa_logit = kappa (dkappa - sum kappa dkappa)
    float dk_mine = 0.f, dot = 0.f;
    for (int j = 0; j < k; ++j) {
      float part = 0.f;
      for (int r = lane; r < V; r += 32) {
        const int e = r / Hn, l = r % Hn;
        int src = (l + j - half) % Hn;
        if (src < 0) src += Hn;
        part = fmaf(dq[r], sf.prow[e * Hn + src], part);
      }
      part = warp_sum(part);
      dot = fmaf(sf.kap[j], part, dot);
      if (lane == j) dk_mine = part;
    }
    if (o.s == 0) {
      if (lane < k) {
        const float dkl = sf.kap[lane] * (dk_mine - dot);
        dtk_b[F + lane] = dkl;
        dtk16_b[F + lane] = dp_half_sat(dkl * DP_GSCALE);
      }
      for (int c = F + k + lane; c < NK; c += 32) {       // padding columns of the stacked projection
        dtk_b[c] = 0.f;
        dtk16_b[c] = __float2half_rn(0.f);
      }
    }
    __syncwarp();
    float pd = 0.f;
    for (int r = lane; r < V; r += 32) pd = fmaf(sf.prow[r], dp[r], pd);
    pd = warp_sum(pd);
    for (int r = lane; r < V; r += 32) sf.aux[r] = sf.prow[r] * (dp[r] - pd);
  }
  __syncthreads();
  dp_weighted_sum(sf, pl.pitchF, V, o.sl.cn, nullptr, sf.aux, dtk_b + o.sl.c0, dtk16_b + o.sl.c0, DP_GSCALE);   // dt[c] = sum_v dz_v feat[v, c]
  // dfeat is written by bwd_b2c between the two halves of this phase's barrier
}

// dfeat[v, c] = q_v dattn[c] + dz_v t[c] from the vectors B2b left in shared memory (an output only, like dctx)
__device__ __noinline__ void bwd_b2c(const dasa_decoder_bwd_t& a, const SmemPlan& pl, unsigned char* raw, int S, int t) {
  const AttnOwner o = dp_owner(a.B, S, a.F, nullptr);
  if (!o.on) return;
  const AttnSmem sf = dp_attn_smem(raw, pl, true);
  const int tid = threadIdx.x, V = a.V;
  const int n4 = o.sl.cn >> 2;
  float* dst_b = a.dfeat + (int64_t)t * a.dfeat_ld_t + (int64_t)o.b * a.dfeat_ld_b + o.sl.c0;
  for (int i = tid; i < V * n4; i += DP_THREADS) {
    const int r = i / n4, col = i % n4;
    const float wr = sf.wrow[r], dz = sf.aux[r];
    const float4 dw = reinterpret_cast<const float4*>(sf.dv)[col];
    const float4 tv = reinterpret_cast<const float4*>(sf.tv)[col];
    const float4 ov = make_float4(fmaf(wr, dw.x, dz * tv.x), fmaf(wr, dw.y, dz * tv.y), fmaf(wr, dw.z, dz * tv.z),
                                  fmaf(wr, dw.w, dz * tv.w));
    stg_stream4(dst_b + (int64_t)r * a.dfeat_ld_row + 4 * col, ov);
  }
}

__device__ __noinline__ void bwd_prologue(const dasa_decoder_bwd_t& a) {
  const int B = a.B, H = a.H;
  const int gtid = blockIdx.x * DP_THREADS + threadIdx.x, gthreads = gridDim.x * DP_THREADS;
  for (int i = gtid; i < B * H; i += gthreads) {      // du of the last action, zero cell-state gradient
    const int64_t j = (int64_t)(a.T - 1) * B * H + i;
    const float ht = __ldg(a.htilde + j);
    const float duv = __ldg(a.d_htilde + j) * (1.f - ht * ht);
    a.du[j] = duv;
    bwd_x16(a).du[j] = dp_half_sat(duv * DP_GSCALE);
    a.dc_carry[i] = a.d_c_last ? __ldg(a.d_c_last + i) : 0.f;
  }
}

// ============================================================================================================= BACKWARD
template <int MT>
__global__ void __launch_bounds__(DP_THREADS, 1) decoder_rollout_bwd_kernel(const __grid_constant__ dasa_decoder_bwd_t a, const int S,
                                                                            const __grid_constant__ SmemPlan pl) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* red = reinterpret_cast<float*>(smem_raw + pl.off_red);
  float* res = reinterpret_cast<float*>(smem_raw + pl.off_res);
  GridBar gb{a.barrier, 0u, gridDim.x};
  if (threadIdx.x == 0) {
    const AttnSmem sf = dp_attn_smem(smem_raw, pl, true), sc = dp_attn_smem(smem_raw, pl, false);
    mbar_init(sf.bar, 1); mbar_init(sc.bar, 1); mbar_fence_init();
  }
  dp_stage_mask(smem_raw, pl, a.B, S, a.L, a.ctx_mask, a.ctx_mask_ld);
  __syncthreads();
  bwd_issue_feat(a, pl, smem_raw, S, a.T - 1);    // the tiles of the LAST action stream in under the prologue and B6
  bwd_issue_ctx(a, pl, smem_raw, S, a.T - 1);
  bwd_prologue(a);
  grid_sync(gb);
  dp_stamp(0);
  for (int t = a.T - 1, it = 0; t >= 0; --t, ++it) {
    const uint32_t par = (uint32_t)(it & 1);
#pragma unroll 1
    for (int ph = 0; ph < 8; ++ph) {
      if (t > 0) {                                          // previous action's tiles, after their last reader's barrier
        if (ph == 3) bwd_issue_ctx(a, pl, smem_raw, S, t - 1);
        if (ph == 7) bwd_issue_feat(a, pl, smem_raw, S, t - 1);
      }
      switch (ph) {
        case 1: bwd_b5a(a, pl, smem_raw, S, t, par); break;
        case 2: bwd_b5b(a, pl, smem_raw, S, t); break;
        case 5: bwd_b2a(a, pl, smem_raw, S, t, par); break;
        case 6: bwd_b2b(a, pl, smem_raw, S, t); break;
        default:                                              // GEMM phases: ONE inlined copy of each K loop
          if (ph == 4 || (ph == 0 && bwd_b6_rows24(a))) bwd_gemm32<MT>(a, t, ph, red, res);   // B3, B6 in 24-row blocks
          else bwd_gemm16<MT>(a, t, ph, red, res);                                            // B6 (other shapes), B4 (ph 3), B1 (ph 7)
          break;
      }
      if (ph == 2 || ph == 6) {                             // big output-only rows leave between the halves of the barrier
        grid_arrive(gb);
        if (ph == 2) bwd_b5c(a, pl, smem_raw, S, t);
        else bwd_b2c(a, pl, smem_raw, S, t);
        grid_wait(gb);
      } else {
        grid_sync(gb);
      }
      dp_stamp(1 + 8 * it + ph);
    }
  }
  const int gtid = blockIdx.x * DP_THREADS + threadIdx.x, gthreads = gridDim.x * DP_THREADS;
  for (int i = gtid; i < a.B * a.H; i += gthreads) a.dc0[i] = ld_cg(a.dc_carry + i);
}

// ---- barrier micro-benchmark (profiling aid): `iters` device-wide barriers with no work in between; out[0] = SM clocks of CTA 0
template <int VARIANT>
__device__ __forceinline__ void grid_sync_variant(GridBar& gb) {
  gb.target += gb.nblk;
  __syncthreads();
  if (threadIdx.x == 0) {
    if (VARIANT == 0) {                     // fences on both sides of a release-add / acquire-poll
      __threadfence();
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(gb.ctr) : "memory");
      unsigned int v, it = 0;
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(gb.ctr) : "memory");
        if (++it > (1u << 22)) __trap();
      } while ((int)(v - gb.target) < 0);
      __threadfence();
    } else if (VARIANT == 1) {                     // release-add / acquire-poll only (no extra fences)
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(gb.ctr) : "memory");
      unsigned int v, it = 0;
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(gb.ctr) : "memory");
        if (++it > (1u << 22)) __trap();
      } while ((int)(v - gb.target) < 0);
    } else if (VARIANT == 3) {              // release-add, RELAXED polls (no L1 invalidate per iteration), one acquire fence at the end
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(gb.ctr) : "memory");
      unsigned int v, it = 0;
      do {
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(gb.ctr) : "memory");
        if (++it > (1u << 22)) __trap();
      } while ((int)(v - gb.target) < 0);
      asm volatile("fence.acq_rel.gpu;" ::: "memory");
    } else if (VARIANT == 4) {              // the same with two polls in flight (detection delay = half a round trip)
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(gb.ctr) : "memory");
      unsigned int v0, v1, it = 0;
      asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v0) : "l"(gb.ctr) : "memory");
      for (;;) {
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v1) : "l"(gb.ctr) : "memory");
        if ((int)(v0 - gb.target) >= 0) break;
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v0) : "l"(gb.ctr) : "memory");
        if ((int)(v1 - gb.target) >= 0) break;
        if (++it > (1u << 22)) __trap();
      }
      asm volatile("fence.acq_rel.gpu;" ::: "memory");
    } else {                                // VARIANT 2: fence + relaxed atomic + volatile poll + fence (cooperative-groups style)
      __threadfence();
      atomicAdd(gb.ctr, 1u);
      unsigned int it = 0;
      while ((int)(*((volatile unsigned int*)gb.ctr) - gb.target) < 0) { if (++it > (1u << 22)) __trap(); }
      __threadfence();
    }
  }
  __syncthreads();
}

template <int VARIANT>
__global__ void __launch_bounds__(DP_THREADS, 1) barrier_bench_kernel(unsigned int* ctr, int iters, long long* out, float* junk) {
  GridBar gb{ctr, 0u, gridDim.x};
  grid_sync_variant<VARIANT>(gb);
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    junk[blockIdx.x * DP_THREADS + threadIdx.x] = (float)i;          // one global write per thread per interval, like a phase
    grid_sync_variant<VARIANT>(gb);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = clock64() - t0;
}

// ------------------------------------------------------------------------------------------------------------- host side
struct LaunchPlan {
  int grid, S, MT;
  SmemPlan pl;
  bool ok;
};

int dp_sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = DASA_NUM_SMS;
  }
  return n;
}

LaunchPlan dp_launch_plan(int B, int H, int E, int F, int V, int L, int D, int NK, int shift_k) {
  LaunchPlan lp{};
  lp.ok = false;
  if (B < 1 || B > 32 || H < 16 || (H % 16) != 0 || ((E + F + H) % 32) != 0 || (NK % 32) != 0 || (D % 32) != 0 || (F % 4) != 0 ||
      (E % 4) != 0 || (H % 32) != 0 || V < 1 || V > DP_MAXROWS || L < 1 || L > DP_MAXROWS || shift_k < 1 || shift_k > DP_MAXK ||
      NK < F + shift_k)
    return lp;
  lp.grid = dp_sm_count();
  lp.MT = (B + 7) / 8;
  int S = lp.grid / B;
  if (S > 8) S = 8;
  if (S < 1) return lp;
  while (S > 1 && ((F >> 2) < S || (D >> 2) < S)) --S;
  lp.S = S;
  lp.pl = dp_plan(lp.MT, V, L, F, D, S);
  lp.ok = lp.pl.total <= 225 * 1024;
  return lp;
}

template <typename Kern, typename Args>
int dp_launch(Kern kern, const Args& a, const LaunchPlan& lp, unsigned int* barrier, cudaStream_t st, const char* name) {
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lp.pl.total);
  if (e != cudaSuccess) { dasa_set_error(name, e); return DASA_ERR_CUDA; }
  e = cudaMemsetAsync(barrier, 0, sizeof(unsigned int), st);
  if (e != cudaSuccess) { dasa_set_error(name, e); return DASA_ERR_CUDA; }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)lp.grid);
  cfg.blockDim = dim3(DP_THREADS);
  cfg.dynamicSmemBytes = lp.pl.total;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;      // all CTAs co-resident, or the launch fails (never a partial grid that deadlocks)
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, kern, a, lp.S, lp.pl);
  if (e != cudaSuccess) { dasa_set_error(name, e); return DASA_ERR_CUDA; }
  return DASA_OK;
}

}  // namespace

extern "C" int dasa_decoder_rollout_supported(int B, int H, int E, int F, int V, int L, int D, int NK, int shift_k) {
  return dp_launch_plan(B, H, E, F, V, L, D, NK, shift_k).ok ? 1 : 0;
}

extern "C" int dasa_debug_barrier_bench(int variant, int iters, unsigned int* ctr, long long* out, float* junk, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(ctr, 0, sizeof(unsigned int), st) != cudaSuccess) return DASA_ERR_CUDA;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)dp_sm_count());
  cfg.blockDim = dim3(DP_THREADS);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e;
  if (variant == 0) e = cudaLaunchKernelEx(&cfg, barrier_bench_kernel<0>, ctr, iters, out, junk);
  else if (variant == 1) e = cudaLaunchKernelEx(&cfg, barrier_bench_kernel<1>, ctr, iters, out, junk);
  else if (variant == 3) e = cudaLaunchKernelEx(&cfg, barrier_bench_kernel<3>, ctr, iters, out, junk);
  else if (variant == 4) e = cudaLaunchKernelEx(&cfg, barrier_bench_kernel<4>, ctr, iters, out, junk);
  else e = cudaLaunchKernelEx(&cfg, barrier_bench_kernel<2>, ctr, iters, out, junk);
  if (e != cudaSuccess) { dasa_set_error("barrier_bench_kernel", e); return DASA_ERR_CUDA; }
  return DASA_OK;
}

extern "C" int dasa_debug_decoder_phase_clocks(long long* out, int n) {
  const int m = 1 + 8 * DP_TIMED_ACTIONS;
  long long h[1 + 8 * DP_TIMED_ACTIONS];
  if (cudaMemcpyFromSymbol(h, g_dp_clock, sizeof(h)) != cudaSuccess) return DASA_ERR_CUDA;
  for (int i = 0; i < n && i < m; ++i) out[i] = h[i];
  return m;
}

extern "C" size_t dasa_decoder_rollout_x16_halves(int T, int B, int H, int E, int F, int D) {
  if (T < 1 || B < 1) return 0;
  return (size_t)T * B * ((size_t)H + (size_t)(E + F + H) + (size_t)(D + H));
}

extern "C" size_t dasa_decoder_rollout_g16_halves(int T, int B, int H, int D, int NK) {
  if (T < 1 || B < 1) return 0;
  return (size_t)T * B * ((size_t)H + (size_t)D + (size_t)4 * H + (size_t)NK);
}

extern "C" size_t dasa_decoder_rollout_scratch_floats(int B) { return (size_t)(B < 1 ? 1 : B) * 8 * DP_MAXROWS; }

extern "C" int dasa_decoder_rollout_fwd(const dasa_decoder_fwd_t* a, void* stream) {
  if (a == nullptr || a->T <= 0) return a == nullptr ? DASA_ERR_BAD_SHAPE : DASA_OK;
  const LaunchPlan lp = dp_launch_plan(a->B, a->H, a->E, a->F, a->V, a->L, a->D, a->NK, a->shift_k);
  if (!lp.ok) return DASA_ERR_UNSUPPORTED;
  if (a->headings <= 0 || (a->V % a->headings) != 0) return DASA_ERR_BAD_SHAPE;
  if (!dasa_aligned16(a->feat) || !dasa_aligned16(a->ctx) || (a->feat_ld_row & 3) || (a->feat_ld_b & 3) || (a->feat_ld_t & 3) ||
      (a->ctx_ld_row & 3) || (a->ctx_ld_b & 3) || (a->ctx_ld_t & 3))
    return DASA_ERR_BAD_ALIGN;
  if ((a->H & 7) || ((a->E + a->F + a->H) & 7) || ((a->D + a->H) & 7) || (a->D & 7) || ((a->E + a->F) & 7)) return DASA_ERR_BAD_ALIGN;
  const void* w[] = {a->w_feat, a->w_lstm, a->w_att_in, a->w_att_out, a->hprev_drop, a->xh, a->cat, a->tk, a->t2, a->x16};
  for (const void* p : w)
    if (p == nullptr || !dasa_aligned16(p)) return DASA_ERR_BAD_ALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  switch (lp.MT) {
    case 1: return dp_launch(decoder_rollout_fwd_kernel<1>, *a, lp, a->barrier, st, "decoder_rollout_fwd_kernel<1>");
    case 2: return dp_launch(decoder_rollout_fwd_kernel<2>, *a, lp, a->barrier, st, "decoder_rollout_fwd_kernel<2>");
    case 3: return dp_launch(decoder_rollout_fwd_kernel<3>, *a, lp, a->barrier, st, "decoder_rollout_fwd_kernel<3>");
    default: return dp_launch(decoder_rollout_fwd_kernel<4>, *a, lp, a->barrier, st, "decoder_rollout_fwd_kernel<4>");
  }
}

extern "C" int dasa_decoder_rollout_bwd(const dasa_decoder_bwd_t* a, void* stream) {
  if (a == nullptr || a->T <= 0) return a == nullptr ? DASA_ERR_BAD_SHAPE : DASA_OK;
  const LaunchPlan lp = dp_launch_plan(a->B, a->H, a->E, a->F, a->V, a->L, a->D, a->NK, a->shift_k);
  if (!lp.ok) return DASA_ERR_UNSUPPORTED;
  if (a->headings <= 0 || (a->V % a->headings) != 0) return DASA_ERR_BAD_SHAPE;
  if (!dasa_aligned16(a->feat) || !dasa_aligned16(a->ctx) || (a->feat_ld_row & 3) || (a->feat_ld_b & 3) || (a->feat_ld_t & 3) ||
      (a->ctx_ld_row & 3) || (a->ctx_ld_b & 3) || (a->ctx_ld_t & 3) || (a->dfeat_ld_row & 3) || (a->dfeat_ld_b & 3) ||
      (a->dfeat_ld_t & 3) || (a->ld_w_feat_t & 7) || (a->ld_w_lstm_t & 7) || (a->ld_w_att_in_t & 7) || (a->ld_w_att_out_t & 7) ||
      (a->H & 7) || (a->D & 7) || (a->NK & 7))
    return DASA_ERR_BAD_ALIGN;
  const void* w[] = {a->w_feat_t, a->w_lstm_t, a->w_att_in_t, a->w_att_out_t, a->du, a->dt2, a->dgates, a->dtk, a->dfeat, a->dctx,
                     a->dcat, a->dattn, a->g16};
  for (const void* p : w)
    if (p == nullptr || !dasa_aligned16(p)) return DASA_ERR_BAD_ALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  switch (lp.MT) {
    case 1: return dp_launch(decoder_rollout_bwd_kernel<1>, *a, lp, a->barrier, st, "decoder_rollout_bwd_kernel<1>");
    case 2: return dp_launch(decoder_rollout_bwd_kernel<2>, *a, lp, a->barrier, st, "decoder_rollout_bwd_kernel<2>");
    case 3: return dp_launch(decoder_rollout_bwd_kernel<3>, *a, lp, a->barrier, st, "decoder_rollout_bwd_kernel<3>");
    default: return dp_launch(decoder_rollout_bwd_kernel<4>, *a, lp, a->barrier, st, "decoder_rollout_bwd_kernel<4>");
  }
}
