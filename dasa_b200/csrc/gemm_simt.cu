// FP32 (FFMA) GEMM with fused epilogues — the exact-fp32 precision mode of dasa_gemm, and the path every
// projection takes when its operands do not satisfy the TMA alignment rules of the tcgen05 kernel (gemm_tc.cu).
// Register-tiled, double-buffered through registers, split-K with a deterministic two-pass reduction.
#include <cuda_fp16.h>
#include "common.cuh"
#define DASA_GELU_EXACT 1
#include "gemm_common.cuh"

namespace {

template <int BM, int BN, int BK, int TM, int TN, bool AK, bool BKM>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
sgemm_kernel(int M, int N, int K, float alpha, const float* __restrict__ A, int64_t lda, const float* __restrict__ B,
             int64_t ldb, float beta, float* __restrict__ C, int64_t ldc, int epilogue, EpiParams ep,
             float* __restrict__ partial, int k_per_split) {
  constexpr int NT = (BM / TM) * (BN / TN);
  constexpr int PAD = 4;
  __shared__ __align__(16) float As[2][BK][BM + PAD];
  __shared__ __align__(16) float Bs[2][BK][BN + PAD];

  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kbeg = blockIdx.z * k_per_split;
  const int kend = min(K, kbeg + k_per_split);
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);

  // element-wise tile loaders (coalesced along the contiguous axis of each operand)
  constexpr int A_ELEMS = BM * BK / NT, B_ELEMS = BN * BK / NT;
  static_assert(BM * BK % NT == 0 && BN * BK % NT == 0, "tile/thread mismatch");
  float ra[A_ELEMS], rb[B_ELEMS];

  auto load_a = [&](int k0) {
#pragma unroll
    for (int i = 0; i < A_ELEMS; ++i) {
      const int e = tid + i * NT;
      int m, k;
      if (AK) { k = e % BK; m = e / BK; } else { m = e % BM; k = e / BM; }
      const int gm = m0 + m, gk = k0 + k;
      float v = 0.f;
      if (gm < M && gk < kend) v = AK ? A[(int64_t)gm * lda + gk] : A[(int64_t)gk * lda + gm];
      ra[i] = v;
    }
  };
  auto load_b = [&](int k0) {
#pragma unroll
    for (int i = 0; i < B_ELEMS; ++i) {
      const int e = tid + i * NT;
      int n, k;
      if (BKM) { k = e % BK; n = e / BK; } else { n = e % BN; k = e / BN; }
      const int gn = n0 + n, gk = k0 + k;
      float v = 0.f;
      if (gn < N && gk < kend) v = BKM ? B[(int64_t)gn * ldb + gk] : B[(int64_t)gk * ldb + gn];
      rb[i] = v;
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int i = 0; i < A_ELEMS; ++i) {
      const int e = tid + i * NT;
      int m, k;
      if (AK) { k = e % BK; m = e / BK; } else { m = e % BM; k = e / BM; }
      As[buf][k][m] = ra[i];
    }
#pragma unroll
    for (int i = 0; i < B_ELEMS; ++i) {
      const int e = tid + i * NT;
      int n, k;
      if (BKM) { k = e % BK; n = e / BK; } else { n = e % BN; k = e / BN; }
      Bs[buf][k][n] = rb[i];
    }
  };

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  int buf = 0;
  if (kbeg < kend) {
    load_a(kbeg);
    load_b(kbeg);
    store_tiles(0);
  }
  __syncthreads();
  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    const bool more = (k0 + BK) < kend;
    if (more) { load_a(k0 + BK); load_b(k0 + BK); }
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; i += 4) {
        const float4 v = *reinterpret_cast<const float4*>(&As[buf][k][ty * TM + i]);
        a[i] = v.x; a[i + 1] = v.y; a[i + 2] = v.z; a[i + 3] = v.w;
      }
#pragma unroll
      for (int j = 0; j < TN; j += 4) {
        const float4 v = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * TN + j]);
        b[j] = v.x; b[j + 1] = v.y; b[j + 2] = v.z; b[j + 3] = v.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (more) {
      store_tiles(buf ^ 1);
      __syncthreads();
      buf ^= 1;
    }
  }

  // epilogue
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + ty * TM + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * TN + j;
      if (n >= N) continue;
      if (partial != nullptr) {
        partial[((int64_t)blockIdx.z * M + m) * N + n] = acc[i][j];
      } else {
        // direct store path: only the plain / bias epilogues (others are routed through the reduction kernel)
        float v = alpha * acc[i][j];
        if (beta != 0.f) v += beta * C[(int64_t)m * ldc + n];
        if (epilogue == DASA_EPI_BIAS && ep.bias != nullptr) v += __ldg(ep.bias + n);
        C[(int64_t)m * ldc + n] = v;
      }
    }
  }
}

__global__ void splitk_reduce_kernel(const float* __restrict__ partial, int S, int M, int N, float alpha, float beta,
                                     float* __restrict__ C, int64_t ldc, int epilogue, EpiParams ep) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)M * N) return;
  const int m = idx / N, n = idx % N;
  float s = 0.f;
  for (int z = 0; z < S; ++z) s += partial[(int64_t)z * M * N + idx];
  float v = alpha * s;
  if (beta != 0.f) v += beta * C[(int64_t)m * ldc + n];
  C[(int64_t)m * ldc + n] = apply_epilogue(v, m, n, N, epilogue, ep);
}

// out[n] (+)= sum_m X[m, n]: grid (N/32 column groups, row slabs); 32x8 threads. The slabs of a column group are combined in slab
// order by the LAST block of that group to finish (ticket counter): bit-reproducible bias gradients, no float atomics.
constexpr int COLSUM_PART_FLOATS = 1 << 17;
constexpr int COLSUM_MAX_GROUPS = 1 << 12;
__device__ float g_colsum_part[COLSUM_PART_FLOATS];
__device__ unsigned int g_colsum_ticket[COLSUM_MAX_GROUPS];

__device__ __forceinline__ float colsum_ld(const float* p) { return *p; }
__device__ __forceinline__ float colsum_ld(const __half* p) { return __half2float(*p); }

// T = float, or __half with `scale` applied to the sums (the scaled fp16 gradient copies of the fp16-operand weight-gradient path:
// half the bytes of this read-once, HBM-bound pass)
template <typename T>
__global__ void colsum_kernel(const T* __restrict__ X, int64_t ldx, int M, int N, float* __restrict__ out, int rows_per_slab, float scale) {
  __shared__ float red[8][33];
  __shared__ bool last;
  const int n = blockIdx.x * 32 + threadIdx.x;
  const int m0 = blockIdx.y * rows_per_slab, m1 = min(M, m0 + rows_per_slab);
  float s = 0.f;
  if (n < N)
    for (int m = m0 + threadIdx.y; m < m1; m += 8) s += colsum_ld(X + (int64_t)m * ldx + n);
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  float t = 0.f;
  if (threadIdx.y == 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x];
  }
  if (gridDim.y == 1) {
    if (threadIdx.y == 0 && n < N) out[n] += t * scale;
    return;
  }
  const int ngrp = gridDim.x * 32;
  if (threadIdx.y == 0) g_colsum_part[(size_t)blockIdx.y * ngrp + blockIdx.x * 32 + threadIdx.x] = t;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0 && threadIdx.y == 0) last = atomicAdd(&g_colsum_ticket[blockIdx.x], 1u) == gridDim.y - 1;
  __syncthreads();
  if (!last) return;
  __threadfence();
  if (threadIdx.y == 0) {
    float v = 0.f;
    for (int z = 0; z < (int)gridDim.y; ++z) v += __ldcg(g_colsum_part + (size_t)z * ngrp + blockIdx.x * 32 + threadIdx.x);
    if (n < N) out[n] += v * scale;
    if (threadIdx.x == 0) g_colsum_ticket[blockIdx.x] = 0;
  }
}

__global__ void zero_kernel(float* __restrict__ p, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = 0.f;
}

struct SimtPlan { bool skinny; int splits; int k_per_split; };

SimtPlan plan_simt(int M, int N, int K) {
  SimtPlan p;
  p.skinny = (M <= 48);
  const int bm = p.skinny ? 32 : 128, bn = p.skinny ? 64 : 128, bk = p.skinny ? 16 : 8;
  const int64_t tiles = dasa_cdiv(M, bm) * dasa_cdiv(N, bn);
  int s = 1;
  if (tiles < 2 * DASA_NUM_SMS && K >= 8 * bk) {
    s = (int)dasa_cdiv(2 * DASA_NUM_SMS, tiles);
    s = (int)min((int64_t)s, (int64_t)K / (4 * bk));
    s = max(1, min(s, 64));
  }
  int kps = (int)dasa_cdiv(K, s);
  kps = (int)dasa_cdiv(kps, bk) * bk;
  p.splits = (int)dasa_cdiv(K, kps);
  p.k_per_split = kps;
  return p;
}

template <int BM, int BN, int BK, int TM, int TN>
int launch_simt(int a_k, int b_k, int M, int N, int K, float alpha, const float* A, int64_t lda, const float* B,
                int64_t ldb, float beta, float* C, int64_t ldc, int epilogue, const EpiParams& ep, float* partial,
                const SimtPlan& p, cudaStream_t st) {
  dim3 grid((unsigned)dasa_cdiv(N, BN), (unsigned)dasa_cdiv(M, BM), (unsigned)p.splits);
  dim3 block((BM / TM) * (BN / TN));
#define DASA_LAUNCH(AKV, BKV)                                                                                      \
  sgemm_kernel<BM, BN, BK, TM, TN, AKV, BKV><<<grid, block, 0, st>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, \
                                                                       epilogue, ep, partial, p.k_per_split)
  if (a_k && b_k) DASA_LAUNCH(true, true);
  else if (a_k && !b_k) DASA_LAUNCH(true, false);
  else if (!a_k && b_k) DASA_LAUNCH(false, true);
  else DASA_LAUNCH(false, false);
#undef DASA_LAUNCH
  return dasa_check_launch("sgemm_kernel");
}

}  // namespace

size_t dasa_gemm_simt_workspace(int M, int N, int K) {
  SimtPlan p = plan_simt(M, N, K);
  return (size_t)p.splits * M * N * sizeof(float);          // also the staging buffer of the two-pass activation epilogues
}

int dasa_gemm_simt(int a_kmajor, int b_kmajor, int M, int N, int K, float alpha, const float* A, int64_t lda,
                   const float* B, int64_t ldb, float beta, float* C, int64_t ldc, int epilogue, const EpiParams& ep,
                   void* workspace, size_t workspace_bytes, cudaStream_t st) {
  if (M <= 0 || N <= 0) return DASA_OK;
  if (K <= 0) return DASA_ERR_BAD_SHAPE;
  SimtPlan p = plan_simt(M, N, K);
  float* partial = nullptr;
  const bool simple = (epilogue == DASA_EPI_NONE || epilogue == DASA_EPI_BIAS) && ep.drop_mask == nullptr;
  if (p.splits > 1 || !simple) {
    const size_t need = (size_t)p.splits * M * N * sizeof(float);
    if (workspace == nullptr || workspace_bytes < need) {
      if (!simple) return DASA_ERR_WORKSPACE;                 // fused activations need the staging buffer
      p.splits = 1;                                           // run unsplit rather than fail
      p.k_per_split = K;
    } else {
      partial = static_cast<float*>(workspace);
    }
  }
  int rc;
  if (p.skinny)
    rc = launch_simt<32, 64, 16, 4, 4>(a_kmajor, b_kmajor, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, epilogue, ep,
                                       partial, p, st);
  else
    rc = launch_simt<128, 128, 8, 8, 8>(a_kmajor, b_kmajor, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, epilogue, ep,
                                        partial, p, st);
  if (rc != DASA_OK) return rc;
  if (partial != nullptr) {
    const int64_t total = (int64_t)M * N;
    splitk_reduce_kernel<<<(unsigned)dasa_cdiv(total, 256), 256, 0, st>>>(partial, p.splits, M, N, alpha, beta, C, ldc,
                                                                           epilogue, ep);
    rc = dasa_check_launch("splitk_reduce_kernel");
  }
  return rc;
}

template <typename T>
static int colsum_launch(const T* X, int64_t ldx, int M, int N, float scale, float* out, int accumulate, cudaStream_t st) {
  if (N <= 0) return DASA_OK;
  if (!accumulate) {
    zero_kernel<<<(unsigned)dasa_cdiv(N, 256), 256, 0, st>>>(out, N);
    if (dasa_check_launch("zero_kernel") != DASA_OK) return DASA_ERR_CUDA;
  }
  if (M <= 0) return DASA_OK;
  const int col_groups = (int)dasa_cdiv(N, 32);
  int slabs = (int)dasa_cdiv(4 * DASA_NUM_SMS, col_groups);          // ~4 CTAs per SM in total
  const int max_slabs = (int)dasa_cdiv(M, 64);
  slabs = slabs < 1 ? 1 : (slabs > max_slabs ? max_slabs : slabs);
  // the slab partials live in a static scratch (launches on one stream are ordered; one training process per device)
  if (col_groups > COLSUM_MAX_GROUPS) slabs = 1;
  while (slabs > 1 && (int64_t)slabs * col_groups * 32 > COLSUM_PART_FLOATS) --slabs;
  const int rows_per_slab = (int)dasa_cdiv(M, slabs);
  dim3 grid((unsigned)col_groups, (unsigned)dasa_cdiv(M, rows_per_slab));
  colsum_kernel<T><<<grid, dim3(32, 8), 0, st>>>(X, ldx, M, N, out, rows_per_slab, scale);
  return dasa_check_launch("colsum_kernel");
}

extern "C" int dasa_colsum(const float* X, int64_t ldx, int M, int N, float* out, int accumulate, void* stream) {
  return colsum_launch<float>(X, ldx, M, N, 1.f, out, accumulate, (cudaStream_t)stream);
}

// fp16 input, 8 columns (one 128-bit load) per thread: a warp row covers 256 columns. Same slab / ticket scheme as colsum_kernel
// (bit-reproducible). N % 8 == 0, ldx % 8 == 0.
__global__ void __launch_bounds__(256) colsum_h8_kernel(const __half* __restrict__ X, int64_t ldx, int M, int N, float* __restrict__ out,
                                                        int rows_per_slab, float scale) {
  __shared__ float red[8][32 * 8 + 8];
  __shared__ bool last;
  const int n0 = (blockIdx.x * 32 + threadIdx.x) * 8;
  const int m0 = blockIdx.y * rows_per_slab, m1 = min(M, m0 + rows_per_slab);
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  if (n0 < N)
    for (int m = m0 + threadIdx.y; m < m1; m += 8) {
      uint4 v;
      asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                   : "l"(X + (int64_t)m * ldx + n0));
      const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __half22float2(h[e]);
        acc[2 * e] += f.x; acc[2 * e + 1] += f.y;
      }
    }
#pragma unroll
  for (int e = 0; e < 8; ++e) red[threadIdx.y][threadIdx.x * 8 + e] = acc[e];
  __syncthreads();
  const int tid = threadIdx.y * 32 + threadIdx.x;            // 256 threads = the 256 columns of this group
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) t += red[i][tid];
  const int n = blockIdx.x * 256 + tid;
  if (gridDim.y == 1) {
    if (n < N) out[n] += t * scale;
    return;
  }
  const int ngrp = gridDim.x * 256;
  g_colsum_part[(size_t)blockIdx.y * ngrp + blockIdx.x * 256 + tid] = t;
  __threadfence();
  __syncthreads();
  if (tid == 0) last = atomicAdd(&g_colsum_ticket[blockIdx.x], 1u) == gridDim.y - 1;
  __syncthreads();
  if (!last) return;
  __threadfence();
  float v = 0.f;
  for (int z = 0; z < (int)gridDim.y; ++z) v += __ldcg(g_colsum_part + (size_t)z * ngrp + blockIdx.x * 256 + tid);
  if (n < N) out[n] += v * scale;
  if (tid == 0) g_colsum_ticket[blockIdx.x] = 0;
}

extern "C" int dasa_colsum_h(const dasa_half_t* X, int64_t ldx, int M, int N, float scale, float* out, int accumulate, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const __half* Xh = reinterpret_cast<const __half*>(X);
  if ((N & 7) || (ldx & 7) || (reinterpret_cast<uintptr_t>(X) & 15)) return colsum_launch<__half>(Xh, ldx, M, N, scale, out, accumulate, st);
  if (N <= 0) return DASA_OK;
  if (!accumulate) {
    zero_kernel<<<(unsigned)dasa_cdiv(N, 256), 256, 0, st>>>(out, N);
    if (dasa_check_launch("zero_kernel") != DASA_OK) return DASA_ERR_CUDA;
  }
  if (M <= 0) return DASA_OK;
  const int col_groups = (int)dasa_cdiv(N, 256);
  int slabs = (int)dasa_cdiv(4 * DASA_NUM_SMS, col_groups);
  const int max_slabs = (int)dasa_cdiv(M, 64);
  slabs = slabs < 1 ? 1 : (slabs > max_slabs ? max_slabs : slabs);
  if (col_groups > COLSUM_MAX_GROUPS) slabs = 1;
  while (slabs > 1 && (int64_t)slabs * col_groups * 256 > COLSUM_PART_FLOATS) --slabs;
  const int rows_per_slab = (int)dasa_cdiv(M, slabs);
  dim3 grid((unsigned)col_groups, (unsigned)dasa_cdiv(M, rows_per_slab));
  colsum_h8_kernel<<<grid, dim3(32, 8), 0, st>>>(Xh, ldx, M, N, out, rows_per_slab, scale);
  return dasa_check_launch("colsum_h8_kernel");
}
