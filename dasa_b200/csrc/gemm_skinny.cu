// Skinny TF32 GEMM for the decoder's per-action projections: C[M,N] = epi(alpha * A[M,K] * B[N,K]^T + beta*C) with M <= 32
// rows (M = the 20 episodes of a rollout). These are weight-streaming problems (N*K*4 bytes of W for 2*M*N*K FLOP): the
// tcgen05 tile kernels pad M to 128 rows, need split-K plus a second reduction launch to fill the machine, and spend most of
// their ~13 us in prologue / epilogue latency. Here every warp streams a [32 rows of W] x [K slice] panel straight from
// global memory with 128-bit loads (no shared-memory staging: each weight element is used exactly once), feeds it to
// mma.sync.m16n8k8 TF32 as the A operand (W rows = MMA M, the batch rows = MMA N), and the K slices are reduced in a fixed
// order: across the 8 warps of a CTA through shared memory, across the CTAs of a thread-block cluster (up to 8 along K)
// through distributed shared memory. One launch, deterministic, epilogue fused.
//
// k-index trick: a lane loads float4 = 4 CONSECUTIVE k of its W row and of its A row and uses component j in MMA step j, so
// "MMA k index t / t+4 of step j" means physical k = 4t+j / 16+4t+j of the 32-wide chunk for both operands: any bijection of
// the reduction index is a valid GEMM, and every global access stays 128-bit.
#include <cuda.h>
#include <stdlib.h>
#include "common.cuh"
#include "gemm_common.cuh"

namespace {

constexpr int SK_WARPS = 8;
constexpr int SK_THREADS = SK_WARPS * 32;
constexpr int SK_ROWS = 32;          // W rows (output columns) per CTA: two m16 MMA tiles per warp
constexpr int SK_CHUNK = 32;         // k per chunk

struct SkinnyParams {
  const float* A; int64_t lda;       // [M, K]
  const float* W; int64_t ldb;       // [N, K]
  float* C; int64_t ldc;
  int M, N, K;
  float alpha, beta;
  int epilogue; EpiParams ep;
  int cluster;                       // CTAs along K (cluster size)
};

__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x)); return r;
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t sk_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void sk_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void sk_st_remote(float* local_ptr, uint32_t rank, float v) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(local_ptr)), "r"(rank));
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(ra), "f"(v) : "memory");
}

// MT = number of 8-row batch tiles (M <= 8*MT)
template <int MT>
__global__ void __launch_bounds__(SK_THREADS, 2) gemm_skinny_tf32_kernel(SkinnyParams p) {
  extern __shared__ __align__(16) float sk_smem[];
  constexpr int OUTS = SK_ROWS * 8 * MT;                       // outputs of one CTA
  float* red = sk_smem;                                        // [SK_WARPS][2][MT][4][32] per-warp accumulator fragments
  float* part = sk_smem + SK_WARPS * 2 * MT * 4 * 32;          // [cluster-1][OUTS] partial sums pushed by the other CTAs (rank 0 only)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const uint32_t crank = p.cluster > 1 ? sk_ctarank() : 0u;
  if (p.cluster > 1) sk_cluster_sync();                        // every CTA of the cluster is resident before DSMEM is touched
  const int n0 = blockIdx.x * SK_ROWS;
  const int nchunks = p.K / SK_CHUNK;
  const int units = SK_WARPS * p.cluster;
  const int unit = (int)crank * SK_WARPS + warp;

  // the four W rows this lane reads (rows past N are clamped; their results are never stored)
  const float* wrow[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int n = n0 + g + 8 * i;
    n = n < p.N ? n : p.N - 1;
    wrow[i] = p.W + (int64_t)n * p.ldb + 4 * t;
  }
  const float* arow[MT];
  bool aok[MT];
#pragma unroll
  for (int i = 0; i < MT; ++i) {
    const int m = 8 * i + g;
    aok[i] = m < p.M;
    arow[i] = p.A + (int64_t)(aok[i] ? m : 0) * p.lda + 4 * t;
  }

  float acc[2][MT][4];
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[h][i][e] = 0.f;

#pragma unroll 1
  for (int c = unit; c < nchunks; c += units) {
    const int kc = c * SK_CHUNK;
    float4 w[4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      w[i][0] = ldg_stream4(wrow[i] + kc);
      w[i][1] = ldg_stream4(wrow[i] + kc + 16);
    }
    float4 x[MT][2];
#pragma unroll
    for (int i = 0; i < MT; ++i) {
      if (aok[i]) {
        x[i][0] = __ldg(reinterpret_cast<const float4*>(arow[i] + kc));
        x[i][1] = __ldg(reinterpret_cast<const float4*>(arow[i] + kc + 16));
      } else {
        x[i][0] = make_float4(0.f, 0.f, 0.f, 0.f);
        x[i][1] = x[i][0];
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint32_t bf[MT][2];
#pragma unroll
      for (int i = 0; i < MT; ++i) {
        bf[i][0] = to_tf32(reinterpret_cast<const float*>(&x[i][0])[j]);
        bf[i][1] = to_tf32(reinterpret_cast<const float*>(&x[i][1])[j]);
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t af[4];
        af[0] = to_tf32(reinterpret_cast<const float*>(&w[2 * h][0])[j]);       // (row g,     k = t)
        af[1] = to_tf32(reinterpret_cast<const float*>(&w[2 * h + 1][0])[j]);   // (row g + 8, k = t)
        af[2] = to_tf32(reinterpret_cast<const float*>(&w[2 * h][1])[j]);       // (row g,     k = t + 4)
        af[3] = to_tf32(reinterpret_cast<const float*>(&w[2 * h + 1][1])[j]);   // (row g + 8, k = t + 4)
#pragma unroll
        for (int i = 0; i < MT; ++i) mma_tf32(acc[h][i], af, bf[i][0], bf[i][1]);
      }
    }
  }

  // ---- reduce the K slices: warps of this CTA (shared memory), then CTAs of the cluster (DSMEM), always in index order
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) red[(((warp * 2 + h) * MT + i) * 4 + e) * 32 + lane] = acc[h][i][e];
  __syncthreads();
  float mine[(OUTS + SK_THREADS - 1) / SK_THREADS];
#pragma unroll
  for (int r = 0; r < (OUTS + SK_THREADS - 1) / SK_THREADS; ++r) {
    const int o = threadIdx.x + r * SK_THREADS;
    float s = 0.f;
    if (o < OUTS) {
      // output o -> (m, nl): nl fastest so that the final stores are contiguous along N
      const int nl = o % SK_ROWS, m = o / SK_ROWS;
      const int h = nl >> 4, r16 = nl & 15, i = m >> 3, mm = m & 7;
      const int e = (mm & 1) + (r16 >= 8 ? 2 : 0), ln = (r16 & 7) * 4 + (mm >> 1);
#pragma unroll
      for (int w2 = 0; w2 < SK_WARPS; ++w2) s += red[(((w2 * 2 + h) * MT + i) * 4 + e) * 32 + ln];
    }
    mine[r] = s;
  }
  if (p.cluster > 1) {
    if (crank != 0) {
#pragma unroll
      for (int r = 0; r < (OUTS + SK_THREADS - 1) / SK_THREADS; ++r) {
        const int o = threadIdx.x + r * SK_THREADS;
        if (o < OUTS) sk_st_remote(part + (crank - 1) * OUTS + o, 0, mine[r]);
      }
    }
    sk_cluster_sync();                                         // release / acquire: the pushed partials are visible to rank 0
    if (crank != 0) return;
  }
#pragma unroll
  for (int r = 0; r < (OUTS + SK_THREADS - 1) / SK_THREADS; ++r) {
    const int o = threadIdx.x + r * SK_THREADS;
    if (o >= OUTS) continue;
    const int nl = o % SK_ROWS, m = o / SK_ROWS, n = n0 + nl;
    if (m >= p.M || n >= p.N) continue;
    float s = mine[r];
    for (int c = 1; c < p.cluster; ++c) s += part[(c - 1) * OUTS + o];
    float v = p.alpha * s;
    float* cp = p.C + (int64_t)m * p.ldc + n;
    if (p.beta != 0.f) v += p.beta * *cp;
    *cp = apply_epilogue(v, m, n, p.N, p.epilogue, p.ep);
  }
}

template <int MT>
int launch_skinny(const SkinnyParams& p, cudaStream_t st) {
  constexpr int OUTS = SK_ROWS * 8 * MT;
  const size_t smem = sizeof(float) * ((size_t)SK_WARPS * 2 * MT * 4 * 32 + (size_t)7 * OUTS);
  auto kern = gemm_skinny_tf32_kernel<MT>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { dasa_set_error("gemm_skinny attr", e); return DASA_ERR_CUDA; }
    attr_set = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)dasa_cdiv(p.N, SK_ROWS), (unsigned)p.cluster, 1);
  cfg.blockDim = dim3(SK_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1;
  attr[0].val.clusterDim.y = (unsigned)p.cluster;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, p);
  if (e != cudaSuccess) { dasa_set_error("gemm_skinny_tf32_kernel", e); return DASA_ERR_CUDA; }
  return DASA_OK;
}

}  // namespace

static int g_skinny_mode = -1;        // -1: read DASA_SKINNY (default 1), 0: off, 1: on where it wins, 2: every eligible shape (tests)
extern "C" int dasa_debug_gemm_skinny(int on) {
  g_skinny_mode = on < 0 ? 0 : (on > 2 ? 2 : on);
  return DASA_OK;
}

bool dasa_gemm_skinny_supported(int a_kmajor, int b_kmajor, int M, int N, int K, const float* A, int64_t lda, const float* B,
                                int64_t ldb) {
  if (g_skinny_mode < 0) { const char* e = getenv("DASA_SKINNY"); g_skinny_mode = e ? atoi(e) : 1; }
  if (!g_skinny_mode || !a_kmajor || !b_kmajor) return false;
  if (M < 1 || M > 32 || N < 16 || K < SK_CHUNK || (K % SK_CHUNK) != 0) return false;
  // measured (scripts/skinny_gemm.py, M = 20): 7.8-10 us here vs 11-14 us on the tcgen05 tile kernels (+ split-K reduce launch) while
  // each warp has at most ~2 chunks to walk; longer K walks per warp are latency-bound here and the tile kernels win (12 vs 15 us)
  if (g_skinny_mode != 2 && (int64_t)N * K > (int64_t)4608 * 1024) return false;
  return dasa_aligned16(A) && dasa_aligned16(B) && (lda % 4) == 0 && (ldb % 4) == 0;
}

int dasa_gemm_skinny(int M, int N, int K, float alpha, const float* A, int64_t lda, const float* B, int64_t ldb, float beta,
                     float* C, int64_t ldc, int epilogue, const EpiParams& ep, cudaStream_t st) {
  SkinnyParams p{};
  p.A = A; p.lda = lda; p.W = B; p.ldb = ldb; p.C = C; p.ldc = ldc; p.M = M; p.N = N; p.K = K;
  p.alpha = alpha; p.beta = beta; p.epilogue = epilogue; p.ep = ep;
  // CTAs along K: fill ~2 CTAs per SM, never more K slices than there are chunks per warp set
  const int nblocks = (int)dasa_cdiv(N, SK_ROWS);
  int cs = 1;
  while (cs < 8 && nblocks * cs * 2 <= 2 * DASA_NUM_SMS && (K / SK_CHUNK) >= SK_WARPS * cs * 2) cs *= 2;
  p.cluster = cs;
  const int mt = (M + 7) / 8;
  switch (mt) {
    case 1: return launch_skinny<1>(p, st);
    case 2: return launch_skinny<2>(p, st);
    case 3: return launch_skinny<3>(p, st);
    default: return launch_skinny<4>(p, st);
  }
}
