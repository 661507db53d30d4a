// tcgen05 (5th-gen tensor core) TF32 GEMM for the dense projections: C[M,N] = epi(alpha * A[M,K] * B[N,K]^T + beta*C).
// fp32 operands stay fp32 in HBM; TMA (cp.async.bulk.tensor, 128B swizzle, OOB zero fill) stages 128x32 / BNx32 tiles
// in shared memory as TFLOAT32, one elected thread issues tcgen05.mma.kind::tf32 (UMMA 128 x BN x 8) with the fp32
// accumulator living in TMEM, tcgen05.commit hands smem stages back to the TMA producer through mbarriers, and the
// four warps read the accumulator back with tcgen05.ld for the fused epilogue (bias / tanh / GELU / ReLU / sigmoid-gate).
// K-major ("Linear") operand layout only; the other layouts of dasa_gemm stay on the FFMA kernel.
#include <cuda.h>
#include <stdlib.h>
#include "common.cuh"
#include "gemm_common.cuh"

namespace {

constexpr int TC_BM = 128;
constexpr int TC_BK = 32;            // 32 fp32 = 128 bytes = one swizzle row
constexpr int TC_UMMA_K = 8;         // tf32: 32 bytes per MMA along K
constexpr int TC_THREADS = 128;

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
// multicast variant: the box lands at the same shared-memory offset of every CTA in cta_mask and signals each one's barrier
__device__ __forceinline__ void tma_load_2d_mc(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar, uint16_t cta_mask) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
               ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(cta_mask) : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(cta_mask) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// K-major, 128B-swizzled operand tile: rows of 128 bytes, 8-row swizzle atoms 1024 bytes apart (SBO); LBO unused.
__device__ __forceinline__ uint64_t make_smem_desc_sw128(const void* smem_ptr) {
  const uint32_t addr = smem_u32(smem_ptr);
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);          // start address, bits [0,14)
  d |= (uint64_t)0 << 16;                          // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                // stride byte offset, bits [32,46)
  d |= (uint64_t)1 << 46;                          // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                          // layout type SWIZZLE_128B
  return d;
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ unsigned long long g_tc_ts[64];
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define TC_STAMP(i) do { if ((p.debug & 2) && blockIdx.x == gridDim.x - 1 && blockIdx.y == gridDim.y - 1 && blockIdx.z == 0) g_tc_ts[i] = gtimer(); } while (0)

struct TcParams {
  int M, N, K;
  float alpha, beta;
  float* C; int64_t ldc;
  int epilogue; EpiParams ep;
  float* partial;        // split-K partial sums [S, M, N] or nullptr
  int k_per_split;       // multiple of TC_BK
  int debug;             // bit0: skip epilogue global stores (timing experiments only)
};

// MC = 1: launched as clusters of two CTAs along N (same 128 rows of A, neighbouring column tiles). Each CTA loads HALF of the
// A tile and multicasts it into both CTAs, so the L2 -> SM operand traffic per tile drops from (A + B) to (A/2 + B): the fp32
// operand stream (32 FLOP/B at 128x128) is what bounds this kernel, not the tensor pipe. A stage is recycled only after BOTH
// CTAs' MMAs have consumed it (tcgen05.commit multicast onto both `empty` barriers, which then count 2 arrivals).
template <int BN, int STAGES, int EPI, int MC = 0>
__global__ void __launch_bounds__(TC_THREADS, (STAGES <= 4 ? 2 : 1))
gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, TcParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  constexpr int A_BYTES = TC_BM * TC_BK * 4, B_BYTES = BN * TC_BK * 4, STAGE_BYTES = A_BYTES + B_BYTES;
  // 1024-byte alignment is required by the 128B swizzle; the dynamic smem base is aligned by the launch
  unsigned char* tiles = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps shared-space provenance (LDS/STS, not generic)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tiles + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * TC_BM, n0 = blockIdx.x * BN;
  const int kbeg = blockIdx.z * p.k_per_split;
  const int kend = min(p.K, kbeg + p.k_per_split);
  const int nkb = (p.debug & 4) ? 0 : (kend - kbeg + TC_BK - 1) / TC_BK;   // debug bit2: skip TMA + MMA

  if (threadIdx.x == 0) TC_STAMP(0);
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], MC ? 2 : 1); }
    mbar_init(tmem_full, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, BN);       // BN fp32 accumulator columns x 128 lanes
  tc_fence_before();
  __syncthreads();
  if constexpr (MC) cluster_sync_all();           // the peer's barriers exist before any multicast load / commit reaches them
  tc_fence_after();
  const uint32_t crank = MC ? cluster_ctarank() : 0u;
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) TC_STAMP(1);

  if (warp == 0 && lane == 0) {
    // ---------------------------------------------------------------- TMA producer
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % STAGES;
      const uint32_t ph = (kb / STAGES) & 1;
      mbar_wait(&empty_bar[s], ph ^ 1);
      mbar_expect_tx(&full_bar[s], STAGE_BYTES);
      unsigned char* a_dst = tiles + s * STAGE_BYTES;
      if constexpr (MC) {                       // my 64-row half of A goes to both CTAs; the peer supplies the other half
        tma_load_2d_mc(a_dst + crank * (A_BYTES / 2), &tmA, kbeg + kb * TC_BK, m0 + (int)crank * (TC_BM / 2), &full_bar[s], (uint16_t)3);
      } else {
        tma_load_2d(a_dst, &tmA, kbeg + kb * TC_BK, m0, &full_bar[s]);
      }
      tma_load_2d(a_dst + A_BYTES, &tmB, kbeg + kb * TC_BK, n0, &full_bar[s]);
      if (kb == 0) TC_STAMP(2);
    }
    TC_STAMP(3);
  } else if (warp == 1 && lane == 0) {
    // ---------------------------------------------------------------- MMA issuer (single thread)
    // instruction descriptor: D=f32, A=B=tf32, both K-major, N=BN, M=128
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % STAGES;
      const uint32_t ph = (kb / STAGES) & 1;
      mbar_wait(&full_bar[s], ph);
      if (kb == 0) TC_STAMP(4);
      tc_fence_after();
      unsigned char* a_src = tiles + s * STAGE_BYTES;
      const uint64_t da = make_smem_desc_sw128(a_src);
      const uint64_t db = make_smem_desc_sw128(a_src + A_BYTES);
#pragma unroll
      for (int k = 0; k < TC_BK / TC_UMMA_K; ++k) {
        // advance the start address by k*32 bytes inside the 128-byte swizzled row (address field is in 16-byte units)
        umma_tf32(tmem_base, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0 ? 1u : 0u);
      }
      if constexpr (MC) umma_commit_mc(&empty_bar[s], (uint16_t)3);   // frees the stage in BOTH CTAs' books
      else umma_commit(&empty_bar[s]);  // smem stage free once these MMAs have consumed it
    }
    umma_commit(tmem_full);            // accumulator complete
    TC_STAMP(5);
  }

  // -------------------------------------------------------------------- epilogue: all four warps
  __syncwarp();
  mbar_wait(tmem_full, 0);
  tc_fence_after();
  if (threadIdx.x == 64) TC_STAMP(6);
  // Phase 1: TMEM -> shared. Thread = tile row (TMEM lane); the pipeline smem is idle now (all MMAs retired, all TMA
  // loads consumed) and is reused as a [128][BN+4] fp32 staging tile (the +4 keeps the 128-bit row writes conflict-free).
  constexpr int LDS = BN + 4;
  float* stage = reinterpret_cast<float*>(tiles);
  if (!(p.debug & 8)) {   // debug bit3: skip the whole epilogue
    float* srow = stage + (warp * 32 + lane) * LDS;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t r[32];
      tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, r);
#pragma unroll
      for (int q = 0; q < 8; ++q)
        *reinterpret_cast<float4*>(srow + c0 + 4 * q) =
            make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]), __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3]));
    }
  }
  __syncwarp();     // each warp re-reads only the 32 rows it staged itself
  // Phase 2: shared -> global, one (or two) full tile rows per warp instruction: every global access of the epilogue
  // (C, beta*C, bias, gate source / gate output, drop mask) is coalesced.
  constexpr int LPR = BN / 4;          // lanes per row
  constexpr int RPI = 32 / LPR;        // rows per warp instruction
  const int col4 = lane % LPR, rsub = lane / LPR;
  const int n = n0 + 4 * col4;
  const bool vec_ok = ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) && (n + 3 < p.N);
  float bias4[4] = {0.f, 0.f, 0.f, 0.f};      // per-column, loop invariant: loaded once (not once per row)
  if constexpr (epi_has_bias<EPI>()) {
    if (p.ep.bias != nullptr && p.partial == nullptr) {
#pragma unroll
      for (int e = 0; e < 4; ++e) if (n + e < p.N) bias4[e] = __ldg(p.ep.bias + n + e);
    }
  }
  const float alpha = p.alpha, beta = p.beta;
  if (!(p.debug & 9)) {
#pragma unroll 4
    for (int rr = 0; rr < 32; rr += RPI) {
      const int r = warp * 32 + rr + rsub;
      const int m = m0 + r;
      if (m >= p.M || n >= p.N) continue;
      const float4 a4 = *reinterpret_cast<const float4*>(stage + r * LDS + 4 * col4);
      float v[4] = {a4.x, a4.y, a4.z, a4.w};
      if (p.partial != nullptr) {
        float* dst = p.partial + ((int64_t)blockIdx.z * p.M + m) * p.N + n;
        if (((p.N & 3) == 0) && (n + 3 < p.N)) *reinterpret_cast<float4*>(dst) = a4;
        else
#pragma unroll
          for (int e = 0; e < 4; ++e) if (n + e < p.N) dst[e] = v[e];
        continue;
      }
      float* crow = p.C + (int64_t)m * p.ldc + n;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (n + e < p.N) {
          float x = alpha * v[e];
          if (beta != 0.f) x += beta * crow[e];
          v[e] = apply_activation_t<EPI>(x + bias4[e], m, n + e, p.N, p.ep);
        }
      }
      if (vec_ok) *reinterpret_cast<float4*>(crow) = make_float4(v[0], v[1], v[2], v[3]);
      else
#pragma unroll
        for (int e = 0; e < 4; ++e) if (n + e < p.N) crow[e] = v[e];
    }
  }
  if (threadIdx.x == 64) TC_STAMP(7);
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, BN);
  if (threadIdx.x == 32) TC_STAMP(8);
  if constexpr (MC) cluster_sync_all();           // the peer may still commit onto / multicast into this CTA's shared memory
}

__global__ void tc_splitk_reduce_kernel(const float* __restrict__ partial, int S, int M, int N, float alpha, float beta,
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

struct TcPlan { int bn; int splits; int k_per_split; };

TcPlan plan_tc(int M, int N, int K) {
  // Measured on B200 (scripts/gemm_sweep.py, gemm_timeline.py): a launch costs ~14 us fixed, a k-block 0.1-0.3 us depending on
  // how many CTAs share the L2 ports; a second wave or a split-K reduction pass costs more than it gains unless K is long.
  TcPlan pl;
  const int64_t tm = dasa_cdiv(M, TC_BM);
  const int64_t t128 = tm * dasa_cdiv(N, 128), t64 = tm * dasa_cdiv(N, 64);
  if (N <= 64) pl.bn = 64;
  else if (t128 >= 120) pl.bn = 128;
  else if (t64 <= 2 * DASA_NUM_SMS) pl.bn = 64;     // two CTAs are resident per SM
  else pl.bn = 128;
  const int64_t tiles = tm * dasa_cdiv(N, pl.bn);
  const int nkb = (int)dasa_cdiv(K, TC_BK);
  int s = 1;
  if (tiles * 2 <= 2 * DASA_NUM_SMS && nkb >= 64) {     // two CTAs are resident per SM (4-stage variant)
    s = (int)(2 * DASA_NUM_SMS / tiles);
    s = s > nkb / 16 ? nkb / 16 : s;
    s = s < 1 ? 1 : (s > 16 ? 16 : s);
  }
  int kps = (int)dasa_cdiv(nkb, s) * TC_BK;
  pl.splits = (int)dasa_cdiv(K, kps);
  pl.k_per_split = kps;
  return pl;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() { return reinterpret_cast<EncodeTiledFn>(dasa_tensormap_encoder()); }

// 2-D fp32 tensor [rows, K] with row stride ld (elements), K contiguous; box = [box_rows x 32]
bool make_map(CUtensorMap* map, const float* base, int64_t rows, int64_t K, int64_t ld, int box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (enc == nullptr) return false;
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

template <int BN, int STAGES, int EPI>
int launch_tc_e(const CUtensorMap& ta, const CUtensorMap& tb, const TcParams& p, int splits, cudaStream_t st) {
  constexpr size_t smem = (size_t)STAGES * (TC_BM * TC_BK * 4 + BN * TC_BK * 4) + 1024 + 256;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tf32_kernel<BN, STAGES, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { dasa_set_error("gemm_tf32 attr", e); return DASA_ERR_CUDA; }
    attr_set = true;
  }
  dim3 grid((unsigned)dasa_cdiv(p.N, BN), (unsigned)dasa_cdiv(p.M, TC_BM), (unsigned)splits);
  gemm_tf32_kernel<BN, STAGES, EPI><<<grid, TC_THREADS, smem, st>>>(ta, tb, p);
  return dasa_check_launch("gemm_tf32_kernel");
}

// cluster (2,1,1) launch of the A-multicast variant; ta must have been built with 64-row boxes
template <int BN, int STAGES, int EPI>
int launch_tc_mc_e(const CUtensorMap& ta, const CUtensorMap& tb, const TcParams& p, cudaStream_t st) {
  constexpr size_t smem = (size_t)STAGES * (TC_BM * TC_BK * 4 + BN * TC_BK * 4) + 1024 + 256;
  auto kern = gemm_tf32_kernel<BN, STAGES, EPI, 1>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { dasa_set_error("gemm_tf32<mc> attr", e); return DASA_ERR_CUDA; }
    attr_set = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)dasa_cdiv(p.N, BN), (unsigned)dasa_cdiv(p.M, TC_BM), 1);
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ta, tb, p);
  if (e != cudaSuccess) { dasa_set_error("gemm_tf32_kernel<mc>", e); return DASA_ERR_CUDA; }
  return DASA_OK;
}

template <int BN, int STAGES>
int launch_tc_mc(const CUtensorMap& ta, const CUtensorMap& tb, const TcParams& p, cudaStream_t st) {
  switch (p.epilogue) {
    case DASA_EPI_BIAS: return launch_tc_mc_e<BN, STAGES, DASA_EPI_BIAS>(ta, tb, p, st);
    case DASA_EPI_BIAS_GELU: return launch_tc_mc_e<BN, STAGES, DASA_EPI_BIAS_GELU>(ta, tb, p, st);
    case DASA_EPI_GATE: return launch_tc_mc_e<BN, STAGES, DASA_EPI_GATE>(ta, tb, p, st);
    case DASA_EPI_NONE: return launch_tc_mc_e<BN, STAGES, DASA_EPI_NONE>(ta, tb, p, st);
    default: return DASA_ERR_UNSUPPORTED;
  }
}

template <int BN, int STAGES>
int launch_tc(const CUtensorMap& ta, const CUtensorMap& tb, const TcParams& p, int splits, cudaStream_t st) {
  // split-K launches only store raw partial sums: one (NONE) instantiation serves them all
  const int epi = (p.partial != nullptr) ? DASA_EPI_NONE : p.epilogue;
  switch (epi) {
    case DASA_EPI_BIAS: return launch_tc_e<BN, STAGES, DASA_EPI_BIAS>(ta, tb, p, splits, st);
    case DASA_EPI_BIAS_TANH: return launch_tc_e<BN, STAGES, DASA_EPI_BIAS_TANH>(ta, tb, p, splits, st);
    case DASA_EPI_BIAS_GELU: return launch_tc_e<BN, STAGES, DASA_EPI_BIAS_GELU>(ta, tb, p, splits, st);
    case DASA_EPI_BIAS_RELU: return launch_tc_e<BN, STAGES, DASA_EPI_BIAS_RELU>(ta, tb, p, splits, st);
    case DASA_EPI_GATE: return launch_tc_e<BN, STAGES, DASA_EPI_GATE>(ta, tb, p, splits, st);
    case DASA_EPI_TANH: return launch_tc_e<BN, STAGES, DASA_EPI_TANH>(ta, tb, p, splits, st);
    default: return launch_tc_e<BN, STAGES, DASA_EPI_NONE>(ta, tb, p, splits, st);
  }
}

}  // namespace

// A-multicast variant: measured NO gain on B200 (M=20300 N=3072 K=768: 315 us vs 297 us; N=768 K=3072: 209 vs 205 us) - at 46-67 %
// of the TF32 peak the kernel is bound by the per-tile prologue / epilogue, not by the L2 -> SM operand stream - so it is off
// unless DASA_TC_MULTICAST=1 or dasa_debug_gemm_multicast(1); kept (and tested) as the base of a persistent multicast kernel.
static int g_use_mc = -1;
extern "C" int dasa_debug_gemm_multicast(int on) {
  g_use_mc = on ? 1 : 0;
  return DASA_OK;
}

extern "C" int dasa_debug_tc_timestamps(unsigned long long* host16) {
  return cudaMemcpyFromSymbol(host16, g_tc_ts, 16 * sizeof(unsigned long long)) == cudaSuccess ? 0 : DASA_ERR_CUDA;
}

bool dasa_gemm_tc_supported(int a_kmajor, int b_kmajor, int M, int N, int K, const float* A, int64_t lda, const float* B,
                            int64_t ldb, const float* C, int64_t ldc) {
  (void)C; (void)ldc;
  if (!a_kmajor || !b_kmajor) return false;                 // Linear layout only (for now)
  if (M <= 0 || N <= 0 || K < TC_BK) return false;
  if (!dasa_aligned16(A) || !dasa_aligned16(B) || (lda % 4) != 0 || (ldb % 4) != 0) return false;   // TMA stride/base rules
  return get_encode_fn() != nullptr;
}

size_t dasa_gemm_tc_workspace(int M, int N, int K) {
  if (dasa_gemm_pair_plan(M, N, K)) return 0;
  TcPlan pl = plan_tc(M, N, K);
  return pl.splits > 1 ? (size_t)pl.splits * M * N * sizeof(float) : 0;
}

int dasa_gemm_tc(int a_kmajor, int b_kmajor, int M, int N, int K, float alpha, const float* A, int64_t lda, const float* B,
                 int64_t ldb, float beta, float* C, int64_t ldc, int epilogue, const EpiParams& ep, void* workspace,
                 size_t workspace_bytes, cudaStream_t st) {
  (void)a_kmajor; (void)b_kmajor;
  // many-tile problems: persistent CTA-pair kernel (gemm_tc2.cu), 256 x BN tiles, half the L2 -> SM operand bytes per FLOP
  if (const int bn2 = dasa_gemm_pair_plan(M, N, K))
    return dasa_gemm_tc_pair(bn2, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, epilogue, ep, st);
  TcPlan pl = plan_tc(M, N, K);
  float* partial = nullptr;
  if (pl.splits > 1) {
    const size_t need = (size_t)pl.splits * M * N * sizeof(float);
    if (workspace == nullptr || workspace_bytes < need) { pl.splits = 1; pl.k_per_split = (int)dasa_cdiv(K, TC_BK) * TC_BK; }
    else partial = static_cast<float*>(workspace);
  }
  static int dbg = -1;
  if (dbg < 0) { const char* e = getenv("DASA_TC_DEBUG"); dbg = e ? atoi(e) : 0; }
  if (g_use_mc < 0) { const char* e = getenv("DASA_TC_MULTICAST"); g_use_mc = e ? atoi(e) : 0; }
  const int use_mc = g_use_mc;
  TcParams p{M, N, K, alpha, beta, C, ldc, epilogue, ep, partial, pl.k_per_split, dbg};
  CUtensorMap ta, tb;
  // many-wave 128x128 problems with an even number of column tiles: pairs of CTAs share their A rows through TMA multicast
  const int64_t ntn = dasa_cdiv(N, 128);
  if (use_mc && pl.bn == 128 && pl.splits == 1 && (ntn % 2) == 0 && dasa_cdiv(M, TC_BM) * ntn > 2 * DASA_NUM_SMS &&
      (epilogue == DASA_EPI_BIAS || epilogue == DASA_EPI_BIAS_GELU || epilogue == DASA_EPI_GATE || epilogue == DASA_EPI_NONE)) {
    if (!make_map(&ta, A, M, K, lda, TC_BM / 2) || !make_map(&tb, B, N, K, ldb, 128)) return DASA_ERR_UNSUPPORTED;
    return launch_tc_mc<128, 3>(ta, tb, p, st);
  }
  if (!make_map(&ta, A, M, K, lda, TC_BM) || !make_map(&tb, B, N, K, ldb, pl.bn)) return DASA_ERR_UNSUPPORTED;
  // 3 x 32 KB / 4 x 24 KB stages = 96 KB of pipeline per CTA: two CTAs co-reside on an SM, so one CTA's epilogue and
  // prologue overlap the other's main loop and a 192-tile problem needs no second wave.
  // otherwise (one CTA per SM anyway) the deep 6/8-stage ring keeps 192 KB of loads in flight per SM.
  const int64_t ctas = dasa_cdiv(M, TC_BM) * dasa_cdiv(N, pl.bn) * pl.splits;
  int rc;
  if (ctas > DASA_NUM_SMS)
    rc = (pl.bn == 128) ? launch_tc<128, 3>(ta, tb, p, pl.splits, st) : launch_tc<64, 4>(ta, tb, p, pl.splits, st);
  else
    rc = (pl.bn == 128) ? launch_tc<128, 6>(ta, tb, p, pl.splits, st) : launch_tc<64, 8>(ta, tb, p, pl.splits, st);
  if (rc != DASA_OK) return rc;
  if (partial != nullptr) {
    const int64_t total = (int64_t)M * N;
    tc_splitk_reduce_kernel<<<(unsigned)dasa_cdiv(total, 256), 256, 0, st>>>(partial, pl.splits, M, N, alpha, beta, C, ldc,
                                                                              epilogue, ep);
    rc = dasa_check_launch("tc_splitk_reduce_kernel");
  }
  return rc;
}
