// GPU probe (not part of the library): pure streaming rate of the TMA engine for the tile shapes the view-attention kernel
// could stage: persistent CTAs, NS-stage ring, consumer = one thread that waits and releases.  nvcc -arch=sm_100a -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t it = 0; it < (1u << 26); ++it) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (done) return;
  }
  __trap();
}
__device__ __forceinline__ void tma3(void* dst, const CUtensorMap* m, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void bulk(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// mode 0: f32 map, nbox boxes [rows x boxw]; mode 1: f64 map, one box [rows x chunk/2]; mode 2: per-row bulk copies;
// mode 3: one contiguous bulk copy of rows*chunk*4 bytes
struct P { const float* x; int B, rows, D, chunk, nbox, boxw, ns, mode, nsplit; };

__global__ void __launch_bounds__(64, 1) probe(const __grid_constant__ CUtensorMap m32, const __grid_constant__ CUtensorMap m64, P p) {
  extern __shared__ __align__(128) unsigned char smem[];
  float* tile = (float*)smem;
  const size_t stage = (size_t)p.rows * p.chunk;
  uint64_t* full = (uint64_t*)(smem + (size_t)p.ns * stage * 4);
  uint64_t* empty = full + p.ns;
  const int units = p.B * p.nsplit;                 // unit = (sample, channel slice)
  if (threadIdx.x == 0) {
    for (int i = 0; i < p.ns; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  int n = 0;
  for (int u = blockIdx.x; u < units; u += gridDim.x) ++n;
  if (threadIdx.x == 0) {
    for (int j = 0; j < n; ++j) {
      const int u = blockIdx.x + j * gridDim.x;
      const int b = u / p.nsplit, sl = u % p.nsplit, st = j % p.ns;
      if (j >= p.ns) mbar_wait(&empty[st], (j / p.ns - 1) & 1);
      float* dst = tile + st * stage;
      mbar_expect_tx(&full[st], (uint32_t)(stage * 4));
      if (p.mode == 0) {
        for (int sb = 0; sb < p.nbox; ++sb) tma3(dst + (size_t)sb * p.rows * p.boxw, &m32, sl * p.chunk + sb * p.boxw, 0, b, &full[st]);
      } else if (p.mode == 1) {
        tma3(dst, &m64, sl * p.chunk / 2, 0, b, &full[st]);
      } else if (p.mode == 2) {
        for (int r = 0; r < p.rows; ++r)
          bulk(dst + (size_t)r * p.chunk, p.x + ((size_t)b * p.rows + r) * p.D + sl * p.chunk, p.chunk * 4, &full[st]);
      } else {
        bulk(dst, p.x + (size_t)b * p.rows * p.D + (size_t)sl * stage, (uint32_t)(stage * 4), &full[st]);
      }
    }
  } else if (threadIdx.x == 32) {
    for (int j = 0; j < n; ++j) {
      const int st = j % p.ns;
      mbar_wait(&full[st], (j / p.ns) & 1);
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[st])) : "memory");
    }
  }
}

typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                        const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int B = 4096, rows = 36, D = 2176;
  float* x;
  cudaMalloc(&x, (size_t)B * rows * D * 4);
  cudaMemset(x, 0, (size_t)B * rows * D * 4);
  char* flush;
  cudaMalloc(&flush, 256 << 20);
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  Enc enc = (Enc)fp;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  struct Cfg { int mode, nsplit, ns, grid; };
  Cfg cfgs[] = {{0, 8, 5, 148}, {0, 8, 5, 120}, {1, 8, 5, 148}, {2, 8, 5, 148}, {3, 8, 5, 148}, {3, 8, 5, 120}, {0, 8, 3, 148}, {1, 8, 3, 148},
                {0, 4, 2, 148}, {1, 4, 2, 148}, {3, 4, 2, 148}, {0, 16, 10, 148}, {1, 16, 10, 148}, {3, 16, 10, 148}, {0, 8, 5, 296}, {1, 8, 2, 296}};
  for (Cfg c : cfgs) {
    P p{x, B, rows, D, D / c.nsplit, 0, 0, c.ns, c.mode, c.nsplit};
    p.nbox = (p.chunk + 255) / 256;
    p.boxw = p.chunk / p.nbox;
    CUtensorMap m32, m64;
    cuuint64_t d32[3] = {(cuuint64_t)D, (cuuint64_t)rows, (cuuint64_t)B}, s32[2] = {(cuuint64_t)D * 4, (cuuint64_t)D * rows * 4};
    cuuint32_t b32[3] = {(cuuint32_t)p.boxw, (cuuint32_t)rows, 1}, es[3] = {1, 1, 1};
    CUresult r1 = enc(&m32, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, x, d32, s32, b32, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    cuuint64_t d64[3] = {(cuuint64_t)D / 2, (cuuint64_t)rows, (cuuint64_t)B};
    cuuint32_t b64[3] = {(cuuint32_t)(p.chunk / 2), (cuuint32_t)rows, 1};
    CUresult r2 = (p.chunk / 2 <= 256) ? enc(&m64, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, x, d64, s32, b64, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) : CUDA_ERROR_INVALID_VALUE;
    if ((c.mode == 0 && r1 != CUDA_SUCCESS) || (c.mode == 1 && r2 != CUDA_SUCCESS)) { printf("mode %d nsplit %d: encode failed %d %d\n", c.mode, c.nsplit, r1, r2); continue; }
    if (c.mode == 1 && r2 != CUDA_SUCCESS) continue;
    if (r2 != CUDA_SUCCESS) m64 = m32;
    const size_t smem = (size_t)c.ns * rows * p.chunk * 4 + 16 * c.ns + 64;
    float best = 1e9;
    for (int it = 0; it < 5; ++it) {
      cudaMemsetAsync(flush, it, 256 << 20);
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0);
      probe<<<c.grid, 64, smem>>>(m32, m64, p);
      cudaEventRecord(e1);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (ms < best) best = ms;
    }
    printf("mode %d nsplit %2d (chunk %4d, nbox %d) stages %2d grid %3d smem %6zu: %.3f ms  %.0f GB/s\n", c.mode, c.nsplit, p.chunk, p.nbox, c.ns,
           c.grid, smem, best, (double)B * rows * D * 4 / best / 1e6);
  }
  return 0;
}
