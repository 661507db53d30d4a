// Micro-experiment: fixed cost of launching N CTAs with large dynamic smem / TMEM allocation (graph-replayed).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int MODE>
__global__ void __launch_bounds__(128, 1) k(float* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t slot;
  if (MODE >= 2) {
    if (threadIdx.x / 32 == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(128u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  if (MODE >= 1) smem[threadIdx.x] = (unsigned char)threadIdx.x;
  if (threadIdx.x == 0 && out) out[blockIdx.x] = 1.f;
  if (MODE >= 2) {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x / 32 == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(128u) : "memory");
  }
}
template <int MODE>
void run(const char* name, int grid, size_t smem) {
  float* out; cudaMalloc(&out, 4096 * 4);
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaStream_t st; cudaStreamCreate(&st);
  for (int i = 0; i < 3; ++i) k<MODE><<<grid, 128, smem, st>>>(out);
  cudaGraph_t g; cudaGraphExec_t ge;
  cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal);
  for (int i = 0; i < 50; ++i) k<MODE><<<grid, 128, smem, st>>>(out);
  cudaStreamEndCapture(st, &g);
  cudaGraphInstantiate(&ge, g, 0);
  cudaGraphLaunch(ge, st); cudaStreamSynchronize(st);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0, st); cudaGraphLaunch(ge, st); cudaEventRecord(e1, st); cudaStreamSynchronize(st);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("%-28s grid=%4d smem=%6zu : %.2f us per launch (%s)\n", name, grid, smem, ms * 1e3 / 50, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  for (int grid : {1, 18, 126, 148, 296}) {
    run<0>("empty", grid, 0);
    run<1>("smem 197KB", grid, 197 * 1024);
    run<1>("smem 64KB", grid, 64 * 1024);
    run<2>("smem 197KB + tmem alloc", grid, 197 * 1024);
    run<2>("smem 32KB + tmem alloc", grid, 32 * 1024);
  }
  return 0;
}
