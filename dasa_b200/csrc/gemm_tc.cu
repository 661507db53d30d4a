// tcgen05 (TF32) GEMM — placeholder until the tensor-core kernel lands; reports "unsupported" so dasa_gemm uses FFMA.
#include "common.cuh"
#include "gemm_common.cuh"
bool dasa_gemm_tc_supported(int, int, int, int, int, const float*, int64_t, const float*, int64_t, const float*, int64_t) { return false; }
size_t dasa_gemm_tc_workspace(int, int, int) { return 0; }
int dasa_gemm_tc(int, int, int, int, int, float, const float*, int64_t, const float*, int64_t, float, float*, int64_t, int,
                 const EpiParams&, void*, size_t, cudaStream_t) { return DASA_ERR_UNSUPPORTED; }
