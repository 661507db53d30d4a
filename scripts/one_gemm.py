import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dasa_b200 import ops
M, N, K = [int(x) for x in sys.argv[1:4]] if len(sys.argv) > 3 else (900, 3072, 768)
A = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda"); b = torch.randn(N, device="cuda"); C = torch.empty(M, N, device="cuda")
for _ in range(5):
    ops.gemm(A, K, 1, W, K, 1, C, N, M, N, K, epilogue=ops.EPI_BIAS_GELU, bias=b, precision=ops.PREC_TF32)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(20):
        ops.gemm(A, K, 1, W, K, 1, C, N, M, N, K, epilogue=ops.EPI_BIAS_GELU, bias=b, precision=ops.PREC_TF32)
g.replay(); torch.cuda.synchronize()
e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
print("graph-replayed: %.2f us per GEMM, %.1f TFLOP/s" % (e0.elapsed_time(e1) * 1e3 / 20, 2 * M * N * K * 20 / e0.elapsed_time(e1) / 1e9))
