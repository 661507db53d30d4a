import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dasa_b200 import ops
def t(M, N, K, reps=20):
    A = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda"); b = torch.randn(N, device="cuda"); C = torch.empty(M, N, device="cuda")
    for _ in range(3): ops.gemm(A, K, 1, W, K, 1, C, N, M, N, K, epilogue=ops.EPI_BIAS, bias=b, precision=ops.PREC_TF32)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps): ops.gemm(A, K, 1, W, K, 1, C, N, M, N, K, epilogue=ops.EPI_BIAS, bias=b, precision=ops.PREC_TF32)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    print("M=%5d N=%5d K=%5d : %7.2f us  %6.1f TFLOP/s" % (M, N, K, us, 2 * M * N * K / us / 1e6), flush=True)
print("DASA_TC_DEBUG=", os.environ.get("DASA_TC_DEBUG"))
for K in (32, 64, 128, 256, 768, 3072):
    t(896, 2304, K)
for (M, N) in ((128, 128), (128, 2304), (896, 128), (900, 768), (900, 3072), (896, 768), (896, 3072), (1792, 2304)):
    t(M, N, 768)
