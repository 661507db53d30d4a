import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dasa_b200 import ops
def t(M, N, K, reps=30):
    A = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda"); b = torch.randn(N, device="cuda"); C = torch.empty(M, N, device="cuda")
    for _ in range(3): ops.gemm(A, K, 1, W, K, 1, C, N, M, N, K, epilogue=ops.EPI_BIAS, bias=b, precision=ops.PREC_TF32)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps): ops.gemm(A, K, 1, W, K, 1, C, N, M, N, K, epilogue=ops.EPI_BIAS, bias=b, precision=ops.PREC_TF32)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps
print("DEBUG=%s  1cta(128x128x768): %.2f us   126cta(896x2304x32): %.2f us   126cta(896x2304x768): %.2f us" % (os.environ.get("DASA_TC_DEBUG"), t(128, 128, 768), t(896, 2304, 32), t(896, 2304, 768)))
