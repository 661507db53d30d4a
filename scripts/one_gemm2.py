import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dasa_b200 import ops
M, N, K = 896, 2304, 32
A = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda"); b = torch.randn(N, device="cuda"); C = torch.empty(M, N, device="cuda")
for _ in range(6):
    ops.gemm(A, K, 1, W, K, 1, C, N, M, N, K, epilogue=ops.EPI_BIAS, bias=b, precision=ops.PREC_TF32)
torch.cuda.synchronize()
print("ok")
