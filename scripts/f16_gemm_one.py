"""GPU helper for ncu: one fp16 GELU GEMM of the frozen stack (20300 x 3072 x 768, fp16 out)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dasa_b200 import ops
ops.set_precision("tf32")
M, N, K = 20300, 3072, 768
if len(sys.argv) >= 5:
    M, N, K = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
epi = ops.EPI_BIAS_GELU if (len(sys.argv) < 2 or sys.argv[1] == "gelu") else ops.EPI_BIAS
x = (torch.randn(M, K, device="cuda") * 0.5).half()
w = (torch.randn(N, K, device="cuda") * 0.05).half()
b = torch.randn(N, device="cuda")
for _ in range(3):
    ops.linear_f16(x, w, b, epi, True)
torch.cuda.synchronize()
