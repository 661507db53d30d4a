"""GPU helper for ncu: one fp16 MN-major weight-gradient GEMM of the bi-LSTM (dW_hh: 4096 x 1024 over 19810 tokens)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dasa_b200 import ops
ops.set_precision("tf32")
rows, N, K = 19810, 4096, 1024
dy = (torch.randn(rows, N, device="cuda") * 0.3).half()
x = (torch.randn(rows, K, device="cuda") * 0.5).half()
dw = torch.zeros(N, K, device="cuda")
for _ in range(3):
    ops.linear_bwd_weight_f16(dy, x, dw, 1.0 / 256.0, True)
torch.cuda.synchronize()
