"""One MN-major and one K-major CTA-pair GEMM of the bi-LSTM weight-gradient shape (for ncu --set full captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dasa_b200 import lib, ops
lib.load()
M, N, K = 35700, 4096, 1024                       # rows, dW[N, K]
dY = torch.randn(M, N, device="cuda"); X = torch.randn(M, K, device="cuda"); dW = torch.zeros(N, K, device="cuda")
dYt, Xt = ops.transpose(dY), ops.transpose(X)
torch.cuda.synchronize()
torch.cuda.profiler.start()
ops.gemm(dY, N, 0, X, K, 0, dW, K, N, K, M, beta=1.0, precision=ops.PREC_TF32)                              # MN-major A and B
ops.gemm(dYt, dYt.stride(0), 1, Xt, Xt.stride(0), 1, dW, K, N, K, M, beta=1.0, precision=ops.PREC_TF32)     # K-major, transposed copies
torch.cuda.synchronize()
torch.cuda.profiler.stop()
