"""GPU helper for ncu: a few launches of the fused gate -> shift attention kernel at one shape (argv: B C)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dasa_b200 import ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
C = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
V, A = 36, 128
f = torch.rand(B, V, C + A, device="cuda"); gp = torch.randn(B, V, C, device="cuda")
t = torch.randn(B, C + A, device="cuda") * 0.05; kl = torch.randn(B, 5, device="cuda")
for _ in range(3):
    ops.gate_shift_attention_fwd(f, gp, t, kl, 5, 12)
torch.cuda.synchronize()
