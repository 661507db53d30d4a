"""GPU helper: shift-attention forward at batch B against the HBM roofline (L2 flushed between launches)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dasa_b200 import ops
B, V = int(sys.argv[1]) if len(sys.argv) > 1 else 1024, 36
F = int(sys.argv[2]) if len(sys.argv) > 2 else 2176
f = torch.rand(B, V, F, device="cuda"); t = torch.randn(B, F, device="cuda") * 0.05; kl = torch.randn(B, 5, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = []
for i in range(8):
    flush.zero_()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record(); ops.row_attention_fwd(f, t, None, 5, 12, kl); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ts.sort(); ms = ts[len(ts) // 2]
byt = 4 * (B * V * F + 2 * B * F + B * V + B * 5)
print("B=%d F=%d pipe_min_b=%s shift_attention_fwd: %.3f ms  %.1f GB/s" % (B, F, os.environ.get("DASA_RA_PIPE_MIN_B", "64"), ms, byt / ms / 1e6))
