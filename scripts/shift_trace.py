"""GPU helper: per-sample hand-off timeline of the pipelined view-attention kernel (SM clock cycles, CTA-local)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dasa_b200 import ops, lib
B, V, F = int(sys.argv[1]) if len(sys.argv) > 1 else 4096, 36, 2176
f = torch.rand(B, V, F, device="cuda"); t = torch.randn(B, F, device="cuda") * 0.05; kl = torch.randn(B, 5, device="cuda")
GATED = len(sys.argv) > 2 and sys.argv[2] == "gate"      # the fused gate variant (stamp 7 = gate warps start)
gp = torch.randn(B, V, F - 128, device="cuda") if GATED else None
run = (lambda: ops.gate_shift_attention_fwd(f, gp, t, kl, 5, 12)) if GATED else (lambda: ops.row_attention_fwd(f, t, None, 5, 12, kl))
for _ in range(3):
    run()
torch.cuda.synchronize()
NCL = 15                         # clusters resident on this part (cudaOccupancyMaxActiveClusters)
nper = (B + NCL - 1) // NCL
tr = torch.zeros(NCL * 8, nper, 8, dtype=torch.int64, device="cuda")
lib.load().dasa_debug_row_attention_trace(tr.data_ptr())
run()
torch.cuda.synchronize()
lib.load().dasa_debug_row_attention_trace(None)
tr = tr.cpu().double()
names = ["P issue", "D full", "D pushed", "S zfull", "S publish", "W start", "W release"]
for cta in (0, 3, 60):
    x = tr[cta]
    ok = (x[:, :7] > 0).all(1)
    n = int(ok.sum())
    x = x[:n]
    mid = slice(20, n - 20)
    print("CTA %d: %d samples; period per sample (cycles): %s" % (cta, n, ["%.0f" % float((x[mid, k][1:] - x[mid, k][:-1]).mean()) for k in range(7)]))
    for k in range(6):
        d = x[mid, k + 1] - x[mid, k]
        print("   %-10s -> %-10s mean %7.0f  p10 %7.0f  p90 %7.0f" % (names[k], names[k + 1], float(d.mean()), float(d.quantile(0.1)), float(d.quantile(0.9))))
    if GATED:
        d = x[mid, 7] - x[mid, 0]
        print("   P issue -> G start (full + gfull) mean %7.0f  p10 %7.0f  p90 %7.0f" % (float(d.mean()), float(d.quantile(0.1)), float(d.quantile(0.9))))
        d = x[mid, 1] - x[mid, 7]
        print("   G start -> D gated mean %7.0f  p10 %7.0f  p90 %7.0f" % (float(d.mean()), float(d.quantile(0.1)), float(d.quantile(0.9))))
    NS = 3 if GATED else 5
    d = x[NS:, 0][20:-20] - x[:-NS, 6][20:-20]
    print("   W release(i) -> P issue(i+%d) mean %7.0f" % (NS, float(d.mean())))
    d = x[mid, 6] - x[mid, 0]
    print("   stage residency (P issue -> W release) mean %7.0f" % float(d.mean()))
