"""GPU helper: time the persistent decoder-rollout kernel (fwd / bwd) at the benchmark geometry and print the per-phase SM-clock
breakdown of CTA 0 (dasa_debug_decoder_phase_clocks)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dasa_b200 import functions as Fn, lib, ops
from dasa_b200.config import FULL

B = int(sys.argv[1]) if len(sys.argv) > 1 else 20
T = int(sys.argv[2]) if len(sys.argv) > 2 else 35
L = int(sys.argv[3]) if len(sys.argv) > 3 else 80
cfg = FULL
dev = "cuda"
ops.set_precision("tf32")
H, E, F, D, k, V = cfg.hidden, cfg.action_emb, cfg.feat, cfg.ctx_dim, cfg.shift_kernel, cfg.views
g = torch.Generator(device=dev).manual_seed(0)
rn = lambda *s, sc=1.0: torch.randn(*s, device=dev, generator=g) * sc
w = dict(w_in=rn(F, H, sc=H ** -0.5), w_shift=rn(k, H, sc=H ** -0.5), b_shift=rn(k, sc=0.1), w_ih=rn(4 * H, E + F, sc=(E + F) ** -0.5),
         w_hh=rn(4 * H, H, sc=H ** -0.5), b_ih=rn(4 * H, sc=0.1), b_hh=rn(4 * H, sc=0.1), w_att_in=rn(D, H, sc=H ** -0.5),
         w_att_out=rn(H, D + H, sc=(D + H) ** -0.5))
for v in w.values():
    v.requires_grad_(True)
emb, feat, ctx = torch.tanh(rn(T, B, E)), rn(T, B, V, F, sc=0.5).abs(), rn(T, B, L, D, sc=0.5)
mask = torch.zeros(B, L, dtype=torch.uint8, device=dev)
h0, c0 = torch.tanh(rn(B, H)), rn(B, H, sc=0.5)
m1 = (torch.rand(T, B, H, device=dev) > 0.5).to(torch.uint8)
m2 = (torch.rand(T, B, H, device=dev) > 0.5).to(torch.uint8)
gh = rn(T, B, H)
feat.requires_grad_(True); ctx.requires_grad_(True); emb.requires_grad_(True)
Fn.defer_weight_grads(True)

def clocks():
    buf = (ctypes.c_longlong * 33)()
    n = lib.load().dasa_debug_decoder_phase_clocks(buf, 33)
    return [int(x) for x in buf][:n]

def run(bwd):
    ht, h1, c = Fn.DecoderRolloutFn.apply(emb, feat, ctx, mask, h0, c0, m1, m2, 2.0, w["w_in"], w["w_shift"], w["b_shift"], w["w_ih"],
                                          w["w_hh"], w["b_ih"], w["b_hh"], w["w_att_in"], w["w_att_out"], 12)
    torch.cuda.synchronize()
    cf = clocks()
    cb = None
    if bwd:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        (ht * gh).sum().backward()
        e1.record()
        torch.cuda.synchronize()
        cb = clocks()
        Fn.discard_weight_grads()
    return cf, cb

for _ in range(2):
    run(True)
ts = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    with torch.no_grad():
        Fn.DecoderRolloutFn.apply(emb, feat, ctx, mask, h0, c0, m1, m2, 2.0, w["w_in"], w["w_shift"], w["b_shift"], w["w_ih"], w["w_hh"],
                                  w["b_ih"], w["b_hh"], w["w_att_in"], w["w_att_out"], 12)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print("B=%d T=%d L=%d  forward launch (incl. host-side buffer allocation): %s ms -> %.1f us/action" % (B, T, L, ["%.3f" % x for x in ts], 1e3 * min(ts) / T))
cf, cb = run(True)
names_f = ["P1 tk", "P2a dots", "P2b shift+wsum", "P3 gates+cell", "P4 t2", "P5a dots", "P5b softmax+wsum", "P6 h~"]
names_b = ["B6 dcat", "B5a dalpha", "B5b dctx", "B4 dh1+cell", "B3 dxh", "B2a dq", "B2b dfeat", "B1 dh~"]
mhz = 1965.0
for nm, c, names in (("forward", cf, names_f), ("backward", cb, names_b)):
    print(nm, "phase times of CTA 0 in us (at %.0f MHz), actions 1..3 of the launch:" % mhz)
    for ph in range(8):
        row = []
        for i in range(1, 4):
            idx = 1 + 8 * i + ph
            row.append((c[idx] - c[idx - 1]) / mhz)
        print("  %-18s %s" % (names[ph], "  ".join("%6.2f" % x for x in row)))
    print("  per action: %s" % "  ".join("%6.2f" % ((c[1 + 8 * i + 7] - c[8 * i]) / mhz) for i in range(1, 4)))
