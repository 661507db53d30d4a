"""GPU helper: small-batch bi-LSTM (B = 20, L = 45, H = 1024) forward + backward, persistent kernels vs per-step kernels."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dasa_b200 import functions as Fn, lib, ops
B, L, In, H = 20, 45, 768, 1024
gen = torch.Generator().manual_seed(0)
names = ["w_ih", "w_hh", "b_ih", "b_hh"]
P = []
for sfx in range(2):
    P += [(torch.randn(4 * H, In, generator=gen) / math.sqrt(In)).cuda().requires_grad_(True),
          (torch.randn(4 * H, H, generator=gen) / math.sqrt(H)).cuda().requires_grad_(True),
          (torch.randn(4 * H, generator=gen) * 0.1).cuda().requires_grad_(True), (torch.randn(4 * H, generator=gen) * 0.1).cuda().requires_grad_(True)]
lengths = torch.tensor([45, 40, 39, 39, 37, 35, 34, 34, 34, 32, 30, 29, 28, 27, 25, 22, 10, 9, 9, 8], dtype=torch.int32).cuda()
x = torch.randn(B, L, In, generator=gen).cuda().requires_grad_(True)
gout = torch.randn(B, L, 2 * H, generator=gen).cuda()
ops.set_precision("tf32")
orig = ops.call
for mode in (0, 1):
    lib.load().dasa_debug_bilstm_persist(mode)
    ev = {}
    def hooked(name, *a):
        if not name.startswith("dasa_bilstm_seq"):
            return orig(name, *a)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); rc = orig(name, *a); e1.record()
        ev.setdefault(name, []).append((e0, e1))
        return rc
    for it in range(4):
        ops.call = hooked if it == 3 else orig
        for p_ in P:
            p_.grad = None
        out, h, c = Fn.BiLSTMFn.apply(x, lengths, *P)
        (out * gout).sum().backward()
    torch.cuda.synchronize()
    ops.call = orig
    print("persist=%d " % mode + "  ".join("%s %.1f us" % (k, sum(a.elapsed_time(b) for a, b in v) * 1e3) for k, v in ev.items()))
