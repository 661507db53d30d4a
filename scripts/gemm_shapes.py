"""Event-time every dasa_gemm launch of one training rollout (bench config) and aggregate by (layout, M, N, K, epilogue)."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dasa_b200 import lib, ops, synth, functions as Fn, modules as M
from dasa_b200.config import FULL
from dasa_b200.rollout import DeviceEpisodes, NavPolicy
lib.load(); ops.set_precision("tf32"); Fn.defer_weight_grads(True)
T = 35
pol = NavPolicy(FULL, synth.policy_state(FULL, 0), "cuda").train(); pol.flatten_parameters()
ep = DeviceEpisodes(synth.Episodes(20, T, FULL, seed=100), "cuda")
src = M.DropoutSource(seed=1, device_seed=True, device="cuda")
def run(hook=None):
    pol.zero_grad(); src.advance()
    with M.use_dropout_source(src):
        loss, _, _ = pol.teacher_rollout(ep, T, tag_steps=False)
    pol.backward(loss); torch.cuda.synchronize()
run(); run()
events = []
orig = lib.call
def hooked(name, *a):
    if name != "dasa_gemm":
        return orig(name, *a)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); rc = orig(name, *a); e1.record()
    events.append((e0, e1, (a[0], a[1], a[2], a[3], a[4], a[13])))
    return rc
ops.call = hooked
run()
ops.call = orig
agg = collections.defaultdict(lambda: [0, 0.0])
for e0, e1, k in events:
    agg[k][0] += 1; agg[k][1] += e0.elapsed_time(e1) * 1e3
tot = sum(v[1] for v in agg.values())
print("total GEMM time %.2f ms over %d launches" % (tot / 1e3, len(events)))
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
    ak, bk, m, nn, kk, epi = k
    fl = 2.0 * m * nn * kk * n
    print("ak=%d bk=%d M=%6d N=%5d K=%6d epi=%d  n=%3d  %8.1f us total  %7.1f us avg  %6.1f TF/s" % (ak, bk, m, nn, kk, epi, n, us, us / n, fl / us / 1e6))
