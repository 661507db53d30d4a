"""GPU helper: per-C-ABI-call device time of one training rollout (CUDA events around every call, warm caches, in stream)."""
import sys, os, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dasa_b200 import lib, modules as M, ops, synth
from dasa_b200.config import FULL
from dasa_b200.rollout import DeviceEpisodes, NavPolicy

prec = sys.argv[1] if len(sys.argv) > 1 else "tf32"
T = int(sys.argv[2]) if len(sys.argv) > 2 else 3
SCHED = sys.argv[3] if len(sys.argv) > 3 else "batched"
ops.set_precision(prec)
import dasa_b200.functions as Fn
Fn.defer_weight_grads(True)
cfg = FULL
pol = NavPolicy(cfg, synth.policy_state(cfg, 0)).train()
pol.schedule = SCHED
ep = DeviceEpisodes(synth.Episodes(20, T, cfg, seed=1))
src = M.DropoutSource(seed=3)

def run():
    pol.zero_grad()
    with M.use_dropout_source(src):
        loss, _, _ = pol.teacher_rollout(ep, T, 0.4, tag_steps=False)
    pol.backward(loss)
    pol.optim_step(1e-4)

run(); run()
torch.cuda.synchronize()
events = []
orig = ops.call
def hooked(name, *a):
    key = name
    if name == "dasa_gemm":
        key = "gemm a%d b%d M=%d N=%d K=%d epi=%d" % (a[0], a[1], a[2], a[3], a[4], a[13])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); rc = orig(name, *a); e1.record()
    events.append((key, e0, e1))
    return rc
ops.call = hooked
import dasa_b200.functions as Fn
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record(); run(); t1.record()
torch.cuda.synchronize()
ops.call = orig
agg = collections.defaultdict(lambda: [0, 0.0])
for k, a, b in events:
    agg[k][0] += 1; agg[k][1] += a.elapsed_time(b)
tot = sum(v[1] for v in agg.values())
print("precision %s, T=%d: wall (device) %.2f ms, sum of bracketed calls %.2f ms, %d calls" % (prec, T, t0.elapsed_time(t1), tot, len(events)))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print("%-52s n=%5d total %8.3f ms (%4.1f%%) avg %7.1f us" % (k, v[0], v[1], 100 * v[1] / tot, 1e3 * v[1] / v[0]))
