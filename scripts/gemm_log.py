"""Log every dasa_gemm call (layout, shape, leading dimensions, pointer alignment) of one training rollout at the bench config."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dasa_b200 import lib, ops, synth, functions as Fn, modules as M
from dasa_b200.config import FULL
from dasa_b200.rollout import DeviceEpisodes, NavPolicy
lib.load(); ops.set_precision("tf32"); Fn.defer_weight_grads(True)
T = 3
pol = NavPolicy(FULL, synth.policy_state(FULL, 0), "cuda").train(); pol.flatten_parameters()
ep = DeviceEpisodes(synth.Episodes(20, T, FULL, seed=100), "cuda")
src = M.DropoutSource(seed=1, device_seed=True, device="cuda")
log = collections.Counter()
orig = lib.call
def hooked(name, *a):
    if name == "dasa_gemm":
        ak, bk, m, n, k = a[0], a[1], a[2], a[3], a[4]
        A, lda, B, ldb, C, ldc = a[6], a[7], a[8], a[9], a[11], a[12]
        al = "A%d B%d C%d" % ((A or 0) % 16, (B or 0) % 16, (C or 0) % 16)
        log[(ak, bk, m, n, k, lda % 4, ldb % 4, al, a[13], a[15])] += 1
    return orig(name, *a)
ops.call = hooked
with M.use_dropout_source(src):
    loss, _, _ = pol.teacher_rollout(ep, T, tag_steps=False)
pol.backward(loss)
torch.cuda.synchronize()
for k, v in sorted(log.items(), key=lambda kv: (kv[0][2], kv[0][3])):
    if k[2] <= 64:
        print("n=%3d ak=%d bk=%d M=%5d N=%5d K=%5d lda%%4=%d ldb%%4=%d %s epi=%d prec=%d" % ((v,) + k))
