"""GPU helper for ncu: the fp16 attention kernel on the packed language self-attention of the batched rollout."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dasa_b200 import ops
B, H, hd, L2 = 700, 12, 768, 80
torch.manual_seed(0)
lens = torch.randn(B).mul(11).add(29).round().clamp(8, L2).to(torch.int32)
off = torch.cumsum(lens, 0).to(torch.int32) - lens
ntok = int(lens.sum())
qkvh = torch.randn(ntok, 3 * hd, device="cuda").half()
pk = (off.cuda(), lens.cuda())
ds = ops.DropStream(None, 7, 0, 0.1)
for _ in range(3):
    o = ops.mha_fwd_h16(qkvh[:, :hd], qkvh[:, hd:2 * hd], qkvh[:, 2 * hd:], H, pk, pk, L2, L2, None, ds, 1 / 0.9)
torch.cuda.synchronize()
x = torch.randn(ntok, hd, device="cuda").half()
res = torch.randn(ntok, hd, device="cuda")
gm, bt = torch.ones(hd, device="cuda"), torch.zeros(hd, device="cuda")
for _ in range(3):
    ops.dropout_residual_layernorm_fwd(x, res, gm, bt, 1e-12, ds, 1 / 0.9, half_copy=True)
torch.cuda.synchronize()
