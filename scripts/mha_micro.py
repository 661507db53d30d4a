"""GPU helper: la-layer self-attention shape of the batched rollout (700 sequences x 12 heads x 45 tokens x 64)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dasa_b200 import ops
ops.set_precision("tf32")
B, L, H, hd = 700, 45, 12, 768
qkv = torch.randn(B, L, 3 * hd, device="cuda")
pad = (torch.arange(L, device="cuda")[None, :] >= torch.randint(8, L + 1, (B, 1), device="cuda")).to(torch.uint8)
mask = (torch.rand(B, H, L, L, device="cuda") >= 0.1).to(torch.uint8)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = []
for i in range(8):
    flush.zero_()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record(); ops.mha_fwd(qkv[..., :hd], qkv[..., hd:2 * hd], qkv[..., 2 * hd:], H, pad, mask, 1 / 0.9); e1.record()
    torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
ts.sort(); ms = ts[len(ts) // 2]
byt = 4 * B * L * hd * 4 + B * H * L * L
print("mha_fwd B=%d L=%d: %.1f us, %.0f GB/s" % (B, L, ms * 1e3, byt / ms / 1e6))
# packed (variable-length) self-attention as the batched rollout issues it: 700 sequences, lengths ~ clip(N(29, 11), 8, 80), L = 80
L2 = 80
lens = torch.randn(B).mul(11).add(29).round().clamp(8, L2).to(torch.int32)
off = torch.cumsum(lens, 0).to(torch.int32) - lens
ntok = int(lens.sum())
qkvp = torch.randn(ntok, 3 * hd, device="cuda")
pk = (off.cuda(), lens.cuda())
mask2 = (torch.rand(B, H, L2, L2, device="cuda") >= 0.1).to(torch.uint8)
ts = []
for i in range(8):
    flush.zero_()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record(); ops.mha_fwd_varlen(qkvp[:, :hd], qkvp[:, hd:2 * hd], qkvp[:, 2 * hd:], H, pk, pk, L2, L2, mask2, 1 / 0.9); e1.record()
    torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
ts.sort(); ms = ts[len(ts) // 2]
byt = 4 * ntok * hd * 4 + int((lens.long() ** 2).sum()) * H
print("mha_fwd_varlen %d tokens (700 seqs, max %d): %.1f us, %.0f GB/s" % (ntok, int(lens.max()), ms * 1e3, byt / ms / 1e6))
