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


def timeit(fn, n=8):
    ts = []
    for i in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


# the fp16 attention kernel on the same packed problem (in-place dropout draws: no mask tensor)
qkvh = qkvp.half()
ds = ops.DropStream(None, 7, 0, 0.1)
for name, drop in (("stream dropout", ds), ("mask tensor", mask2), ("no dropout", None)):
    ms = timeit(lambda: ops.mha_fwd_h16(qkvh[:, :hd], qkvh[:, hd:2 * hd], qkvh[:, 2 * hd:], H, pk, pk, L2, L2, None, drop, 1 / 0.9))
    byt = 2 * ntok * hd * 4
    print("mha_fwd_h16 packed self-attention, %s: %.1f us, %.0f GB/s (q/k/v/out fp16)" % (name, ms * 1e3, byt / ms / 1e6))
# cross attention: packed language queries over the 36 dense views of 700 samples, and the reverse
vis = torch.randn(B, 36, 2 * hd, device="cuda").half()
qv = torch.randn(B, 36, hd, device="cuda").half()
ms = timeit(lambda: ops.mha_fwd_h16(qkvh[:, :hd], vis[..., :hd], vis[..., hd:], H, pk, None, L2, 36, None, ds, 1 / 0.9))
print("mha_fwd_h16 lang q over views: %.1f us, %.0f GB/s" % (ms * 1e3, 2 * (2 * ntok * hd + 2 * B * 36 * hd) / ms / 1e6))
ms = timeit(lambda: ops.mha_fwd_h16(qv, qkvh[:, hd:2 * hd], qkvh[:, 2 * hd:], H, None, pk, 36, L2, None, ds, 1 / 0.9))
print("mha_fwd_h16 view q over lang: %.1f us, %.0f GB/s" % (ms * 1e3, 2 * (2 * ntok * hd + 2 * B * 36 * hd) / ms / 1e6))
vqkv = torch.randn(B, 36, 3 * hd, device="cuda").half()
ms = timeit(lambda: ops.mha_fwd_h16(vqkv[..., :hd], vqkv[..., hd:2 * hd], vqkv[..., 2 * hd:], H, drop=ds, drop_scale=1 / 0.9))
print("mha_fwd_h16 view self-attention: %.1f us, %.0f GB/s" % (ms * 1e3, 2 * 4 * B * 36 * hd / ms / 1e6))
# residual LayerNorm of the frozen stack: fp32 x + mask tensor (old) vs fp16 x + in-place draws
R = ntok
x = torch.randn(R, hd, device="cuda")
res = torch.randn(R, hd, device="cuda")
gm, bt = torch.ones(hd, device="cuda"), torch.zeros(hd, device="cuda")
mk = (torch.rand(R, hd, device="cuda") >= 0.1).to(torch.uint8)
ms = timeit(lambda: ops.dropout_residual_layernorm(x, res, gm, bt, 1e-12, mk, 1 / 0.9, half_copy=True))
print("LN fp32 x + mask tensor, %d rows: %.1f us, %.0f GB/s" % (R, ms * 1e3, R * hd * (4 + 4 + 4 + 2 + 1) / ms / 1e6))
ms = timeit(lambda: ops.dropout_residual_layernorm_fwd(x, res, gm, bt, 1e-12, ds, 1 / 0.9, half_copy=True))
print("LN fp32 x + stream: %.1f us, %.0f GB/s" % (ms * 1e3, R * hd * (4 + 4 + 4 + 2) / ms / 1e6))
xh = x.half()
ms = timeit(lambda: ops.dropout_residual_layernorm_fwd(xh, res, gm, bt, 1e-12, ds, 1 / 0.9, half_copy=True))
print("LN fp16 x + stream: %.1f us, %.0f GB/s" % (ms * 1e3, R * hd * (2 + 4 + 4 + 2) / ms / 1e6))
