"""GPU tier (B200): every C-ABI kernel against the CPU oracle (oracle/restated.py) or a float64 torch restatement on
identical seeded inputs. Bars: fp32 paths 1e-4 relative (north_star); index outputs bit-exact."""
import math

import pytest
import torch

from dasa_b200 import synth
from dasa_b200.config import FULL, SMALL
from oracle import restated as R

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from dasa_b200 import functions as Fn
    from dasa_b200 import modules as M
    from dasa_b200 import ops

DEV = "cuda"


def rel_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def assert_close(a, b, tol=1e-4, what=""):
    e = rel_err(a, b)
    assert e <= tol, "%s: relative error %.3e > %.1e" % (what, e, tol)


def g(seed):
    return torch.Generator().manual_seed(seed)


# ------------------------------------------------------------------------------------------------------ GEMM
@pytest.mark.parametrize("M,N,K", [(20, 4096, 3264), (20, 2176, 1024), (1, 5, 1024), (720, 768, 2176), (1600, 3072, 768),
                                   (37, 53, 29), (128, 128, 8), (257, 130, 1000), (1000, 2048, 2048)])
@pytest.mark.parametrize("layout", [(1, 1), (1, 0), (0, 0), (0, 1)])
def test_gemm_fp32(M, N, K, layout):
    ak, bk = layout
    if (M * N * K > 3e9 / 4) and layout != (1, 1):
        pytest.skip("large shapes only in the Linear layout")
    gen = g(M + N + K)
    A = torch.randn((M, K) if ak else (K, M), generator=gen)
    Bm = torch.randn((N, K) if bk else (K, N), generator=gen)
    C0 = torch.randn(M, N, generator=gen)
    bias = torch.randn(N, generator=gen)
    ref = (A if ak else A.t()).double() @ (Bm.t() if bk else Bm).double()
    Ad, Bd = A.to(DEV), Bm.to(DEV)
    C = C0.to(DEV).clone()
    ops.gemm(Ad, Ad.stride(0), ak, Bd, Bd.stride(0), bk, C, N, M, N, K, alpha=0.5, beta=2.0, epilogue=ops.EPI_BIAS,
             bias=bias.to(DEV), precision=ops.PREC_FP32)
    assert_close(C, 0.5 * ref + 2.0 * C0.double() + bias.double(), 2e-5, "gemm")


@pytest.mark.parametrize("epi", ["tanh", "gelu", "relu", "bias_tanh"])
def test_gemm_epilogues(epi):
    gen = g(5)
    M, N, K = 70, 200, 300
    A, W, b = torch.randn(M, K, generator=gen), torch.randn(N, K, generator=gen) / math.sqrt(K), torch.randn(N, generator=gen)
    z = A.double() @ W.double().t()
    code, ref = {"tanh": (ops.EPI_TANH, torch.tanh(z)), "gelu": (ops.EPI_BIAS_GELU, R.gelu_erf(z + b.double())),
                 "relu": (ops.EPI_BIAS_RELU, torch.relu(z + b.double())),
                 "bias_tanh": (ops.EPI_BIAS_TANH, torch.tanh(z + b.double()))}[epi]
    y = ops.linear_fwd(A.to(DEV), W.to(DEV), None if epi == "tanh" else b.to(DEV), code, precision=ops.PREC_FP32)
    assert_close(y, ref, 2e-5, epi)


def test_linear_backward_helpers():
    gen = g(6)
    M, N, K = 45, 96, 130
    x, w, dy = torch.randn(M, K, generator=gen), torch.randn(N, K, generator=gen), torch.randn(M, N, generator=gen)
    dx = ops.linear_bwd_input(dy.to(DEV), w.to(DEV), precision=ops.PREC_FP32)
    assert_close(dx, dy.double() @ w.double(), 2e-5, "dx")
    dw0 = torch.randn(N, K, generator=gen)
    dw = dw0.to(DEV).clone()
    ops.linear_bwd_weight(dy.to(DEV), x.to(DEV), dw, True, precision=ops.PREC_FP32)
    assert_close(dw, dw0.double() + dy.double().t() @ x.double(), 2e-5, "dw")
    out = torch.zeros(N, device=DEV)
    ops.colsum(dy.to(DEV), out, False)
    assert_close(out, dy.double().sum(0), 2e-5, "colsum")


# ------------------------------------------------------------------------------------------------ AdaIN family
@pytest.mark.parametrize("cfg,B", [(SMALL, 3), (FULL, 2)])
def test_adain_gate_fused_and_epilogue(cfg, B):
    st = synth.adain_state(cfg, 0)
    ep = synth.Episodes(B, 1, cfg, seed=2)
    C = cfg.rgb_size
    f, d = ep.f_t[0], ep.d_t[0]
    want = R.adain_channel_gate(st, f[..., :C], d[..., :C])
    mod = M.DGAdaChannel(C).to(DEV)
    mod.load_state_dict(st)
    fd, dd = f.to(DEV), d.to(DEV)
    got = mod(fd[..., :C], dd[..., :C])                        # strided slices, read in place
    assert_close(got, want, 1e-4, "DGAdaChannel")
    full = mod.gate_features(fd, dd)
    assert_close(full[..., :C], want, 1e-4, "gate_features rgb")
    assert torch.equal(full[..., C:].cpu(), f[..., C:])          # angle part carried over bit-exactly
    # epilogue-only form given the pre-activation
    gpre = torch.nn.functional.linear(d[..., :C], st["a_fc.weight"], st["a_fc.bias"]).to(DEV)
    out = torch.empty_like(gpre)
    ops.gate_modulate(gpre, fd[..., :C], out)
    assert_close(out, want, 1e-5, "gate_modulate")
    # with an injected drop_env mask
    keep = torch.rand(B, cfg.views, C, generator=g(3)) >= 0.4
    full2 = mod.gate_features(fd, dd, keep.to(torch.uint8).to(DEV), 1 / 0.6)
    assert_close(full2[..., :C], want * keep / 0.6, 1e-4, "gate + mask")


@pytest.mark.parametrize("cfg,B", [(SMALL, 4), (FULL, 2)])
def test_adain_stat_mean_default(cfg, B):
    ep = synth.Episodes(B, 1, cfg, seed=4, stress=True)
    C = cfg.rgb_size
    f, d = ep.f_t[0][..., :C], ep.d_t[0][..., :C]
    fd, dd = ep.f_t[0].to(DEV)[..., :C], ep.d_t[0].to(DEV)[..., :C]
    assert_close(ops.view_stats(dd), R.view_stats(d), 1e-5, "view_stats")
    for kind, cls, fn in (("stat", M.DGAdaStatChannel, R.adain_stat_channel), ("mean", M.DGAdaMeanChannel, R.adain_mean_channel)):
        st = synth.adain_state(cfg, 0, kind)
        mod = cls(C).to(DEV)
        mod.load_state_dict(st)
        assert_close(mod(fd, dd), fn(st, f, d), 1e-4, kind)
    assert_close(M.adaptive_instance_normalization(fd, dd), R.adain_default(f, d), 1e-4, "adain default")


@pytest.mark.parametrize("cfg,B", [(SMALL, 4), (FULL, 3)])
@pytest.mark.parametrize("kind", ["stat", "mean"])
def test_adain_stat_mean_backward(cfg, B, kind):
    """a2 backward: gradients of a_fc / b_fc (and of f when it requires grad) against autograd through the oracle."""
    cls, fn = {"stat": (M.DGAdaStatChannel, R.adain_stat_channel), "mean": (M.DGAdaMeanChannel, R.adain_mean_channel)}[kind]
    ep = synth.Episodes(B, 1, cfg, seed=6)
    C = cfg.rgb_size
    f, d = ep.f_t[0][..., :C].clone().requires_grad_(True), ep.d_t[0][..., :C]
    st = {k: v.clone().requires_grad_(True) for k, v in synth.adain_state(cfg, 0, kind).items()}
    w = torch.randn(B, cfg.views, C, generator=g(9))
    (fn(st, f, d) * w).sum().backward()
    mod = cls(C).to(DEV)
    mod.load_state_dict({k: v.detach() for k, v in st.items()})
    f_all = ep.f_t[0].to(DEV)                                   # strided RGB slice of the [B, V, C+A] buffer, read in place
    fd = f_all[..., :C].detach().requires_grad_(True)
    dd = ep.d_t[0].to(DEV)[..., :C]
    out = mod(fd, dd)
    (out * w.to(DEV)).sum().backward()
    for k, p in mod.state_dict(keep_vars=True).items():
        assert_close(p.grad, st[k].grad, 2e-4, "%s d%s" % (kind, k))
    assert_close(fd.grad, f.grad, 1e-4, kind + " df")
    # env-data form: f carries no gradient -> df is not produced
    mod.zero_grad()
    out = mod(f_all[..., :C], dd)
    (out * w.to(DEV)).sum().backward()
    assert_close(mod.a_fc.weight.grad, st["a_fc.weight"].grad, 2e-4, kind + " da_fc (no df)")
    with pytest.raises(NotImplementedError):
        mod(f_all[..., :C], dd.clone().requires_grad_(True))


# ------------------------------------------------------------------------------------------- decoder attention
@pytest.mark.parametrize("cfg,B", [(SMALL, 5), (FULL, 3), (FULL, 40)])
def test_shift_attention_fwd_bwd(cfg, B):
    st = {k: v.clone().requires_grad_(True) for k, v in synth.decoder_state(cfg, 0).items()}
    ep = synth.Episodes(B, 1, cfg, seed=8)
    gen = g(9)
    h = torch.tanh(torch.randn(B, cfg.hidden, generator=gen)).requires_grad_(True)
    ctxt = ep.f_t[0].clone().requires_grad_(True)
    wc, p = R.shift_soft_dot_attention(st, "feat_att_layer.", h, ctxt, cfg.headings)
    dwc = torch.randn(wc.shape, generator=gen)
    wc.backward(dwc)
    mod = M.ShiftSoftDotAttention(cfg.hidden, cfg.feat, cfg.shift_kernel).to(DEV)
    mod.load_state_dict({k.split(".", 1)[1]: v.detach() for k, v in st.items() if k.startswith("feat_att_layer.")})
    hd = h.detach().to(DEV).requires_grad_(True)
    cd = ctxt.detach().to(DEV).requires_grad_(True)
    wc2, p2 = mod(hd, cd, output_tilde=False)
    assert_close(p2, p, 1e-4, "view softmax")
    assert_close(wc2, wc, 1e-4, "weighted context")
    wc2.backward(dwc.to(DEV))
    assert_close(hd.grad, h.grad, 2e-4, "dh")
    assert_close(cd.grad, ctxt.grad, 2e-4, "dctx")
    assert_close(mod.linear_in.weight.grad, st["feat_att_layer.linear_in.weight"].grad, 2e-4, "dW_in")
    assert_close(mod.linear_shift.weight.grad, st["feat_att_layer.linear_shift.weight"].grad, 2e-4, "dW_shift")
    assert_close(mod.linear_shift.bias.grad, st["feat_att_layer.linear_shift.bias"].grad, 2e-4, "db_shift")


def _row_attention_f64(ctx, t, k, headings, kl):
    """model.py:318-353 restated in float64: softmax over the rows, circular k-tap shift along the heading axis."""
    ctx, t = ctx.double(), t.double()
    p = torch.softmax(torch.einsum("brd,bd->br", ctx, t), 1)
    w = p
    if k > 0:
        kap = torch.softmax(kl.double(), 1)
        B, rows = p.shape
        pe = p.view(B, rows // headings, headings)
        w = torch.zeros_like(pe)
        for j in range(k):
            w = w + kap[:, j, None, None] * torch.roll(pe, -(j - k // 2), 2)
        w = w.reshape(B, rows)
    return torch.einsum("br,brd->bd", w, ctx), p, w


@pytest.mark.parametrize("B,rows,D,k,headings", [(300, 36, 2176, 5, 12), (1000, 36, 2176, 5, 12), (150, 36, 4224, 5, 12),
                                                 (200, 36, 520, 3, 12), (257, 50, 2048, 0, 1), (64, 36, 3200, 5, 12)])
def test_row_attention_fwd_large_batch(B, rows, D, k, headings):
    """Batches >= 64 without a mask take the persistent pipelined cluster kernel (row_attention_pipe.cu)."""
    gen = g(B + D)
    ctx = torch.relu(torch.randn(B, rows, D, generator=gen)) * 0.5
    t = torch.randn(B, D, generator=gen) * 0.05
    kl = torch.randn(B, max(k, 1), generator=gen)
    wc, attn, q, kappa = ops.row_attention_fwd(ctx.to(DEV), t.to(DEV), None, k, headings, kl.to(DEV) if k else None, want_q=True)
    wc_ref, p_ref, w_ref = _row_attention_f64(ctx, t, k, headings, kl)
    assert_close(attn, p_ref, 1e-5, "softmax")
    assert_close(q, w_ref, 1e-5, "shifted weights")
    assert_close(wc, wc_ref, 1e-5, "weighted context")
    if k:
        assert_close(kappa, torch.softmax(kl.double(), 1), 1e-5, "kappa")
    # strided context (the RGB slice of a wider feature row), output into a strided buffer
    if D > 128:
        wide = torch.zeros(B, rows, D + 128)
        wide[..., :D] = ctx
        wd = wide.to(DEV)
        wc2, attn2, _, _ = ops.row_attention_fwd(wd[..., :D], t.to(DEV), None, k, headings, kl.to(DEV) if k else None)
        assert torch.equal(wc2, wc) and torch.equal(attn2, attn)


@pytest.mark.parametrize("cfg,B", [(SMALL, 5), (FULL, 3), (FULL, 24)])
def test_soft_dot_attention_fwd_bwd(cfg, B):
    st = {k: v.clone().requires_grad_(True) for k, v in synth.decoder_state(cfg, 0).items()}
    gen = g(10)
    seq, mask, lens = synth.instructions(B, cfg, seed=3)
    L = mask.shape[1]
    h = torch.tanh(torch.randn(B, cfg.hidden, generator=gen)).requires_grad_(True)
    ctxt = (torch.randn(B, L, cfg.ctx_dim, generator=gen) * 0.3 * (~mask).unsqueeze(-1)).requires_grad_(True)
    ht, alpha = R.soft_dot_attention(st, "attention_layer.", h, ctxt, mask)
    dht = torch.randn(ht.shape, generator=gen)
    ht.backward(dht)
    mod = M.SoftDotAttention(cfg.hidden, cfg.ctx_dim).to(DEV)
    mod.load_state_dict({k.split(".", 1)[1]: v.detach() for k, v in st.items() if k.startswith("attention_layer.")})
    hd = h.detach().to(DEV).requires_grad_(True)
    cd = ctxt.detach().to(DEV).requires_grad_(True)
    ht2, alpha2 = mod(hd, cd, mask.to(DEV))
    assert_close(alpha2, alpha, 1e-4, "alpha")
    assert float(alpha2[mask.to(DEV)].abs().max()) == 0.0
    assert_close(ht2, ht, 1e-4, "h_tilde")
    ht2.backward(dht.to(DEV))
    assert_close(hd.grad, h.grad, 2e-4, "dh")
    assert_close(cd.grad, ctxt.grad, 2e-4, "dctx")
    assert_close(mod.linear_in.weight.grad, st["attention_layer.linear_in.weight"].grad, 2e-4, "dW_in")
    assert_close(mod.linear_out.weight.grad, st["attention_layer.linear_out.weight"].grad, 2e-4, "dW_out")


@pytest.mark.parametrize("cfg,B", [(SMALL, 5), (FULL, 6)])
def test_candidate_logits_fwd_bwd(cfg, B):
    st = {k: v.clone().requires_grad_(True) for k, v in synth.decoder_state(cfg, 0).items()}
    ep = synth.Episodes(B, 1, cfg, seed=12)
    gen = g(13)
    h = torch.tanh(torch.randn(B, cfg.hidden, generator=gen)).requires_grad_(True)
    cand = ep.cand_feat[0].clone().requires_grad_(True)
    leng = ep.cand_leng[0]
    lg = R.candidate_logits(st, "candidate_att_layer.", h, cand)
    lg = lg.masked_fill(R.length2mask(leng, lg.shape[1]), -float("inf"))
    dl = torch.randn(lg.shape, generator=gen) * torch.isfinite(lg)
    torch.where(torch.isfinite(lg), lg, torch.zeros_like(lg)).backward(dl)
    mod = M.SoftDotAttention(cfg.hidden, cfg.feat).to(DEV)
    mod.load_state_dict({k.split(".", 1)[1]: v.detach() for k, v in st.items() if k.startswith("candidate_att_layer.")})
    hd = h.detach().to(DEV).requires_grad_(True)
    cd = cand.detach().to(DEV).requires_grad_(True)
    _, lg2 = mod(hd, cd, output_prob=False, cand_leng=leng.to(DEV), rgb_channels=cfg.rgb_size)
    fin = torch.isfinite(lg)
    assert torch.equal(torch.isfinite(lg2).cpu(), fin)
    assert_close(lg2.cpu()[fin], lg[fin], 1e-4, "logits")
    lg2.backward(dl.to(DEV))
    assert_close(hd.grad, h.grad, 2e-4, "dh")
    assert_close(cd.grad[..., :cfg.rgb_size], cand.grad[..., :cfg.rgb_size], 2e-4, "dcand rgb")
    assert_close(mod.linear_in.weight.grad, st["candidate_att_layer.linear_in.weight"].grad, 2e-4, "dW_in")


# ------------------------------------------------------------------------------------------------------- LSTM
@pytest.mark.parametrize("cfg,B", [(SMALL, 5), (FULL, 20)])
def test_lstm_cell_fwd_bwd(cfg, B):
    st = {k: v.clone().requires_grad_(True) for k, v in synth.decoder_state(cfg, 0).items() if k.startswith("lstm.")}
    gen = g(14)
    x = torch.randn(B, cfg.action_emb + cfg.feat, generator=gen).requires_grad_(True)
    h = torch.tanh(torch.randn(B, cfg.hidden, generator=gen)).requires_grad_(True)
    c = torch.randn(B, cfg.hidden, generator=gen).requires_grad_(True)
    h1, c1 = R.lstm_cell(st["lstm.weight_ih"], st["lstm.weight_hh"], st["lstm.bias_ih"], st["lstm.bias_hh"], x, h, c)
    dh1, dc1 = torch.randn(h1.shape, generator=gen), torch.randn(c1.shape, generator=gen)
    (h1 * dh1).sum().add((c1 * dc1).sum()).backward()
    P = {k: v.detach().to(DEV).requires_grad_(True) for k, v in st.items()}
    xd, hd, cd = (t.detach().to(DEV).requires_grad_(True) for t in (x, h, c))
    xh = torch.cat((xd, hd), 1)                                   # the decoder's one concatenation [emb ; attn_feat ; prev_h1]
    h2, c2 = Fn.LSTMCellFn.apply(xh, cd, P["lstm.weight_ih"], P["lstm.weight_hh"], P["lstm.bias_ih"], P["lstm.bias_hh"])
    assert_close(h2, h1, 1e-4, "h1")
    assert_close(c2, c1, 1e-4, "c1")
    torch.autograd.backward([h2, c2], [dh1.to(DEV), dc1.to(DEV)])
    for a, b, n in ((xd, x, "dx"), (hd, h, "dh"), (cd, c, "dc")):
        assert_close(a.grad, b.grad, 2e-4, n)
    for k in st:
        assert_close(P[k].grad, st[k].grad, 2e-4, k)


def test_masked_ce_and_argmax():
    gen = g(15)
    B, Nc = 37, 11
    lg = torch.randn(B, Nc, generator=gen) * 3
    leng = torch.randint(2, Nc + 1, (B,), generator=gen)
    lg = lg.masked_fill(R.length2mask(leng, Nc), -float("inf"))
    lg[3, :2] = 1.25                                              # exact tie -> first index wins
    tgt = torch.stack([torch.randint(0, int(n), (1,), generator=gen)[0] for n in leng])
    tgt[::5] = -100
    lgr = lg.clone().requires_grad_(True)
    want = torch.nn.functional.cross_entropy(lgr, tgt, ignore_index=-100, reduction="sum")
    want.backward()
    lgd = lg.to(DEV).requires_grad_(True)
    loss, act = Fn.MaskedCEFn.apply(lgd, tgt.to(DEV), -100)
    assert_close(loss, want, 1e-5, "ce")
    loss.backward()
    assert_close(lgd.grad, lgr.grad, 1e-5, "dlogit")
    assert torch.equal(act.cpu(), lg.argmax(1))
    _, a2, lp, ent = ops.masked_ce(lg.to(DEV), None, -100, 0.0, None, want_grad=False, want_stats=True)
    dist = torch.distributions.Categorical(logits=lg)
    assert_close(ent, dist.entropy(), 1e-4, "entropy")
    assert_close(lp, torch.log_softmax(lg, 1).gather(1, lg.argmax(1, keepdim=True)).squeeze(1), 1e-4, "logprob")


def test_rmsprop_and_clip_match_torch():
    gen = g(16)
    p0, g0 = torch.randn(1000, 33, generator=gen), torch.randn(1000, 33, generator=gen) * 3
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.RMSprop([ref], lr=1e-4)
    pd, sq = p0.to(DEV).clone(), torch.zeros_like(p0, device=DEV)
    ssq, coef = torch.zeros(1, device=DEV), torch.ones(1, device=DEV)
    for it in range(3):
        ref.grad = g0.clone() * (it + 1)
        torch.nn.utils.clip_grad_norm_([ref], 40.0)
        opt.step()
        gd = (g0 * (it + 1)).to(DEV)
        ssq.zero_()
        ops.sumsq(gd, ssq)
        ops.clip_coef(ssq, 40.0, coef)
        ops.rmsprop_step(pd, gd, sq, 1e-4, 0.99, 1e-8, 0.0, coef)
    assert_close(pd, ref.detach(), 1e-5, "rmsprop")


def test_dropout_mask_rate_and_determinism():
    m1 = ops.dropout_mask((1 << 20,), 0.4, 7, 0)
    m2 = ops.dropout_mask((1 << 20,), 0.4, 7, 0)
    assert torch.equal(m1, m2)
    assert abs(float(m1.float().mean()) - 0.6) < 5e-3
    assert not torch.equal(m1, ops.dropout_mask((1 << 20,), 0.4, 8, 0))


@pytest.mark.parametrize("B,L,In,H", [(3, 9, 96, 128), (20, 24, 768, 128), (20, 80, 768, 1024), (7, 33, 64, 64), (45, 11, 96, 128)])
def test_bilstm_fused_fwd_bwd(B, L, In, H):
    """Fused per-step bi-LSTM kernels (packed-sequence semantics, ragged lengths) against the oracle + autograd."""
    gen = g(17 + B)
    names = ["weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0"]
    sd = {}
    for sfx in ("", "_reverse"):
        sd["lstm.weight_ih_l0" + sfx] = torch.randn(4 * H, In, generator=gen) / math.sqrt(In)
        sd["lstm.weight_hh_l0" + sfx] = torch.randn(4 * H, H, generator=gen) / math.sqrt(H)
        sd["lstm.bias_ih_l0" + sfx] = torch.randn(4 * H, generator=gen) * 0.1
        sd["lstm.bias_hh_l0" + sfx] = torch.randn(4 * H, generator=gen) * 0.1
    sd = {k: v.requires_grad_(True) for k, v in sd.items()}
    lengths = torch.randint(1, L + 1, (B,), generator=gen).sort(descending=True).values
    lengths[0] = L
    x = torch.randn(B, L, In, generator=gen).requires_grad_(True)
    out, ((hf, cf), (hb, cb)) = R.bilstm(sd, x, lengths, H)
    wts = [torch.randn(t.shape, generator=gen) for t in (out, hf, cf, hb, cb)]
    sum((t * w).sum() for t, w in zip((out, hf, cf, hb, cb), wts)).backward()
    P = {k: v.detach().to(DEV).requires_grad_(True) for k, v in sd.items()}
    xd = x.detach().to(DEV).requires_grad_(True)
    l32 = lengths.to(DEV).to(torch.int32)
    o2, hfin, cfin = Fn.BiLSTMFn.apply(xd, l32, *[P["lstm." + n] for n in names], *[P["lstm." + n + "_reverse"] for n in names])
    assert_close(o2, out, 1e-4, "seq out")
    assert_close(hfin[0], hf, 1e-4, "h fwd"); assert_close(hfin[1], hb, 1e-4, "h bwd")
    assert_close(cfin[0], cf, 1e-4, "c fwd"); assert_close(cfin[1], cb, 1e-4, "c bwd")
    pad = torch.arange(L)[None, :] >= lengths[:, None]
    if bool(pad.any()):
        assert float(o2[pad.to(DEV)].abs().max()) == 0.0
    dhf = torch.stack((wts[1], wts[3])).to(DEV)
    dcf = torch.stack((wts[2], wts[4])).to(DEV)
    torch.autograd.backward([o2, hfin, cfin], [wts[0].to(DEV), dhf, dcf])
    assert_close(xd.grad, x.grad, 3e-4, "dx")
    for k in sd:
        assert_close(P[k].grad, sd[k].grad, 3e-4, k)


@pytest.mark.parametrize("B,L,In,H", [(20, 45, 768, 1024), (5, 12, 64, 128), (300, 14, 96, 128), (700, 9, 64, 256), (37, 21, 64, 1024)])
def test_bilstm_tf32_tensor_core_path(B, L, In, H):
    """TF32 recurrence (tf32 precision mode): mma.sync per-step kernels up to B = 20, above that the grouped CTA-pair tcgen05 GEMM
    of both directions + one pointwise launch per step (split-K partials summed by the next step's pointwise kernel).
    Stated bound 1e-2 relative against the fp32 oracle."""
    gen = g(23 + B)
    names = ["weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0"]
    sd = {}
    for sfx in ("", "_reverse"):
        sd["lstm.weight_ih_l0" + sfx] = torch.randn(4 * H, In, generator=gen) / math.sqrt(In)
        sd["lstm.weight_hh_l0" + sfx] = torch.randn(4 * H, H, generator=gen) / math.sqrt(H)
        sd["lstm.bias_ih_l0" + sfx] = torch.randn(4 * H, generator=gen) * 0.1
        sd["lstm.bias_hh_l0" + sfx] = torch.randn(4 * H, generator=gen) * 0.1
    sd = {k: v.requires_grad_(True) for k, v in sd.items()}
    lengths = torch.randint(1, L + 1, (B,), generator=gen).sort(descending=True).values
    lengths[0] = L
    x = torch.randn(B, L, In, generator=gen).requires_grad_(True)
    out, ((hf, cf), (hb, cb)) = R.bilstm(sd, x, lengths, H)
    wts = [torch.randn(t.shape, generator=gen) for t in (out, hf, cf, hb, cb)]
    sum((t * w).sum() for t, w in zip((out, hf, cf, hb, cb), wts)).backward()
    P = {k: v.detach().to(DEV).requires_grad_(True) for k, v in sd.items()}
    xd = x.detach().to(DEV).requires_grad_(True)
    l32 = lengths.to(DEV).to(torch.int32)
    ops.set_precision("tf32")
    try:
        o2, hfin, cfin = Fn.BiLSTMFn.apply(xd, l32, *[P["lstm." + n] for n in names], *[P["lstm." + n + "_reverse"] for n in names])
        dhf = torch.stack((wts[1], wts[3])).to(DEV)
        dcf = torch.stack((wts[2], wts[4])).to(DEV)
        torch.autograd.backward([o2, hfin, cfin], [wts[0].to(DEV), dhf, dcf])
    finally:
        ops.set_precision("fp32")
    errs = {"out": rel_err(o2, out), "h_f": rel_err(hfin[0], hf), "h_b": rel_err(hfin[1], hb), "dx": rel_err(xd.grad, x.grad)}
    for k in sd:
        errs[k] = rel_err(P[k].grad, sd[k].grad)
    print("tf32 bilstm errors:", {k: "%.2e" % v for k, v in errs.items()})
    assert max(errs.values()) <= 1e-2, errs


@pytest.mark.parametrize("B,L,H", [(20, 45, 1024), (7, 13, 1024), (1, 5, 512), (20, 80, 256)])
def test_bilstm_persistent_kernels_match_per_step_kernels(B, L, H):
    """Small-batch bi-LSTM in the tensor-core mode: the persistent cooperative kernels (whole time loop in one launch, fp16 weights
    resident in shared memory, fp16 state / scaled gate-gradient exchange) against the per-step TF32 kernels they replace
    (dasa_debug_bilstm_persist): outputs, final states and every gradient, ragged lengths incl. length 1. Both carry 11-bit operands."""
    from dasa_b200 import lib
    In = 64
    gen = g(77 + B + L)
    names = ["weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0"]
    base = {}
    for sfx in ("", "_reverse"):
        base["weight_ih_l0" + sfx] = torch.randn(4 * H, In, generator=gen) / math.sqrt(In)
        base["weight_hh_l0" + sfx] = torch.randn(4 * H, H, generator=gen) / math.sqrt(H)
        base["bias_ih_l0" + sfx] = torch.randn(4 * H, generator=gen) * 0.1
        base["bias_hh_l0" + sfx] = torch.randn(4 * H, generator=gen) * 0.1
    lengths = torch.randint(1, L + 1, (B,), generator=gen).sort(descending=True).values
    lengths[0] = L
    lengths[-1] = 1
    x = torch.randn(B, L, In, generator=gen)
    wts = [torch.randn(B, L, 2 * H, generator=gen), torch.randn(2, B, H, generator=gen), torch.randn(2, B, H, generator=gen)]
    l32 = lengths.to(DEV).to(torch.int32)
    res = []
    ops.set_precision("tf32")
    try:
        for mode in (0, 1):
            lib.load().dasa_debug_bilstm_persist(mode)
            P = {k: v.clone().to(DEV).requires_grad_(True) for k, v in base.items()}
            xd = x.clone().to(DEV).requires_grad_(True)
            o2, hfin, cfin = Fn.BiLSTMFn.apply(xd, l32, *[P[n] for n in names], *[P[n + "_reverse"] for n in names])
            torch.autograd.backward([o2, hfin, cfin], [w.to(DEV) for w in wts])
            torch.cuda.synchronize()
            res.append([o2, hfin, cfin, xd.grad] + [P[k].grad for k in sorted(P)])
    finally:
        lib.load().dasa_debug_bilstm_persist(1)
        ops.set_precision("fp32")
    for i, (a, b) in enumerate(zip(res[1], res[0])):
        assert rel_err(a, b) <= 2e-3, (i, rel_err(a, b))
    valid = torch.arange(L, device=DEV).view(1, L) < l32.view(B, 1)
    assert float(res[1][0][~valid].abs().max()) == 0.0


@pytest.mark.parametrize("prec,tol", [("fp32", 2e-5), ("tf32", 3e-3)])
@pytest.mark.parametrize("B,Lq,Lk", [(3, 45, 45), (2, 80, 36), (4, 36, 80), (2, 23, 23)])
def test_mha_forward_both_precisions(prec, tol, B, Lq, Lk):
    """BertSelfAttention / BertOutAttention core (vilmodel.py:203-236): padding mask (-10000), dropout keep mask, P.V."""
    gen = g(31 + Lq + Lk)
    heads, dh = 12, 64
    Hd = heads * dh
    q = torch.randn(B, Lq, Hd, generator=gen)
    kv = torch.randn(B, Lk, 2 * Hd, generator=gen)          # fused K|V buffer: strided views
    k, v = kv[..., :Hd], kv[..., Hd:]
    lens = torch.randint(max(1, Lk // 2), Lk + 1, (B,), generator=gen)
    pad = torch.arange(Lk)[None, :] >= lens[:, None]
    keep = torch.rand(B, heads, Lq, Lk, generator=gen) >= 0.1
    qh = q.view(B, Lq, heads, dh).permute(0, 2, 1, 3).double()
    kh = k.reshape(B, Lk, heads, dh).permute(0, 2, 1, 3).double()
    vh = v.reshape(B, Lk, heads, dh).permute(0, 2, 1, 3).double()
    s = qh @ kh.transpose(-1, -2) / math.sqrt(dh) + (pad.double() * -10000.0)[:, None, None, :]
    p = torch.softmax(s, -1)
    want = ((p * keep / 0.9) @ vh).permute(0, 2, 1, 3).reshape(B, Lq, Hd)
    kvd = kv.to(DEV)
    ops.set_precision(prec)
    try:
        out, probs = ops.mha_fwd(q.to(DEV), kvd[..., :Hd], kvd[..., Hd:], heads, pad.to(DEV), keep.to(torch.uint8).to(DEV), 1 / 0.9,
                                 save_probs=True)
    finally:
        ops.set_precision("fp32")
    assert_close(probs, p, tol, "probs")
    assert_close(out, want, tol, "attention output")


# ------------------------------------------------------------------------------- sampled feedback + A2C kernels
@pytest.mark.parametrize("B,Nc", [(20, 14), (513, 9), (3, 40)])
def test_policy_sample_fwd_bwd(B, Nc):
    gen = g(B * Nc)
    logit = torch.randn(B, Nc, generator=gen) * 3
    leng = torch.randint(2, Nc + 1, (B,), generator=gen)
    logit = logit.masked_fill(R.length2mask(leng, Nc), -float("inf")).requires_grad_(True)
    act = (torch.rand(B, generator=gen) * leng.float()).long().clamp(max=leng - 1)
    c = torch.distributions.Categorical(torch.softmax(logit, 1))
    lp, ent = c.log_prob(act), c.entropy()
    wl, we = torch.randn(B, generator=gen), torch.randn(B, generator=gen)
    (lp * wl + ent * we).sum().backward()
    ld = logit.detach().to(DEV).requires_grad_(True)
    a2, lp2, ent2 = Fn.PolicySampleFn.apply(ld, None, act.to(DEV))
    assert torch.equal(a2.cpu(), act)
    assert_close(lp2, lp, 1e-5, "log_prob")
    assert_close(ent2, ent, 1e-5, "entropy")
    (lp2 * wl.to(DEV) + ent2 * we.to(DEV)).sum().backward()
    assert_close(ld.grad, logit.grad, 1e-5, "dlogit")
    # argmax mode is torch.max; sampling mode is the inverse CDF of the supplied uniforms
    a3, _, _, probs = ops.policy_sample_fwd(ld.detach())
    assert torch.equal(a3.cpu(), logit.detach().argmax(1))
    u = torch.rand(B, generator=gen)
    a4, lp4, _, _ = ops.policy_sample_fwd(ld.detach(), u=u.to(DEV))
    cdf = probs.double().cpu().cumsum(1)
    want = (cdf <= u.double()[:, None]).sum(1).clamp(max=Nc - 1)
    live = torch.isfinite(logit.detach())
    assert bool(live.gather(1, a4.cpu()[:, None]).all())
    near = ((cdf - u.double()[:, None]).abs() < 1e-5).any(1)          # fp32 vs fp64 prefix sums may flip exact ties
    assert torch.equal(a4.cpu()[~near], want[~near])
    assert_close(lp4, torch.log_softmax(logit.detach().double(), 1).gather(1, a4.cpu()[:, None]).squeeze(1), 1e-5, "lp(sample)")


def test_policy_sample_distribution():
    """Sampling frequencies follow softmax(logit) (chi-square-like bound on 200k draws)."""
    logit = torch.tensor([[0.5, -1.0, 2.0, -float("inf"), 0.0]]).repeat(200000, 1).to(DEV)
    torch.manual_seed(1)
    a, _, _, p = ops.policy_sample_fwd(logit, u=torch.rand(200000, device=DEV))
    freq = torch.bincount(a, minlength=5).double() / 200000
    assert float(freq[3]) == 0.0
    assert float((freq - p[0].double()).abs().max()) < 5e-3


@pytest.mark.parametrize("T,B,norm", [(5, 20, "total"), (35, 512, "total"), (4, 7, "batch"), (3, 1500, "none")])
def test_a2c_loss_matches_oracle(T, B, norm):
    gen = g(T * B)
    logp = (-torch.rand(T, B, generator=gen) * 3).requires_grad_(True)
    ent = (torch.rand(T, B, generator=gen) * 2).requires_grad_(True)
    val = torch.randn(T, B, generator=gen).requires_grad_(True)
    last = torch.randn(B, generator=gen)
    rew = torch.randint(-2, 3, (T, B), generator=gen).float()
    ended_at = torch.randint(1, T + 2, (B,), generator=gen)
    mask = (torch.arange(T)[:, None] < ended_at[None, :]).float()
    ended = ended_at <= T
    loss, total = R.a2c_epilogue(list(logp), list(ent), list(val), last, list(rew), list(mask), ended, 0.9, 0.01, norm)
    loss.backward()
    lp2, en2, v2 = (x.detach().to(DEV).requires_grad_(True) for x in (logp, ent, val))
    loss2, total2 = Fn.A2CLossFn.apply(lp2, en2, v2, last.to(DEV), rew.to(DEV), mask.to(DEV), ended.to(torch.uint8).to(DEV),
                                       0.9, 0.01, norm)
    assert float(total2) == total
    assert_close(loss2, loss, 1e-5, "a2c loss")
    loss2.backward()
    assert_close(lp2.grad, logp.grad, 1e-5, "dlogp")
    assert_close(en2.grad, ent.grad, 1e-5, "dent")
    assert_close(v2.grad, val.grad, 1e-5, "dvalue")


def test_nav_reward_bit_exact():
    gen = g(99)
    B = 257
    leng = torch.randint(2, 14, (B,), generator=gen).int()
    a = (torch.rand(B, generator=gen) * leng.float()).long().clamp(max=leng.long() - 1)
    a[::7] = -100
    dist, last = torch.rand(B, generator=gen) * 6, torch.rand(B, generator=gen) * 6
    ended = torch.rand(B, generator=gen) < 0.3
    r, m, e2 = R.nav_reward(a, leng, -100, dist, last, ended)
    ed = ended.to(torch.uint8).to(DEV)
    rd, md = torch.empty(B, device=DEV), torch.empty(B, device=DEV)
    ops.nav_reward(a.to(DEV), leng.to(DEV), -100, dist.to(DEV), last.to(DEV), ed, rd, md)
    assert torch.equal(rd.cpu(), r) and torch.equal(md.cpu(), m) and torch.equal(ed.cpu().bool(), e2)


# ------------------------------------------------------------------- finetune config: attention / LN / GELU backward
@pytest.mark.parametrize("B,Lq,Lk,heads,dh,pad,drop", [(3, 7, 7, 4, 16, True, True), (2, 45, 36, 12, 64, False, True),
                                                        (4, 36, 80, 12, 64, True, False)])
def test_mha_fwd_bwd_matches_torch(B, Lq, Lk, heads, dh, pad, drop):
    gen = g(B * Lq + Lk)
    Hd = heads * dh
    q, k, v = (torch.randn(B, n, Hd, generator=gen, requires_grad=True) for n in (Lq, Lk, Lk))
    key_pad = None
    add = 0.0
    if pad:
        lens = torch.randint(1, Lk + 1, (B,), generator=gen)
        key_pad = torch.arange(Lk)[None, :] >= lens[:, None]
        add = (key_pad.float() * -10000.0)[:, None, None, :]
    keep = (torch.rand(B, heads, Lq, Lk, generator=gen) >= 0.1) if drop else None
    s = torch.matmul(q.view(B, Lq, heads, dh).permute(0, 2, 1, 3), k.view(B, Lk, heads, dh).permute(0, 2, 3, 1)) / math.sqrt(dh) + add
    pr = torch.softmax(s, -1)
    if drop:
        pr = pr * keep.float() / 0.9
    o = torch.matmul(pr, v.view(B, Lk, heads, dh).permute(0, 2, 1, 3)).permute(0, 2, 1, 3).reshape(B, Lq, Hd)
    do = torch.randn(B, Lq, Hd, generator=gen)
    o.backward(do)
    qd, kd, vd = (x.detach().to(DEV).requires_grad_(True) for x in (q, k, v))
    o2 = Fn.MHAFn.apply(qd, kd, vd, heads, key_pad.to(DEV) if pad else None,
                        ops.as_keep_mask(keep.to(DEV)) if drop else None, 1 / 0.9 if drop else 1.0)
    assert_close(o2, o, 1e-4, "mha out")
    o2.backward(do.to(DEV))
    assert_close(qd.grad, q.grad, 2e-4, "dq")
    assert_close(kd.grad, k.grad, 2e-4, "dk")
    assert_close(vd.grad, v.grad, 2e-4, "dv")


@pytest.mark.parametrize("R,Hd,resid,pre,post", [(37, 768, True, True, False), (20, 768, False, False, True), (5, 64, True, False, False)])
def test_dropout_residual_layernorm_bwd(R, Hd, resid, pre, post):
    gen = g(R + Hd)
    x = torch.randn(R, Hd, generator=gen, requires_grad=True)
    r = torch.randn(R, Hd, generator=gen, requires_grad=True) if resid else None
    gamma = (1 + 0.1 * torch.randn(Hd, generator=gen)).requires_grad_(True)
    beta = (0.1 * torch.randn(Hd, generator=gen)).requires_grad_(True)
    km = (torch.rand(R, Hd, generator=gen) >= 0.1) if pre else None
    pm = (torch.rand(R, Hd, generator=gen) >= 0.1) if post else None
    z = x * (km.float() / 0.9) if pre else x
    z = z + r if resid else z
    y = torch.nn.functional.layer_norm(z, (Hd,), gamma, beta, 1e-12)
    y = y * (pm.float() / 0.9) if post else y
    dy = torch.randn(R, Hd, generator=gen)
    y.backward(dy)
    xd = x.detach().to(DEV).requires_grad_(True)
    rd = r.detach().to(DEV).requires_grad_(True) if resid else None
    gd, bd = gamma.detach().to(DEV).requires_grad_(True), beta.detach().to(DEV).requires_grad_(True)
    y2 = Fn.DropResLNFn.apply(xd, rd, gd, bd, 1e-12, ops.as_keep_mask(km.to(DEV)) if pre else None, 1 / 0.9 if pre else 1.0,
                              ops.as_keep_mask(pm.to(DEV)) if post else None, 1 / 0.9 if post else 1.0)
    assert_close(y2, y, 1e-4, "LN out")
    y2.backward(dy.to(DEV))
    assert_close(xd.grad, x.grad, 2e-4, "dx")
    if resid:
        assert_close(rd.grad, r.grad, 2e-4, "dresid")
    assert_close(gd.grad, gamma.grad, 2e-4, "dgamma")
    assert_close(bd.grad, beta.grad, 2e-4, "dbeta")


def test_linear_gelu_fwd_bwd():
    gen = g(17)
    x = torch.randn(6, 11, 48, generator=gen, requires_grad=True)
    w = (torch.randn(96, 48, generator=gen) * 0.2).requires_grad_(True)
    b = torch.randn(96, generator=gen).requires_grad_(True)
    y = R.gelu_erf(torch.nn.functional.linear(x, w, b))
    dy = torch.randn(y.shape, generator=gen)
    y.backward(dy)
    xd, wd, bd = (t.detach().to(DEV).requires_grad_(True) for t in (x, w, b))
    y2 = Fn.linear(xd, wd, bd, "gelu")
    assert_close(y2, y, 1e-4, "gelu(linear)")
    y2.backward(dy.to(DEV))
    Fn.flush_weight_grads()
    assert_close(xd.grad, x.grad, 2e-4, "dx")
    assert_close(wd.grad, w.grad, 2e-4, "dW")
    assert_close(bd.grad, b.grad, 2e-4, "db")


@pytest.mark.parametrize("prec,tol", [("fp32", 2e-5), ("tf32", 3e-3)])
def test_mha_varlen_packed_matches_padded(prec, tol):
    """dasa_mha_fwd_varlen: packed queries and/or keys give the same rows as the padded call with a key padding mask."""
    gen = g(77)
    B, L, V, heads, dh = 6, 45, 36, 12, 64
    Hd = heads * dh
    lens = torch.randint(5, L + 1, (B,), generator=gen)
    lens[0] = L
    pad = torch.arange(L)[None, :] >= lens[:, None]
    x = torch.randn(B, L, 3 * Hd, generator=gen)
    vis = torch.randn(B, V, 3 * Hd, generator=gen)
    keep = torch.rand(B, heads, L, L, generator=gen) >= 0.1
    rows = torch.cat([torch.arange(n) + b * L for b, n in enumerate(lens.tolist())])
    off = torch.cat([torch.zeros(1, dtype=torch.long), lens.cumsum(0)[:-1]]).int().to(DEV)
    len32 = lens.int().to(DEV)
    xd, vd = x.to(DEV), vis.to(DEV)
    xp = xd.view(B * L, -1)[rows.to(DEV)].contiguous()
    km = keep.to(torch.uint8).to(DEV)
    ops.set_precision(prec)
    try:
        ref = ops.mha_fwd(xd[..., :Hd], xd[..., Hd:2 * Hd], xd[..., 2 * Hd:], heads, pad.to(DEV), km, 1 / 0.9)
        got = ops.mha_fwd_varlen(xp[:, :Hd], xp[:, Hd:2 * Hd], xp[:, 2 * Hd:], heads, (off, len32), (off, len32), L, L, km, 1 / 0.9)
        assert_close(got, ref.view(B * L, Hd)[rows.to(DEV)], tol, "self-attention, packed q and k")
        # language queries (packed) over the dense views; dense view queries over the packed language keys
        ref = ops.mha_fwd(xd[..., :Hd], vd[..., Hd:2 * Hd], vd[..., 2 * Hd:], heads)
        got = ops.mha_fwd_varlen(xp[:, :Hd], vd[..., Hd:2 * Hd], vd[..., 2 * Hd:], heads, (off, len32), None, L, V)
        assert_close(got, ref.view(B * L, Hd)[rows.to(DEV)], tol, "packed q, dense k")
        ref = ops.mha_fwd(vd[..., :Hd], xd[..., Hd:2 * Hd], xd[..., 2 * Hd:], heads, pad.to(DEV))
        got = ops.mha_fwd_varlen(vd[..., :Hd], xp[:, Hd:2 * Hd], xp[:, 2 * Hd:], heads, None, (off, len32), V, L)
        assert_close(got, ref, tol, "dense q, packed k")
    finally:
        ops.set_precision("fp32")
    rev = ops.reverse_tokens_packed(xp[:, :Hd].contiguous(), off, len32, L)
    assert torch.equal(rev, ops.reverse_tokens(xd[..., :Hd].contiguous(), len32))


def test_gemm_tf32_multicast_variant():
    """Cluster-of-2 TMA-multicast tcgen05 GEMM (experiment switch): same numbers as the default kernel, all its epilogues."""
    from dasa_b200 import lib
    gen = g(3)
    M, N, K = 5000, 1024, 800                       # 40 x 8 tiles > 2 x 148: eligible
    A, W, b = torch.randn(M, K, generator=gen), torch.randn(N, K, generator=gen), torch.randn(N, generator=gen)
    Ad, Wd, bd = A.to(DEV), W.to(DEV), b.to(DEV)
    ref = A.double() @ W.double().t() + b.double()
    outs = []
    for on in (0, 1):
        lib.load().dasa_debug_gemm_multicast(on)
        lib.load().dasa_debug_gemm_pair(0)
        try:
            C = torch.empty(M, N, device=DEV)
            ops.gemm(Ad, K, 1, Wd, K, 1, C, N, M, N, K, epilogue=ops.EPI_BIAS, bias=bd, precision=ops.PREC_TF32)
            Cg = torch.empty(M, N, device=DEV)
            ops.gemm(Ad, K, 1, Wd, K, 1, Cg, N, M, N, K, epilogue=ops.EPI_BIAS_GELU, bias=bd, precision=ops.PREC_TF32)
        finally:
            lib.load().dasa_debug_gemm_multicast(0)
            lib.load().dasa_debug_gemm_pair(1)
        outs.append((C, Cg))
        assert_close(C, ref, 3e-3, "tf32 gemm (multicast=%d)" % on)
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


@pytest.mark.parametrize("shape", [(256, 256, 32), (777, 515, 96), (3000, 768, 768), (1000, 132, 40), (20300, 768, 160)])
def test_gemm_tf32_pair_kernel(shape):
    """Persistent CTA-pair tcgen05 GEMM (cta_group::2, 256x256 tiles, gemm_tc2.cu) forced on for every shape: all fused
    epilogues, ragged M / N / K edges, beta accumulation, strided output rows, dropout keep mask; against an fp64 reference
    (TF32 products: 3e-3 of the output scale) and against the single-CTA tcgen05 kernel."""
    from dasa_b200 import lib
    M, N, K = shape
    gen = g(11)
    A, W, b = torch.randn(M, K, generator=gen), torch.randn(N, K, generator=gen) * 0.1, torch.randn(N, generator=gen)
    F = torch.rand(M, N + 8, generator=gen)
    keep = (torch.rand(M, N, generator=gen) > 0.3).to(torch.uint8)
    C0 = torch.randn(M, N + 4, generator=gen)
    Ad, Wd, bd, Fd, kd = A.to(DEV), W.to(DEV), b.to(DEV), F.to(DEV), keep.to(DEV)
    acc = A.double() @ W.double().t()
    sig = torch.sigmoid(acc + b.double())
    cases = [
        ("none", dict(epilogue=ops.EPI_NONE), acc),
        ("bias", dict(epilogue=ops.EPI_BIAS, bias=bd), acc + b.double()),
        ("bias+tanh", dict(epilogue=ops.EPI_BIAS_TANH, bias=bd), torch.tanh(acc + b.double())),
        ("bias+gelu", dict(epilogue=ops.EPI_BIAS_GELU, bias=bd), torch.nn.functional.gelu(acc + b.double())),
        ("bias+relu", dict(epilogue=ops.EPI_BIAS_RELU, bias=bd), torch.relu(acc + b.double())),
        ("tanh", dict(epilogue=ops.EPI_TANH), torch.tanh(acc)),
        ("bias+drop", dict(epilogue=ops.EPI_BIAS, bias=bd, drop_mask=kd, drop_scale=1 / 0.7), (acc + b.double()) * keep.double() / 0.7),
    ]
    L = lib.load()
    try:
        for name, kw, ref in cases:
            outs = []
            for mode in (0, 2):
                L.dasa_debug_gemm_pair(mode)
                C = torch.full((M, N + 4), 7.0, device=DEV)                      # strided rows; the pad columns must stay untouched
                ops.gemm(Ad, K, 1, Wd, K, 1, C, N + 4, M, N, K, precision=ops.PREC_TF32, **kw)
                assert float((C[:, N:] - 7.0).abs().max()) == 0.0, name
                outs.append(C[:, :N])
            # tanh has slope 1 where TF32 rounding of |acc| ~ 10 lands: the absolute product error passes straight through
            assert_close(outs[1], ref, 1e-2 if "tanh" in name else 3e-3, "pair kernel, %s" % name)
            assert_close(outs[1], outs[0], 1e-5, "pair vs single-CTA kernel, %s" % name)
        # sigmoid gate: out = sigmoid(acc + b) * f (f strided), the gate itself saved on the side
        L.dasa_debug_gemm_pair(2)
        C, G = torch.empty(M, N, device=DEV), torch.empty(M, N, device=DEV)
        ops.gemm(Ad, K, 1, Wd, K, 1, C, N, M, N, K, precision=ops.PREC_TF32, epilogue=ops.EPI_GATE, bias=bd, gate_src=Fd, ld_gate=N + 8,
                 gate_out=G, ld_gate_out=N)
        assert_close(G, sig, 3e-3, "pair kernel, gate")
        assert_close(C, sig * F[:, :N].double(), 3e-3, "pair kernel, gated output")
        # alpha / beta accumulation into an existing strided C
        C = C0.to(DEV).clone()
        ops.gemm(Ad, K, 1, Wd, K, 1, C, N + 4, M, N, K, alpha=0.5, beta=1.0, precision=ops.PREC_TF32)
        assert_close(C[:, :N], 0.5 * acc + C0[:, :N].double(), 3e-3, "pair kernel, alpha/beta")
        assert torch.equal(C[:, N:].cpu(), C0[:, N:])
    finally:
        L.dasa_debug_gemm_pair(1)


def test_gemm_pair_plan_routes_big_token_major_shapes():
    """The rollout's many-tile GEMMs run on the pair kernel by default and agree with the single-CTA kernel's TF32 result."""
    from dasa_b200 import lib
    L = lib.load()
    gen = g(12)
    M, N, K = 9472, 2304, 768                       # 37 x 9 = 333 pair tiles: 4.5 waves of the 74 TPCs
    A, W, b = torch.randn(M, K, generator=gen).to(DEV), torch.randn(N, K, generator=gen).to(DEV), torch.randn(N, generator=gen).to(DEV)
    outs = []
    try:
        for mode in (0, 1):
            L.dasa_debug_gemm_pair(mode)
            C = torch.empty(M, N, device=DEV)
            ops.gemm(A, K, 1, W, K, 1, C, N, M, N, K, epilogue=ops.EPI_BIAS, bias=b, precision=ops.PREC_TF32)
            outs.append(C)
    finally:
        L.dasa_debug_gemm_pair(1)
    assert_close(outs[1], outs[0], 1e-5, "default routing vs single-CTA kernel")
    assert_close(outs[1], A.double().cpu() @ W.double().cpu().t() + b.double().cpu(), 3e-3, "default routing vs fp64")


@pytest.mark.parametrize("shape", [(20, 4096, 2240), (20, 1024, 4096), (20, 2176, 1024), (1, 64, 128), (32, 48, 64), (7, 2240, 4096),
                                   (20, 1000, 96), (13, 16, 32)])
def test_gemm_skinny_kernel(shape):
    """Weight-streaming mma.sync TF32 kernel for M <= 32 rows (the decoder's per-action projections): K slices reduced through
    shared memory and the cluster's DSMEM. Against fp64 (TF32 products: 3e-3 of the output scale) for every epilogue it serves,
    beta accumulation and strided rows, forced on for every shape; bit-reproducible."""
    from dasa_b200 import lib
    M, N, K = shape
    gen = g(21)
    A, W, b = torch.randn(M, K + 4, generator=gen), torch.randn(N, K, generator=gen) * 0.1, torch.randn(N, generator=gen)
    C0 = torch.randn(M, N + 4, generator=gen)
    Ad, Wd, bd = A.to(DEV), W.to(DEV), b.to(DEV)
    acc = A[:, :K].double() @ W.double().t()
    L = lib.load()
    cases = [("none", dict(epilogue=ops.EPI_NONE), acc), ("bias", dict(epilogue=ops.EPI_BIAS, bias=bd), acc + b.double()),
             ("bias+tanh", dict(epilogue=ops.EPI_BIAS_TANH, bias=bd), torch.tanh(acc + b.double())),
             ("tanh", dict(epilogue=ops.EPI_TANH), torch.tanh(acc)),
             ("bias+relu", dict(epilogue=ops.EPI_BIAS_RELU, bias=bd), torch.relu(acc + b.double()))]
    try:
        for name, kw, ref in cases:
            outs = []
            for mode in (2, 0):
                L.dasa_debug_gemm_skinny(mode)
                C = torch.full((M, N + 4), 7.0, device=DEV)
                ops.gemm(Ad, K + 4, 1, Wd, K, 1, C, N + 4, M, N, K, precision=ops.PREC_TF32, **kw)
                assert float((C[:, N:] - 7.0).abs().max()) == 0.0, name
                outs.append(C[:, :N])
            assert_close(outs[0], ref, 1e-2 if "tanh" in name else 3e-3, "skinny kernel, %s" % name)
            assert_close(outs[0], outs[1], 3e-3, "skinny vs tile kernel, %s" % name)
        L.dasa_debug_gemm_skinny(2)
        C = C0.to(DEV).clone()
        ops.gemm(Ad, K + 4, 1, Wd, K, 1, C, N + 4, M, N, K, alpha=0.5, beta=1.0, precision=ops.PREC_TF32)
        assert_close(C[:, :N], 0.5 * acc + C0[:, :N].double(), 3e-3, "skinny kernel, alpha/beta")
        assert torch.equal(C[:, N:].cpu(), C0[:, N:])
        C2 = C0.to(DEV).clone()
        ops.gemm(Ad, K + 4, 1, Wd, K, 1, C2, N + 4, M, N, K, alpha=0.5, beta=1.0, precision=ops.PREC_TF32)
        assert torch.equal(C, C2), "skinny kernel must be deterministic"
    finally:
        L.dasa_debug_gemm_skinny(1)


@pytest.mark.parametrize("shape", [(256, 256, 32), (512, 260, 96), (777, 516, 1000), (4096, 768, 3000), (2048, 2240, 700)])
@pytest.mark.parametrize("layout", [(1, 0), (0, 0), (0, 1)])
def test_gemm_tf32_pair_kernel_mn_major(shape, layout):
    """MN-major operands on the CTA-pair tcgen05 kernel (A stored [K][M] and / or B stored [K][N]; TMA boxes of 32 columns x 32
    k rows, MN-major shared-memory descriptors): the backward GEMMs dX = dY.W (1,0) and dW += dY^T.X (0,0) without transposed
    copies. Against an fp64 reference (TF32 products) and against the K-major pair kernel fed explicit transposes."""
    from dasa_b200 import lib
    M, N, K = shape
    ak, bk = layout
    gen = g(13 + M + ak)
    pad4 = lambda n: (n + 3) // 4 * 4
    A = torch.randn((M, K) if ak else (K, pad4(M) + 4), generator=gen)   # padded rows: strided operands (TMA: stride % 16 B == 0)
    Bm = torch.randn((N, K) if bk else (K, pad4(N) + 8), generator=gen) * 0.1
    C0 = torch.randn(M, N + 4, generator=gen)
    A_log = A if ak else A[:, :M].t()
    B_log = Bm if bk else Bm[:, :N].t()
    acc = A_log.double() @ B_log.double().t()
    Ad, Bd = A.to(DEV), Bm.to(DEV)
    L = lib.load()
    try:
        L.dasa_debug_gemm_pair(2)
        assert L.dasa_gemm_layout_on_tensor_cores(ak, bk, M, N, K) == 1
        C = torch.full((M, N + 4), 7.0, device=DEV)
        ops.gemm(Ad, Ad.stride(0), ak, Bd, Bd.stride(0), bk, C, N + 4, M, N, K, precision=ops.PREC_TF32)
        assert float((C[:, N:] - 7.0).abs().max()) == 0.0
        assert_close(C[:, :N], acc, 3e-3, "MN-major pair kernel")
        # the same products through the K-major path (explicit transposes): identical tensor-core arithmetic
        At = A_log.contiguous().to(DEV)
        Bt = B_log.contiguous().to(DEV)
        Ck = torch.empty(M, N, device=DEV)
        ops.gemm(At, K, 1, Bt, K, 1, Ck, N, M, N, K, precision=ops.PREC_TF32)
        assert_close(C[:, :N], Ck, 1e-5, "MN-major vs K-major pair kernel")
        # accumulate into an existing gradient (beta = 1), as linear_bwd_weight does
        C2 = C0.to(DEV).clone()
        ops.gemm(Ad, Ad.stride(0), ak, Bd, Bd.stride(0), bk, C2, N + 4, M, N, K, beta=1.0, precision=ops.PREC_TF32)
        assert_close(C2[:, :N], acc + C0[:, :N].double(), 3e-3, "MN-major pair kernel, beta")
        assert torch.equal(C2[:, N:].cpu(), C0[:, N:])
    finally:
        L.dasa_debug_gemm_pair(1)


def test_linear_backward_helpers_tf32_mn_major():
    """ops.linear_bwd_input / linear_bwd_weight at rollout-sized shapes: the MN-major route equals the transposed-copy route."""
    gen = g(21)
    M, N, K = 6000, 4096, 768
    x, w, dy = torch.randn(M, K, generator=gen), torch.randn(N, K, generator=gen) * 0.05, torch.randn(M, N, generator=gen)
    xd, wd, dyd = x.to(DEV), w.to(DEV), dy.to(DEV)
    ops.set_precision("tf32")
    try:
        outs = []
        for flag in (True, False):
            ops.use_mn_major = flag
            dx = ops.linear_bwd_input(dyd, wd)
            dw = torch.zeros(N, K, device=DEV)
            ops.linear_bwd_weight(dyd, xd, dw, True)
            outs.append((dx, dw))
        assert_close(outs[0][0], dy.double() @ w.double(), 3e-3, "dx (MN-major B)")
        assert_close(outs[0][1], dy.double().t() @ x.double(), 3e-3, "dw (MN-major A, B)")
        assert_close(outs[0][0], outs[1][0], 1e-5, "dx: MN-major vs transposed copy")
        # the MN-major route splits this few-tile reduction over K (fp32 partial sums folded in a fixed order)
        assert_close(outs[0][1], outs[1][1], 1e-4, "dw: MN-major vs transposed copy")
    finally:
        ops.use_mn_major = True
        ops.set_precision("fp32")


@pytest.mark.parametrize("rows,N,K", [(3000, 512, 768), (19810, 4096, 1024), (1000, 304, 200), (4100, 1024, 768)])
def test_gemm_f16_mn_major_weight_gradient(rows, N, K):
    """dasa_gemm_f16_mn: dW[N, K] = alpha * dY^T X (+ beta dW) with fp16 operands stored row-per-reduction-index (both MN-major),
    tcgen05 kind::f16, split-K with a deterministic fold, against fp64 on the same fp16 values: only fp32 accumulation error."""
    from dasa_b200 import ops
    g = torch.Generator().manual_seed(rows + N)
    dy = (torch.randn(rows, N, generator=g) * 0.3).half().to(DEV)
    x = (torch.randn(rows, K, generator=g) * 0.5).half().to(DEV)
    dw0 = torch.randn(N, K, generator=g).to(DEV)
    ref = 0.25 * (dy.double().t() @ x.double())
    dw = dw0.clone()
    ops.linear_bwd_weight_f16(dy, x, dw, alpha=0.25, accumulate=True)
    assert rel_err(dw.double(), ref + dw0.double()) <= 1e-4, rel_err(dw.double(), ref + dw0.double())
    dw2 = torch.full_like(dw0, float("nan"))
    ops.linear_bwd_weight_f16(dy, x, dw2, alpha=0.25, accumulate=False)
    assert rel_err(dw2.double(), ref) <= 1e-4
    # strided operands (column slices of wider buffers), deterministic
    wide = torch.zeros(rows, N + 64, dtype=torch.float16, device=DEV)
    wide[:, 8:8 + N] = dy
    dw3 = torch.zeros_like(dw0)
    ops.linear_bwd_weight_f16(wide[:, 8:8 + N], x, dw3, alpha=0.25, accumulate=False)
    assert torch.equal(dw3, dw2)


def test_adain_gate_on_fp16_operands_matches_tf32_path():
    """AdaINGateFn in the tensor-core mode: the gate GEMM over an fp16 copy of the depth features (dasa_gemm_f16, GATE epilogue) and
    its weight gradient over the scaled fp16 gradient copy (dasa_gate_backward_h + dasa_gemm_f16_mn) against the TF32 operand path
    and against fp64: output, saved gate, dW, db. Depth features O(1..10) stay inside fp16's range, so both carry 11-bit operands."""
    from dasa_b200 import functions as Fn
    from dasa_b200 import ops
    R, C, F_all = 36 * 128, 2048, 2176
    g = torch.Generator().manual_seed(11)
    f = torch.randn(R, F_all, generator=g).abs().to(DEV)
    d = (torch.randn(R, F_all, generator=g).abs() * 2.0).to(DEV)
    w0 = (torch.randn(C, C, generator=g) * C ** -0.5)
    b0 = torch.randn(C, generator=g) * 0.1
    gout = (torch.randn(R, F_all, generator=g) * 0.01).to(DEV)
    mask = (torch.rand(R, C, generator=g) > 0.3).to(torch.uint8).to(DEV)
    res = []
    ops.set_precision("tf32")
    try:
        for half in (False, True):
            ops.half_gate = half
            w = w0.clone().to(DEV).requires_grad_(True)
            b = b0.clone().to(DEV).requires_grad_(True)
            out = Fn.AdaINGateFn.apply(f, d, w, b, mask, 1.0 / 0.7, C)
            (out * gout).sum().backward()
            torch.cuda.synchronize()
            res.append((out, w.grad.clone(), b.grad.clone()))
    finally:
        ops.half_gate = True
        ops.set_precision("fp32")
    # fp64 reference
    wd, bd = w0.double().to(DEV).requires_grad_(True), b0.double().to(DEV).requires_grad_(True)
    gate = torch.sigmoid(d[:, :C].double() @ wd.t() + bd)
    ref = torch.cat([gate * f[:, :C].double() * mask.double() / 0.7, f[:, C:].double()], 1)
    (ref * gout.double()).sum().backward()
    for name, (out, dw, db) in zip(("tf32", "fp16"), res):
        assert rel_err(out.double(), ref) <= 3e-3, (name, rel_err(out.double(), ref))
        assert rel_err(dw.double(), wd.grad) <= 3e-3, (name, rel_err(dw.double(), wd.grad))
        assert rel_err(db.double(), bd.grad) <= 3e-3, (name, rel_err(db.double(), bd.grad))
    assert rel_err(res[1][0], res[0][0]) <= 2e-3 and rel_err(res[1][1], res[0][1]) <= 2e-3


@pytest.mark.parametrize("M,N,ld", [(19810, 4096, 4096), (700, 2048, 2112), (333, 40, 40), (64, 1000, 1000)])
def test_colsum_fp16_scaled(M, N, ld):
    """dasa_colsum_h: out[n] (+)= scale * sum_m x16[m, n] (vectorised 8-column path and the scalar fallback) vs fp64; run twice for
    bit-reproducibility."""
    from dasa_b200 import ops
    g = torch.Generator().manual_seed(M + N)
    wide = (torch.randn(M, ld, generator=g) * 3.0).half().to(DEV)
    x = wide[:, :N]
    out0 = torch.randn(N, generator=g).to(DEV)
    ref = out0.double() + x.double().sum(0) / 256.0
    a = ops.colsum_h(x, 1.0 / 256.0, out0.clone(), True)
    b = ops.colsum_h(x, 1.0 / 256.0, out0.clone(), True)
    assert torch.equal(a, b)
    assert rel_err(a.double(), ref) <= 1e-5, rel_err(a.double(), ref)
    c = ops.colsum_h(x, 1.0 / 256.0, torch.full((N,), float("nan"), device=DEV), False)
    assert rel_err(c.double(), ref - out0.double()) <= 1e-5


# ------------------------------------------------------------------------------ padding-free bi-LSTM (bilstm_packed.cu)
@pytest.mark.parametrize("R,L,In,H,seed", [(70, 12, 64, 64, 0), (300, 21, 96, 128, 1), (45, 9, 64, 32, 2)])
def test_packed_bilstm_matches_padded_path(R, L, In, H, seed):
    """PackedBiLSTMFn (length-ranked, position-block token order, per-direction row counts in the grouped tcgen05 GEMM) against
    BiLSTMFn on the padded reversed input in exact fp32: sequence outputs (zero rows past each length), final states in original
    order, and every gradient incl. the packed input's. TF32 recurrence over up to 21 steps: 5e-3 normwise."""
    from dasa_b200 import functions as Fn
    from dasa_b200 import modules as M
    from dasa_b200 import ops
    g = torch.Generator().manual_seed(seed)
    lens = torch.randint(1, L + 1, (R,), generator=g).tolist()
    lens[3] = L
    lens[5] = 1
    pack = M.PackInfo(lens, L, 1, DEV)
    plan = pack.bilstm_plan()
    assert plan.N == sum(lens) and list(plan.n_rows_list)[0] == R
    x = (torch.randn(pack.ntok, In, generator=g) * 0.5).to(DEV)

    def weights():
        gg = torch.Generator().manual_seed(seed + 100)
        return [(torch.randn(*s, generator=gg) * sc).to(DEV).requires_grad_(True)
                for s, sc in (((4 * H, In), In ** -0.5), ((4 * H, H), H ** -0.5), ((4 * H,), 0.1), ((4 * H,), 0.1)) * 2]
    gout = torch.randn(R, L, 2 * H, generator=g).to(DEV)
    ghf = torch.randn(2, R, H, generator=g).to(DEV)
    gcf = torch.randn(2, R, H, generator=g).to(DEV)
    # reference: padded path, exact fp32
    w_ref = weights()
    x_ref = x.clone().requires_grad_(True)
    rev = ReversePackedFn.apply(x_ref, pack, L)
    out_r, h_r, c_r = Fn.BiLSTMFn.apply(rev, pack.len, *w_ref)
    ((out_r * gout).sum() + (h_r * ghf).sum() + (c_r * gcf).sum()).backward()
    # packed path, TF32
    w_pk = weights()
    x_pk = x.clone().requires_grad_(True)
    ops.set_precision("tf32")
    try:
        out_p, h_p, c_p = Fn.PackedBiLSTMFn.apply(x_pk, plan, *w_pk)
        ((out_p * gout).sum() + (h_p * ghf).sum() + (c_p * gcf).sum()).backward()
        torch.cuda.synchronize()
    finally:
        ops.set_precision("fp32")
    valid = torch.arange(L, device=DEV).view(1, L) < pack.len.view(R, 1)
    assert float(out_p[~valid].abs().max()) == 0.0, "rows past a sequence's length must be exactly zero"
    for name, a, b in (("out", out_p, out_r), ("h_fin", h_p, h_r), ("c_fin", c_p, c_r), ("dx", x_pk.grad, x_ref.grad)):
        assert rel_err(a, b) <= 5e-3, "%s: %.3e" % (name, rel_err(a, b))
    for i, nm in enumerate(("w_ih_f", "w_hh_f", "b_ih_f", "b_hh_f", "w_ih_r", "w_hh_r", "b_ih_r", "b_hh_r")):
        assert rel_err(w_pk[i].grad, w_ref[i].grad) <= 5e-3, "grad %s: %.3e" % (nm, rel_err(w_pk[i].grad, w_ref[i].grad))


@pytest.mark.parametrize("R,L,In,H,seed", [(70, 12, 64, 64, 0), (600, 23, 96, 128, 1), (300, 9, 64, 256, 2)])
def test_packed_bilstm_fused_cell_matches_two_launch_form(R, L, In, H, seed):
    """Forward recurrence with the LSTM cell in the GEMM epilogue (fp16 state rows x interleaved fp16 W_hh, tcgen05 kind::f16, one
    launch per step) against the two-launch form (TF32 grouped GEMM + pointwise kernel): outputs, final states, the saved gate
    activations through every gradient. Both carry 11-bit operands, so they agree far inside the TF32 bound; rows past a length stay
    exactly zero. R = 600 crosses the 256-row tile boundary in both directions (rows joining / leaving inside a tile)."""
    from dasa_b200 import functions as Fn
    from dasa_b200 import modules as M
    from dasa_b200 import ops
    g = torch.Generator().manual_seed(seed)
    lens = torch.randint(1, L + 1, (R,), generator=g).tolist()
    lens[3] = L
    lens[5] = 1
    pack = M.PackInfo(lens, L, 1, DEV)
    plan = pack.bilstm_plan()
    x = (torch.randn(pack.ntok, In, generator=g) * 0.5).to(DEV)

    def weights():
        gg = torch.Generator().manual_seed(seed + 100)
        return [(torch.randn(*s, generator=gg) * sc).to(DEV).requires_grad_(True)
                for s, sc in (((4 * H, In), In ** -0.5), ((4 * H, H), H ** -0.5), ((4 * H,), 0.1), ((4 * H,), 0.1)) * 2]
    gout = torch.randn(R, L, 2 * H, generator=g).to(DEV)
    ghf = torch.randn(2, R, H, generator=g).to(DEV)
    gcf = torch.randn(2, R, H, generator=g).to(DEV)
    res = []
    ops.set_precision("tf32")
    try:
        for fused in (False, True):
            ops.fused_lstm_cell = fused
            w = weights()
            xx = x.clone().requires_grad_(True)
            out, h, c = Fn.PackedBiLSTMFn.apply(xx, plan, *w)
            ((out * gout).sum() + (h * ghf).sum() + (c * gcf).sum()).backward()
            torch.cuda.synchronize()
            res.append((out, h, c, xx.grad, [t.grad for t in w]))
    finally:
        ops.fused_lstm_cell = True
        ops.set_precision("fp32")
    (o0, h0, c0, dx0, gw0), (o1, h1, c1, dx1, gw1) = res
    valid = torch.arange(L, device=DEV).view(1, L) < pack.len.view(R, 1)
    assert float(o1[~valid].abs().max()) == 0.0
    for name, a, b in (("out", o1, o0), ("h_fin", h1, h0), ("c_fin", c1, c0), ("dx", dx1, dx0)):
        assert rel_err(a, b) <= 2e-3, "%s: %.3e" % (name, rel_err(a, b))
    for i, (a, b) in enumerate(zip(gw1, gw0)):
        assert rel_err(a, b) <= 2e-3, "grad %d: %.3e" % (i, rel_err(a, b))


class ReversePackedFn(torch.autograd.Function):
    """test helper: packed tokens -> padded reversed [R, L, In] with a gradient back to the packed rows"""

    @staticmethod
    def forward(ctx, x, pack, L):
        from dasa_b200 import ops
        ctx.pack, ctx.L = pack, L
        return ops.reverse_tokens_packed(x, pack.off, pack.len, L)

    @staticmethod
    def backward(ctx, g):
        pack, L = ctx.pack, ctx.L
        R = pack.nseq
        lens = pack.len.to(torch.int64)
        pos = torch.arange(L, device=g.device).view(1, L)
        src_pos = lens.view(R, 1) - 1 - pos                                      # original token index at reversed position
        ok = src_pos >= 0
        rows = (pack.off.to(torch.int64).view(R, 1) + src_pos.clamp(min=0))[ok]
        dx = torch.zeros(pack.ntok, g.shape[2], device=g.device)
        dx[rows] = g[ok]
        return dx, None, None


# ------------------------------------------------------------------------------------- fp16-operand tcgen05 GEMM
@pytest.mark.parametrize("M,N,K,epi,half_out", [(19810, 3072, 768, "gelu", True), (5003, 768, 3072, "bias", False),
                                                (7000, 2304, 768, "bias", False), (3000, 1536, 768, "none", False)])
def test_gemm_f16_pair(M, N, K, epi, half_out):
    """dasa_gemm_f16 (tcgen05 kind::f16, CTA-pair kernel) against fp64 on the SAME fp16-rounded operands: products of fp16 values
    are exact in the fp32 accumulator, so only the accumulation order separates the two (1e-5); the fp16 output adds its own
    2^-11 rounding."""
    import math
    from dasa_b200 import lib, ops
    g = torch.Generator().manual_seed(M + N)
    a = (torch.randn(M, K, generator=g)).to(DEV)
    w = (torch.randn(N, K, generator=g) * K ** -0.5).to(DEV)
    b = torch.randn(N, generator=g).to(DEV) if epi != "none" else None
    a16, w16 = ops.to_half(a), ops.to_half(w)
    assert torch.equal(a16, a.half()) and torch.equal(w16, w.half())
    ops.set_precision("tf32")
    lib.load().dasa_debug_gemm_pair(2)
    lib.gemm_route_counts(reset=True)
    try:
        y = ops.linear_f16(a16, w16, b, ops.EPI_BIAS_GELU if epi == "gelu" else None, out_half=half_out)
        torch.cuda.synchronize()
    finally:
        lib.load().dasa_debug_gemm_pair(1)
        ops.set_precision("fp32")
    assert lib.gemm_route_counts(reset=True)["pair_f16"] == 1
    ref = a16.double() @ w16.double().t()
    if b is not None:
        ref = ref + b.double()
    if epi == "gelu":
        ref = 0.5 * ref * (1.0 + torch.erf(ref / math.sqrt(2.0)))
    e = rel_err(y, ref)
    assert e <= (1e-3 if half_out else 1e-5), e


# ------------------------------------------------------------------------- fp16 attention + in-place dropout (frozen stack)
def _round16(n):
    return (n + 15) // 16 * 16


@pytest.mark.parametrize("stream", [False, True])
def test_mha_h16_matches_fp32_kernel(stream):
    """dasa_mha_fwd_h16 (fp16 q/k/v, persistent cp.async ring, m16n8k16) against the exact-fp32 attention kernel on the SAME
    fp16-representable inputs: packed self-attention, packed q over dense views, dense q over packed keys, dense with key_pad.
    stream=True: the kernel draws its own keep flags; the reference gets the mask materialised from the same stream through
    the documented index map. Bound 3e-3 (P rounded to fp16 before P.V; everything else fp32)."""
    gen = g(91)
    B, L, V, heads, dh = 7, 45, 36, 12, 64
    Hd = heads * dh
    lens = torch.randint(3, L + 1, (B,), generator=gen)
    lens[0], lens[1] = L, 1
    pad = (torch.arange(L)[None, :] >= lens[:, None]).to(DEV)
    x = torch.randn(B, L, 3 * Hd, generator=gen).half()
    vis = torch.randn(B, V, 3 * Hd, generator=gen).half()
    rows = torch.cat([torch.arange(n) + b * L for b, n in enumerate(lens.tolist())]).to(DEV)
    off = torch.cat([torch.zeros(1, dtype=torch.long), lens.cumsum(0)[:-1]]).int().to(DEV)
    len32 = lens.int().to(DEV)
    xh, vh = x.to(DEV), vis.to(DEV)
    xf, vf = xh.float(), vh.float()
    xp = xh.view(B * L, -1)[rows].contiguous()
    p, seed, base = 0.1, 4242, 1000

    def drop(Lq, Lk):
        n = ops.mha_h16_stream_bytes(B, heads, Lq, Lk)
        assert n % 16 == 0
        flat = ops.dropout_mask((n,), p, seed, base)
        km = flat[ops.mha_h16_stream_index(B, heads, Lq, Lk, DEV)].contiguous()
        assert 0.85 < float(km.float().mean()) < 0.95
        return km, (ops.DropStream(None, seed, base, p) if stream else km)

    ops.set_precision("fp32")
    km, d = drop(L, L)
    ref = ops.mha_fwd(xf[..., :Hd], xf[..., Hd:2 * Hd], xf[..., 2 * Hd:], heads, pad, km, 1 / 0.9)
    got = ops.mha_fwd_h16(xp[:, :Hd], xp[:, Hd:2 * Hd], xp[:, 2 * Hd:], heads, (off, len32), (off, len32), L, L, None, d, 1 / 0.9)
    assert got.dtype == torch.float16
    assert_close(got.float(), ref.view(B * L, Hd)[rows], 3e-3, "packed q and k")
    got32 = ops.mha_fwd_h16(xp[:, :Hd], xp[:, Hd:2 * Hd], xp[:, 2 * Hd:], heads, (off, len32), (off, len32), L, L, None, d, 1 / 0.9,
                            out_half=False)
    assert_close(got32, ref.view(B * L, Hd)[rows], 3e-3, "packed q and k, fp32 out")
    km, d = drop(L, V)
    ref = ops.mha_fwd(xf[..., :Hd], vf[..., Hd:2 * Hd], vf[..., 2 * Hd:], heads, None, km, 1 / 0.9)
    got = ops.mha_fwd_h16(xp[:, :Hd], vh[..., Hd:2 * Hd], vh[..., 2 * Hd:], heads, (off, len32), None, L, V, None, d, 1 / 0.9)
    assert_close(got.float(), ref.view(B * L, Hd)[rows], 3e-3, "packed q, dense k")
    km, d = drop(V, L)
    ref = ops.mha_fwd(vf[..., :Hd], xf[..., Hd:2 * Hd], xf[..., 2 * Hd:], heads, pad, km, 1 / 0.9)
    got = ops.mha_fwd_h16(vh[..., :Hd], xp[:, Hd:2 * Hd], xp[:, 2 * Hd:], heads, None, (off, len32), V, L, None, d, 1 / 0.9)
    assert_close(got.float(), ref, 3e-3, "dense q, packed k")
    got = ops.mha_fwd_h16(vh[..., :Hd], xh[..., Hd:2 * Hd], xh[..., 2 * Hd:], heads, key_pad=pad, drop=d, drop_scale=1 / 0.9)
    assert_close(got.float(), ref, 3e-3, "dense q, dense k + key_pad")
    # no dropout (eval)
    ref = ops.mha_fwd(vf[..., :Hd], vf[..., Hd:2 * Hd], vf[..., 2 * Hd:], heads)
    got = ops.mha_fwd_h16(vh[..., :Hd], vh[..., Hd:2 * Hd], vh[..., 2 * Hd:], heads)
    assert_close(got.float(), ref, 3e-3, "views, eval")


def test_mha_h16_many_units_and_long_keys():
    """More (sample, head) units than resident CTAs (the ring wraps, prefetch of the next unit under the current one) and 80-key
    instructions (5 key blocks)."""
    gen = g(92)
    B, L, heads, dh = 300, 80, 12, 64
    Hd = heads * dh
    lens = torch.randint(8, L + 1, (B,), generator=gen)
    lens[5] = L
    rows = torch.cat([torch.arange(n) + b * L for b, n in enumerate(lens.tolist())]).to(DEV)
    off = torch.cat([torch.zeros(1, dtype=torch.long), lens.cumsum(0)[:-1]]).int().to(DEV)
    len32 = lens.int().to(DEV)
    xp = (torch.randn(int(lens.sum()), 3 * Hd, generator=gen) * 0.7).half().to(DEV)
    xpad = torch.zeros(B * L, 3 * Hd, device=DEV)
    xpad[rows] = xp.float()
    xpad = xpad.view(B, L, 3 * Hd)
    pad = (torch.arange(L)[None, :] >= lens[:, None]).to(DEV)
    ops.set_precision("fp32")
    ref = ops.mha_fwd(xpad[..., :Hd], xpad[..., Hd:2 * Hd], xpad[..., 2 * Hd:], heads, pad)
    got = ops.mha_fwd_h16(xp[:, :Hd], xp[:, Hd:2 * Hd], xp[:, 2 * Hd:], heads, (off, len32), (off, len32), L, L)
    assert_close(got.float(), ref.view(B * L, Hd)[rows], 3e-3, "300 x 12 units")


@pytest.mark.parametrize("x_half", [False, True])
@pytest.mark.parametrize("Hd", [768, 256])
def test_layernorm_fwd_stream_dropout_bit_equal_to_mask(x_half, Hd):
    """dasa_dropout_residual_layernorm_fwd: in-place draws == the materialised mask of the same stream, bit for bit; fp16 x; and
    both equal the training-path kernel (dasa_dropout_residual_layernorm) on the same mask."""
    gen = g(93)
    R = 1037
    x = torch.randn(R, Hd, generator=gen).to(DEV)
    if x_half:
        x = x.half()
    resid = torch.randn(R, Hd, generator=gen).to(DEV)
    gamma, beta = torch.rand(Hd, generator=gen).to(DEV) + 0.5, torch.randn(Hd, generator=gen).to(DEV)
    p, seed, base = 0.1, 99, 77
    n = (R * Hd + 15) // 16 * 16
    km = ops.dropout_mask((n,), p, seed, base)[:R * Hd].view(R, Hd)
    a, a16 = ops.dropout_residual_layernorm_fwd(x, resid, gamma, beta, 1e-12, ops.DropStream(None, seed, base, p), 1 / 0.9, True)
    b, b16 = ops.dropout_residual_layernorm_fwd(x, resid, gamma, beta, 1e-12, km.contiguous(), 1 / 0.9, True)
    assert torch.equal(a, b) and torch.equal(a16, b16)
    c = ops.dropout_residual_layernorm(x.float(), resid, gamma, beta, 1e-12, km.contiguous(), 1 / 0.9)
    assert_close(a, c, 2e-6, "forward-only kernel (fused multiply-adds) vs the training-path kernel")
    seed_dev = torch.tensor([seed], dtype=torch.int64, device=DEV)
    d = ops.dropout_residual_layernorm_fwd(x, resid, gamma, beta, 1e-12, ops.DropStream(seed_dev, 0, base, p), 1 / 0.9)
    assert torch.equal(a, d)
    z = x.float() * km.float() / 0.9 + resid
    ref = torch.nn.functional.layer_norm(z.double(), (Hd,), gamma.double(), beta.double(), 1e-12)
    assert_close(a, ref, 2e-5, "LN")
    assert_close(a16.float(), ref, 1e-3, "LN fp16 copy")


@pytest.mark.parametrize("stream", [False, True])
def test_packed_bilstm_fused_output_dropout(stream):
    """`ctx = drop(ctx)` (r2rmodel.py:2357) fused into the packed recurrence's output write and gradient read: identical (bit for
    bit) to the unfused PackedBiLSTMFn followed by Fn.dropout with the same keep mask - a mask tensor, or the flags the kernels
    draw in place (materialised here from the same stream)."""
    R, L, In, H, seed = 90, 14, 64, 64, 5
    g = torch.Generator().manual_seed(seed)
    lens = torch.randint(1, L + 1, (R,), generator=g).tolist()
    lens[0] = L
    pack = M.PackInfo(lens, L, 1, DEV)
    plan = pack.bilstm_plan()
    x = (torch.randn(pack.ntok, In, generator=g) * 0.5).to(DEV)

    def weights():
        gg = torch.Generator().manual_seed(seed + 100)
        return [(torch.randn(*s, generator=gg) * sc).to(DEV).requires_grad_(True)
                for s, sc in (((4 * H, In), In ** -0.5), ((4 * H, H), H ** -0.5), ((4 * H,), 0.1), ((4 * H,), 0.1)) * 2]
    gout = torch.randn(R, L, 2 * H, generator=g).to(DEV)
    p, sd, base = 0.5, 31, 64
    n = R * L * 2 * H
    assert n % 16 == 0
    km = ops.dropout_mask((n,), p, sd, base).view(R, L, 2 * H)
    drop = ops.DropStream(None, sd, base, p) if stream else km
    ops.set_precision("tf32")
    try:
        w0, x0 = weights(), x.clone().requires_grad_(True)
        o0, _, _ = Fn.PackedBiLSTMFn.apply(x0, plan, *w0)
        o0 = Fn.dropout(o0, km, 2.0)
        (o0 * gout).sum().backward()
        w1, x1 = weights(), x.clone().requires_grad_(True)
        o1, _, _ = Fn.PackedBiLSTMFn.apply(x1, plan, *w1, drop, 2.0)
        (o1 * gout).sum().backward()
        torch.cuda.synchronize()
    finally:
        ops.set_precision("fp32")
    assert 0.4 < float((o1 == 0).float().mean()) and torch.equal(o0, o1)
    assert torch.equal(x0.grad, x1.grad)
    for a, b in zip(w0, w1):
        assert torch.equal(a.grad, b.grad)


@pytest.mark.parametrize("B,C", [(5, 2048), (20, 2048), (300, 2048), (9, 3072)])
@pytest.mark.parametrize("noise", [False, True])
def test_fused_gate_shift_attention_matches_unfused(B, C, noise):
    """dasa_gate_shift_attention_fwd (K1 epilogue -> K3 without materialising df_t) against gate_modulate followed by the shift
    attention on the materialised df_t, and against an fp64 restatement (agent_dg.py:1544-1547, model.py:327-345)."""
    gen = g(B + C)
    V, A, k, Hn = 36, 128, 5, 12
    F = C + A
    f = torch.rand(B, V, F, generator=gen).to(DEV)
    gp = torch.randn(B, V, C, generator=gen).to(DEV)
    t = (torch.randn(B, F, generator=gen) * 0.05).to(DEV)
    kl = torch.randn(B, k, generator=gen).to(DEV)
    cs = (torch.rand(C, generator=gen) > 0.4).float().div(0.6).to(DEV) if noise else None
    wc, attn, q, kap = ops.gate_shift_attention_fwd(f, gp, t, kl, k, Hn, cs)
    df = f.clone()
    ops.gate_modulate(gp.view(B * V, C), f[..., :C], df[..., :C])
    if noise:
        df[..., :C] *= cs
    wc0, attn0, q0, kap0 = ops.row_attention_fwd(df, t, None, k, Hn, kl)
    for name, a, b in (("wc", wc, wc0), ("attn", attn, attn0), ("q", q, q0), ("kappa", kap, kap0)):
        assert_close(a, b, 2e-6, "fused vs unfused " + name)
    # fp64 restatement
    d64 = f.double().clone()
    d64[..., :C] = torch.sigmoid(gp.double()) * f[..., :C].double() * (cs.double() if noise else 1.0)
    p = torch.softmax(torch.einsum("bvf,bf->bv", d64, t.double()), 1)
    kp = torch.softmax(kl.double(), 1)
    pe = p.view(B, V // Hn, Hn)
    qq = torch.zeros_like(pe)
    for j in range(k):
        qq += kp[:, j].view(B, 1, 1) * torch.roll(pe, shifts=-(j - k // 2), dims=2)
    ref = torch.einsum("bv,bvf->bf", qq.view(B, V), d64)
    assert_close(attn, p, 1e-4, "softmax over the views")
    assert_close(wc, ref, 1e-4, "weighted context")
