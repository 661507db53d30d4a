"""GPU tier: the persistent decoder-rollout kernel (csrc/decoder_persist.cu, dasa_decoder_rollout_fwd / _bwd) against an
fp64 torch restatement of BAttnDecoderLSTM.forward (model.py:472-554) over T actions, every intermediate buffer and every
gradient; then through the drop-in module (per-action T = 1 form) and the teacher-forced rollout against the CPU oracle.

Tolerances: the kernel multiplies in TF32 (10-bit mantissa operands, fp32 accumulate). For a K-term dot product of O(1)
operands the rounding error is bounded by 2 * 2^-11 * sum|x_k w_k| and behaves like sqrt(K) * 2^-11 * rms; the asserts below use
the normwise bound 1e-2 for forward buffers (measured: 2e-3 .. 5.5e-3 at full geometry after three recurrent actions, the
softmax over 2176-term logits being the amplifier) and for gradients that went through the T-step recurrence (north_star: tensor
-core paths within 1e-2)."""
import pytest
import torch

from dasa_b200 import synth
from dasa_b200.config import FULL, SMALL

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from dasa_b200 import functions as Fn
    from dasa_b200 import modules as M
    from dasa_b200 import ops
    from dasa_b200.rollout import DeviceEpisodes, NavPolicy

DEV = "cuda"


def rel(a, b):
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def reference_rollout(w, emb, feat, ctx, mask, h0, c0, m_hp, m_h1, scale, headings=12):
    """fp64 torch restatement; returns dict of per-action stacks (same names as dasa_decoder_fwd_t)."""
    T, B, _ = emb.shape
    F_all, k = w["w_in"].shape[0], w["w_shift"].shape[0]
    half = k // 2
    V = feat.shape[2]
    E_, Hn = V // headings, headings
    out = {n: [] for n in ("tk", "p", "q", "kappa", "attn", "acts", "c", "h1", "t2", "alpha", "wc", "htilde")}
    h_prev, c_prev = h0, c0
    for t in range(T):
        hpd = h_prev if m_hp is None else h_prev * m_hp[t] * scale
        tvec = hpd @ w["w_in"].T
        kl = hpd @ w["w_shift"].T + w["b_shift"]
        z = torch.einsum("bvf,bf->bv", feat[t], tvec)
        p = torch.softmax(z, 1)
        kappa = torch.softmax(kl, 1)
        pe = p.view(B, E_, Hn)
        q = torch.zeros_like(pe)
        for j in range(k):
            idx = (torch.arange(Hn, device=p.device) + j - half) % Hn
            q = q + kappa[:, j].view(B, 1, 1) * pe[:, :, idx]
        q = q.reshape(B, V)
        attn = torch.einsum("bv,bvf->bf", q, feat[t])
        xh = torch.cat([emb[t], attn, h_prev], 1)
        gates = xh @ torch.cat([w["w_ih"], w["w_hh"]], 1).T + w["b_ih"] + w["b_hh"]
        i, f, g, o = gates.chunk(4, 1)
        i, f, g, o = torch.sigmoid(i), torch.sigmoid(f), torch.tanh(g), torch.sigmoid(o)
        c1 = f * c_prev + i * g
        h1 = o * torch.tanh(c1)
        h1d = h1 if m_h1 is None else h1 * m_h1[t] * scale
        t2 = h1d @ w["w_att_in"].T
        z2 = torch.einsum("bld,bd->bl", ctx[t], t2)
        if mask is not None:
            z2 = z2.masked_fill(mask.bool(), float("-inf"))
        alpha = torch.softmax(z2, 1)
        wc = torch.einsum("bl,bld->bd", alpha, ctx[t])
        ht = torch.tanh(torch.cat([wc, h1d], 1) @ w["w_att_out"].T)
        for n, v in (("tk", torch.cat([tvec, kl], 1)), ("p", p), ("q", q), ("kappa", kappa), ("attn", attn),
                     ("acts", torch.cat([i, f, g, o], 1)), ("c", c1), ("h1", h1), ("t2", t2), ("alpha", alpha), ("wc", wc),
                     ("htilde", ht)):
            out[n].append(v)
        h_prev, c_prev = ht, c1
    return {n: torch.stack(v) for n, v in out.items()}


def _problem(cfg, B, T, L, seed, train):
    g = torch.Generator().manual_seed(seed)
    H, E, F, D, k = cfg.hidden, cfg.action_emb, cfg.feat, cfg.ctx_dim, cfg.shift_kernel
    V = cfg.views

    def rn(*s, scale=1.0):
        return (torch.randn(*s, generator=g) * scale).to(DEV)
    # attention projections scaled so that the view / token logits are O(1..5) like the policy's own (synth.policy_state): with
    # unit-variance logits of 2176 terms the softmax is a near one-hot whose gradient amplifies the TF32 rounding of the logits
    # ~20x, which measures the conditioning of the synthetic problem rather than the kernel
    att = 0.15 if F >= 1024 else 1.0
    w = {"w_in": rn(F, H, scale=att * H ** -0.5), "w_shift": rn(k, H, scale=H ** -0.5), "b_shift": rn(k, scale=0.1),
         "w_ih": rn(4 * H, E + F, scale=(E + F) ** -0.5), "w_hh": rn(4 * H, H, scale=H ** -0.5), "b_ih": rn(4 * H, scale=0.1),
         "b_hh": rn(4 * H, scale=0.1), "w_att_in": rn(D, H, scale=att * H ** -0.5), "w_att_out": rn(H, D + H, scale=(D + H) ** -0.5)}
    emb = torch.tanh(rn(T, B, E))
    feat = rn(T, B, V, F, scale=0.5).abs()
    ctx = rn(T, B, L, D, scale=0.5)
    lens = torch.randint(max(1, L // 3), L + 1, (B,), generator=g)
    lens[0] = L
    mask = (torch.arange(L).unsqueeze(0) >= lens.unsqueeze(1)).to(DEV)
    ctx = ctx * (~mask).view(1, B, L, 1)               # pad_packed_sequence zero rows
    h0, c0 = torch.tanh(rn(B, H)), rn(B, H, scale=0.5)
    m_hp = m_h1 = None
    if train:
        m_hp = (torch.rand(T, B, H, generator=g) >= cfg.dropout).to(torch.uint8).to(DEV)
        m_h1 = (torch.rand(T, B, H, generator=g) >= cfg.dropout).to(torch.uint8).to(DEV)
    gh = rn(T, B, H)                                    # upstream gradient of every h_tilde
    return w, emb, feat, ctx, mask, h0, c0, m_hp, m_h1, gh


@pytest.mark.parametrize("cfg,B,T,L,train", [(SMALL, 3, 3, 11, False), (SMALL, 5, 4, 24, True), (SMALL, 20, 2, 17, True),
                                              (FULL, 20, 3, 80, True), (FULL, 9, 2, 37, False)])
def test_kernel_matches_fp64_reference(cfg, B, T, L, train):
    w, emb, feat, ctx, mask, h0, c0, m_hp, m_h1, gh = _problem(cfg, B, T, L, 7 + B, train)
    scale = 1.0 / (1.0 - cfg.dropout)
    # ---- fp64 reference with autograd
    leaf = {k: v.double().requires_grad_(True) for k, v in w.items()}
    emb64, feat64, ctx64 = (x.double().requires_grad_(True) for x in (emb, feat, ctx))
    h64, c64 = h0.double().requires_grad_(True), c0.double().requires_grad_(True)
    ref = reference_rollout(leaf, emb64, feat64, ctx64, mask, h64, c64, None if m_hp is None else m_hp.double(),
                            None if m_h1 is None else m_h1.double(), scale)
    (ref["htilde"] * gh.double()).sum().backward()
    # ---- kernel
    ops.set_precision("tf32")
    try:
        geom = (B, cfg.hidden, cfg.action_emb, cfg.feat, cfg.views, L, cfg.ctx_dim, (cfg.feat + cfg.shift_kernel + 31) // 32 * 32,
                cfg.shift_kernel)
        assert ops.decoder_rollout_supported(*geom), "geometry should be supported: %s" % (geom,)
        prm = {k: v.clone().requires_grad_(True) for k, v in w.items()}
        e2, f2, x2 = (x.clone().requires_grad_(True) for x in (emb, feat, ctx))
        h2, c2 = h0.clone().requires_grad_(True), c0.clone().requires_grad_(True)
        Wf, bf = ops.stacked_weights((prm["w_in"], prm["w_shift"]), 0, 32, (None, prm["b_shift"]))
        Wl, _ = ops.stacked_weights((prm["w_ih"], prm["w_hh"]), 1)
        raw = ops.decoder_rollout_fwd(emb, feat, ctx, mask, h0, c0, m_hp, m_h1, scale, Wf, bf, Wl, prm["b_ih"], prm["b_hh"],
                                      prm["w_att_in"], prm["w_att_out"], 12, cfg.shift_kernel)
        F_all, k, E = cfg.feat, cfg.shift_kernel, cfg.action_emb
        D = cfg.ctx_dim
        checks = [("tk", raw["tk"][..., :F_all + k], ref["tk"]), ("p", raw["p"], ref["p"]), ("q", raw["q"], ref["q"]),
                  ("kappa", raw["kappa"], ref["kappa"]), ("attn_feat", raw["xh"][..., E:E + F_all], ref["attn"]),
                  ("acts", raw["acts"], ref["acts"]), ("c", raw["c"][1:], ref["c"]), ("h1", raw["h1"], ref["h1"]),
                  ("t2", raw["t2"], ref["t2"]), ("alpha", raw["alpha"], ref["alpha"]), ("wc", raw["cat"][..., :D], ref["wc"]),
                  ("htilde", raw["htilde"], ref["htilde"])]
        errs = {n: rel(a, b) for n, a, b in checks}
        print("forward errors:", {n: "%.2e" % e for n, e in errs.items()})
        for n, e in errs.items():
            assert e <= 1e-2, "forward buffer %s: normwise error %.3e (all: %s)" % (n, e, errs)
        assert float(raw["alpha"][:, mask].abs().max() if bool(mask.any()) else 0.0) == 0.0        # masked tokens: exactly 0
        # ---- autograd Function: outputs + every gradient
        ht, h1, cc = Fn.DecoderRolloutFn.apply(e2, f2, x2, mask.to(torch.uint8), h2, c2, m_hp, m_h1, scale, prm["w_in"],
                                               prm["w_shift"], prm["b_shift"], prm["w_ih"], prm["w_hh"], prm["b_ih"], prm["b_hh"],
                                               prm["w_att_in"], prm["w_att_out"], 12)
        assert torch.equal(ht, raw["htilde"]), "two launches on identical inputs must agree bit for bit (deterministic folds)"
        (ht * gh).sum().backward()
        torch.cuda.synchronize()
    finally:
        ops.set_precision("fp32")
    gerrs = {}
    for n, got, want in [("emb", e2.grad, emb64.grad), ("feat", f2.grad, feat64.grad), ("ctx", x2.grad, ctx64.grad),
                         ("h0", h2.grad, h64.grad), ("c0", c2.grad, c64.grad)] + [(k_, prm[k_].grad, leaf[k_].grad) for k_ in w]:
        assert got is not None, "missing gradient %s" % n
        gerrs[n] = rel(got, want)
    print("gradient errors:", {n: "%.2e" % e for n, e in gerrs.items()})
    for n, e in gerrs.items():
        assert e <= 1e-2, "gradient %s: normwise error %.3e (all: %s)" % (n, e, gerrs)
    if bool(mask.any()):
        assert float(x2.grad[:, mask].abs().max()) == 0.0                  # masked context rows receive an exact zero gradient


def test_unsupported_geometries_are_reported():
    ops.set_precision("tf32")
    try:
        assert not ops.decoder_rollout_supported(512, 1024, 64, 2176, 36, 80, 2048, 2208, 5)      # B > 32: per-op path
        assert ops.decoder_rollout_supported(20, 1024, 64, 2176, 36, 80, 2048, 2208, 5)
    finally:
        ops.set_precision("fp32")
    assert not ops.decoder_rollout_supported(20, 1024, 64, 2176, 36, 80, 2048, 2208, 5)             # exact-fp32 mode: FFMA per-op path


@pytest.mark.parametrize("train", [False, True])
def test_module_step_and_rollout_use_the_persistent_kernel(train):
    """The drop-in BAttnDecoderLSTM.forward (T = 1 form) and the batched teacher-forced rollout, TF32 precision, against the CPU
    oracle with the same injected dropout masks: logits, loss, every trainable gradient; and the per-action schedule (one
    cooperative launch per action) must agree with the whole-rollout launch."""
    from oracle import restated as R
    from tests.test_gpu_policy import _train_masks
    cfg, B, T = SMALL, 4, 3
    st = synth.policy_state(cfg, 5)
    ep = synth.Episodes(B, T, cfg, seed=41)
    L, nc = ep.seq_mask.shape[1], ep.cand_feat.shape[2]
    ost = {grp: {k: v.clone().requires_grad_(True) for k, v in d.items()} for grp, d in st.items()}
    keep = _train_masks(cfg, B, T, L, nc, 99) if train else {}
    drops = R.MaskDrops({k: m.float() / (1 - p) for k, (m, p) in keep.items()}) if train else R.NoDrop()
    loss, logits, _ = R.teacher_rollout(ost, cfg, ep, T, drops=drops)
    loss.backward()
    pol = NavPolicy(cfg, st)
    pol = pol.train() if train else pol.eval()
    dep = DeviceEpisodes(ep)
    src = M.DropoutSource(injected={k: m for k, (m, p) in keep.items()}) if train else M.DropoutSource()
    ops.set_precision("tf32")
    calls = []
    orig = ops.call

    def spy(name, *a):
        calls.append(name)
        return orig(name, *a)
    ops.call = spy
    try:
        with M.use_dropout_source(src):
            loss_b, logits_b, _ = pol.teacher_rollout(dep, T, schedule="batched")
            n_batched = calls.count("dasa_decoder_rollout_fwd")
            loss_s, logits_s, _ = pol.teacher_rollout(dep, T, schedule="sequential")
        n_seq = calls.count("dasa_decoder_rollout_fwd") - n_batched
        loss_b.backward()
        torch.cuda.synchronize()
    finally:
        ops.call = orig
        ops.set_precision("fp32")
    assert n_batched == 1 and n_seq == T, (n_batched, n_seq)
    assert calls.count("dasa_decoder_rollout_bwd") == 1
    assert rel(loss_b, loss) <= 1e-2 and rel(loss_s, loss) <= 1e-2
    lg = torch.stack(logits).detach()
    fin = torch.isfinite(lg)
    for name, l2 in (("batched", logits_b), ("sequential", logits_s)):
        l2 = torch.stack(l2).detach().cpu()
        assert torch.equal(torch.isfinite(l2), fin)
        assert rel(l2[fin], lg[fin]) <= 1e-2, name
    ls_, lb_ = torch.stack(logits_s).detach().cpu(), torch.stack(logits_b).detach().cpu()
    assert rel(ls_[fin], lb_[fin]) <= 2e-3
    for grp, mod in (("adaIn", pol.adaIn), ("decoder", pol.decoder), ("encoder", pol.encoder)):
        for k, prm in mod.named_parameters():
            want = ost[grp][k].grad
            if want is None or float(want.abs().max()) == 0.0:
                continue
            gd, wd = prm.grad.detach().double().cpu(), want.double()
            e2 = float((gd - wd).norm() / wd.norm())
            assert e2 <= 1e-2, "grad %s.%s L2 rel err %.3e" % (grp, k, e2)
