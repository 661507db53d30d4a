"""TEST INFRASTRUCTURE ONLY. Generates tests/golden/*.pt from the REAL reference modules.

Run in the build container (needs /root/reference):   python -m oracle.make_golden
The reference cannot travel to the GPU box, so its outputs on seeded synthetic inputs (dasa_b200/synth.py) are
committed as small fixtures. Weights are NOT stored: they are regenerated from the seed by synth.policy_state and
loaded into the reference modules with load_state_dict (keys/shapes must match exactly — that is itself a check of
the drop-in state_dict contract).

The agent's own loop (agent_dg.py:633-1033) hard-codes .cuda() and needs the simulator, so the rollout fixtures
drive the reference's *modules* with a restated loop body (agent_dg.py:727-851); everything numeric inside a step
is the reference's code.
"""
import contextlib
import io
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from dasa_b200 import synth                                   # noqa: E402
from dasa_b200.config import FULL, SMALL                      # noqa: E402
from oracle import load_reference                             # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def build_reference_modules(ref, cfg, state, adain_kind="channel"):
    ref.BertConfig.OVERRIDES = dict(hidden_size=cfg.bert_hidden, num_attention_heads=cfg.bert_heads,
                                    intermediate_size=cfg.bert_inter, vocab_size=cfg.vocab,
                                    max_position_embeddings=cfg.max_pos)
    a = ref.args
    a.critic_dim, a.angle_feat_size, a.shift_kernel_size = cfg.critic_dim, cfg.angle_size, cfg.shift_kernel
    a.featdropout, a.dropout, a.use_shift = cfg.featdropout, cfg.dropout, True
    enc = ref.r2rmodel.DicEncoder(cfg.feat, cfg.enc_hidden, cfg.hidden, cfg.enc_dropout, True, False, 1, True, True,
                                  cfg.vl_layers, cfg.la_layers, "small", cfg.update_add_layer)
    dec = ref.model.BAttnDecoderLSTM(cfg.action_emb, cfg.hidden, cfg.dropout, feature_size=cfg.feat)
    cri = ref.model.Critic()
    ada = {"channel": ref.agent_dg.DGAdaChannel, "stat": ref.agent_dg.DGAdaStatChannel,
           "mean": ref.agent_dg.DGAdaMeanChannel}[adain_kind](cfg.rgb_size)
    for m, k in ((enc, "encoder"), (dec, "decoder"), (cri, "critic"), (ada, "adaIn")):
        m.load_state_dict(state[k], strict=True)
    return enc, dec, cri, ada


@contextlib.contextmanager
def recorded_dropout(seed, record):
    """Replace F.dropout by a seeded mask generator that records every mask (in call order)."""
    g = torch.Generator().manual_seed(seed)
    orig = F.dropout

    def fake(x, p=0.5, training=True, inplace=False):
        if not training or p == 0.0:
            return x
        m = (torch.rand(x.shape, generator=g) >= p).to(x.dtype) / (1.0 - p)
        record.append(m)
        return x * m

    F.dropout = fake
    try:
        yield
    finally:
        F.dropout = orig


def reference_rollout(ref, mods, cfg, ep, T, ml_weight=0.4, noise=None):
    """Loop body of vl_rollout (agent_dg.py:727-851), feedback='teacher', around the reference modules. noise (a [C] tensor)
    switches on consistent_drop with --env_drop_stage after_adain --depth_drop (agent_dg.py:780-785, 812-820)."""
    enc, dec, cri, ada = mods
    C = cfg.rgb_size
    crit = torch.nn.CrossEntropyLoss(ignore_index=cfg.ignore_id, reduction="sum")   # agent_dg.py:250
    total, logits, hs = 0.0, [], []
    for t in range(T):
        a_t, f_t, d_t, cand, cand_d, leng, target = [x.clone() for x in ep.step(t)]
        df_t = f_t.clone()
        df_t[:, :, :C] = ada(f_t[:, :, :C].clone(), d_t[:, :, :C].clone())
        cand[:, :, :C] = ada(cand[:, :, :C].clone(), cand_d[:, :, :C].clone())
        consistent_drop = noise is not None
        if consistent_drop:
            cand[..., :C] *= noise
            f_t[..., :C] *= noise
            cand_d[..., :C] *= noise
            df_t[..., :C] *= noise
        ctx, en_h, en_c, _, _ = enc(ep.seq, mask=ep.seq_mask, lengths=ep.seq_lengths, f_t_all=f_t.clone())
        if t == 0:
            h_t, c_t, logit, h1, _ = dec(a_t, df_t, cand, en_h, en_h, en_c, ctx, ep.seq_mask, already_dropfeat=consistent_drop)
        else:
            h_t, c_t, logit, h1, _ = dec(a_t, df_t, cand, h_t, h1, c_t, ctx, ep.seq_mask, already_dropfeat=consistent_drop)
        cmask = torch.arange(logit.shape[1]).unsqueeze(0) >= leng.view(-1, 1).long()   # utils.length2mask
        logit.masked_fill_(cmask, -float("inf"))
        total = total + crit(logit, target)
        logits.append(logit)
        hs.append(h_t)
    return total * ml_weight / ep.B, logits, hs


def sample(t, n=257):
    """Strided sample of a big tensor (keeps fixtures small)."""
    f = t.detach().reshape(-1)
    step = max(1, f.numel() // n)
    return f[::step][:n].clone()


def pack_masks(masks):
    """bit-packed keep flags + shapes (unpack with oracle.make_golden.unpack_masks)."""
    import numpy as np
    return [(torch.from_numpy(np.packbits((m != 0).numpy().reshape(-1))), tuple(m.shape)) for m in masks]


def unpack_masks(packed, p_of_shape=None):
    """-> list of bool keep tensors."""
    import numpy as np
    out = []
    for bits, shape in packed:
        n = 1
        for d in shape:
            n *= d
        out.append(torch.from_numpy(np.unpackbits(bits.numpy())[:n].astype(bool)).view(*shape))
    return out


def main():
    with contextlib.redirect_stdout(io.StringIO()):
        ref = load_reference.load()
    os.makedirs(GOLDEN, exist_ok=True)
    torch.set_num_threads(8)

    # ------------------------------------------------------------------ 1. per-module, SMALL geometry, eval
    cfg, seed = SMALL, 0
    state = synth.policy_state(cfg, seed)
    mods = build_reference_modules(ref, cfg, state)
    enc, dec, cri, ada = mods
    for m in mods:
        m.eval()
    ep = synth.Episodes(5, 4, cfg, seed=3)
    out = {"meta": {"cfg": "SMALL", "seed": seed, "episodes": dict(B=5, T=4, seed=3)}}
    C = cfg.rgb_size
    with torch.no_grad():
        a_t, f_t, d_t, cand, cand_d, leng, target = ep.step(1)
        out["adain_channel"] = ada(f_t[..., :C].clone(), d_t[..., :C].clone())
        for kind in ("stat", "mean"):
            st2 = synth.adain_state(cfg, seed, kind)
            m2 = {"stat": ref.agent_dg.DGAdaStatChannel, "mean": ref.agent_dg.DGAdaMeanChannel}[kind](C)
            m2.load_state_dict(st2)
            out["adain_" + kind] = m2(f_t[..., :C].clone(), d_t[..., :C].clone())
        out["adain_default"] = ref.model.adaptive_instance_normalization(f_t[..., :C], d_t[..., :C])
        g = torch.Generator().manual_seed(11)
        h = torch.tanh(torch.randn(5, cfg.hidden, generator=g))
        out["h_query"] = h
        wc, p = dec.feat_att_layer(h, f_t.clone(), output_tilde=False)
        out["shift_wc"], out["shift_attn"] = wc, p
        ctx, en_h, en_c, _, vis = enc(ep.seq, mask=ep.seq_mask, lengths=ep.seq_lengths, f_t_all=f_t.clone())
        out["enc_ctx"], out["enc_h"], out["enc_c"], out["enc_vis"] = ctx, en_h, en_c, vis
        ht, alpha = dec.attention_layer(h, ctx, ep.seq_mask)
        out["softdot_h"], out["softdot_alpha"] = ht, alpha
        _, lg = dec.candidate_att_layer(h, cand.clone(), output_prob=False)
        out["cand_logit"] = lg
        h1, c1, lg2, h_tilde, _ = dec(a_t, f_t.clone(), cand.clone(), en_h, en_h, en_c, ctx, ep.seq_mask)
        out["dec_h1"], out["dec_c1"], out["dec_logit"], out["dec_htilde"] = h1, c1, lg2, h_tilde
        out["critic"] = cri(h1)
        loss, logits, hs = reference_rollout(ref, mods, cfg, ep, 4)
        out["rollout_eval_loss"], out["rollout_eval_logits"] = loss, torch.stack(logits)
    torch.save(out, os.path.join(GOLDEN, "small_eval.pt"))
    print("small_eval.pt written")

    # ------------------------------------------------- 2. SMALL geometry, train mode, recorded masks, gradients
    for m in mods:
        m.train()
        m.zero_grad()
    ep2 = synth.Episodes(3, 2, cfg, seed=5)
    rec = []
    with recorded_dropout(1234, rec):
        loss, logits, hs = reference_rollout(ref, mods, cfg, ep2, 2)
    loss.backward()
    tr = {"meta": {"cfg": "SMALL", "seed": seed, "episodes": dict(B=3, T=2, seed=5), "mask_seed": 1234},
          "masks": pack_masks(rec), "loss": loss.detach(), "logits": torch.stack(logits).detach(), "grads": {}}
    for name, m in (("encoder", enc), ("decoder", dec), ("adaIn", ada)):
        for k, prm in m.named_parameters():
            if prm.grad is not None:
                tr["grads"][name + "." + k] = {"norm": prm.grad.norm(), "sample": sample(prm.grad)}
    tr["no_grad_params"] = [n + "." + k for n, m in (("encoder", enc), ("decoder", dec)) for k, prm in m.named_parameters()
                            if prm.grad is None]
    torch.save(tr, os.path.join(GOLDEN, "small_train.pt"))
    print("small_train.pt written; %d masks, %d params with grad" % (len(rec), len(tr["grads"])))

    # ---------------------------------------------------------------- 3. FULL geometry, one policy step, eval
    cfg = FULL
    state = synth.policy_state(cfg, 0)
    mods = build_reference_modules(ref, cfg, state)
    for m in mods:
        m.eval()
    ep3 = synth.Episodes(3, 2, cfg, seed=7)
    with torch.no_grad():
        loss, logits, hs = reference_rollout(ref, mods, cfg, ep3, 2)
        enc, dec, cri, ada = mods
        a_t, f_t, d_t, cand, cand_d, leng, target = ep3.step(0)
        ctx, en_h, en_c, _, vis = enc(ep3.seq, mask=ep3.seq_mask, lengths=ep3.seq_lengths, f_t_all=f_t.clone())
    full = {"meta": {"cfg": "FULL", "seed": 0, "episodes": dict(B=3, T=2, seed=7)},
            "loss": loss, "logits": torch.stack(logits), "h_t": torch.stack(hs),
            "enc_h": en_h, "enc_c": en_c, "ctx_sample": sample(ctx, 4099), "vis_sample": sample(vis, 4099)}
    torch.save(full, os.path.join(GOLDEN, "full_eval.pt"))
    print("full_eval.pt written")


if __name__ == "__main__":
    main()
