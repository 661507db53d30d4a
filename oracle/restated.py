"""TEST INFRASTRUCTURE ONLY. CPU oracle for the agent_dg navigation-policy hot path.

A plain-PyTorch (CPU, fp32 or fp64) functional restatement of the reference algorithm, written against the
reference's state_dict keys so that the same seeded weights drive the reference modules (when mounted), this
oracle, and the CUDA path. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this file; the product path (dasa_b200/) must never import it.

Parity status: PINNED. The reference carries no tests/golden vectors for this path, so the
oracle is pinned against outputs of the reference's own modules imported in the build container
(oracle/load_reference.py) — tests/test_oracle_vs_reference.py (live, skipped when /root/reference is absent)
and tests/golden/*.pt (fixtures produced by oracle/make_golden.py from the real reference; checked anywhere).

Each function cites the reference lines it follows (paths relative to /root/reference/r2r_src).
"""
import math

import torch
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------------ dropout
class NoDrop:
    """eval(): every nn.Dropout is the identity."""
    training = False

    def __call__(self, x, p, tag):
        return x


class MaskDrops:
    """train() with injected, pre-scaled keep masks looked up by tag (tests feed the CUDA path the same masks)."""
    training = True

    def __init__(self, masks):
        self.masks = masks

    def __call__(self, x, p, tag):
        m = self.masks.get(tag)
        return x if m is None else x * m.to(x.dtype)


class ReplayDrops:
    """train() replaying masks in call order (recorded from a run of the real reference modules)."""
    training = True

    def __init__(self, seq):
        self.seq, self.i = list(seq), 0

    def __call__(self, x, p, tag):
        m = self.seq[self.i]
        self.i += 1
        assert m.shape == x.shape, (tag, m.shape, x.shape)
        return x * m.to(x.dtype)


# ------------------------------------------------------------------------------------- AdaIN family (a1, a2)
def adain_channel_gate(sd, f, d):
    """DGAdaChannel, ab_type='a', a_type='sigmoid' (agent_dg.py:1534-1547): sigmoid(a_fc(d)) * f."""
    return torch.sigmoid(F.linear(d, sd["a_fc.weight"], sd["a_fc.bias"])) * f


def view_stats(d):
    """mean / unbiased std / max / min over the view axis (agent_dg.py:1651-1654) -> [N, 4C]."""
    return torch.cat([d.mean(1), d.std(1), d.max(1)[0], d.min(1)[0]], -1)


def adain_stat_channel(sd, f, d):
    """DGAdaStatChannel (agent_dg.py:1649-1661)."""
    s = view_stats(d)
    a = F.linear(s, sd["a_fc.weight"], sd["a_fc.bias"]).unsqueeze(1)
    b = F.linear(s, sd["b_fc.weight"], sd["b_fc.bias"]).unsqueeze(1)
    return a * f + b


def adain_mean_channel(sd, f, d):
    """DGAdaMeanChannel (agent_dg.py:1630-1636)."""
    s = d.mean(1)
    a = F.linear(s, sd["a_fc.weight"], sd["a_fc.bias"]).unsqueeze(1)
    b = F.linear(s, sd["b_fc.weight"], sd["b_fc.bias"]).unsqueeze(1)
    return a * f + b


def adain_default(f, d, eps=1e-5):
    """model.adaptive_instance_normalization (model.py:1822-1840): stats over the channel axis per view,
    unbiased variance + eps."""
    mu_f, sd_f = f.mean(-1, keepdim=True), (f.var(-1, keepdim=True) + eps).sqrt()
    mu_d, sd_d = d.mean(-1, keepdim=True), (d.var(-1, keepdim=True) + eps).sqrt()
    return (f - mu_f) / sd_f * sd_d + mu_d


# --------------------------------------------------------------------------------------- attention (a3-a5)
def shift_probs(p, kappa, headings=12):
    """Circular cross-correlation of the view distribution along the heading axis with a per-sample kernel
    (model.py:337-344): q[b,e,l] = sum_j kappa[b,j] * p[b,e,(l + j - k//2) mod 12]."""
    B, V = p.shape
    k = kappa.shape[1]
    p3 = p.view(B, V // headings, headings)
    q = torch.zeros_like(p3)
    for j in range(k):
        q = q + kappa[:, j, None, None] * torch.roll(p3, shifts=-(j - k // 2), dims=2)
    return q.reshape(B, V)


def shift_soft_dot_attention(sd, pre, h, context, headings=12):
    """ShiftSoftDotAttention.forward with mask=None, output_tilde=False (model.py:318-353).
    Returns (weighted_context, attn) where attn is the PRE-shift softmax (model.py:336,351-353)."""
    target = F.linear(h, sd[pre + "linear_in.weight"])
    logit = torch.bmm(context, target.unsqueeze(2)).squeeze(2)
    p = torch.softmax(logit, 1)
    kappa = torch.softmax(F.linear(h, sd[pre + "linear_shift.weight"], sd[pre + "linear_shift.bias"]), -1)
    q = shift_probs(p, kappa, headings)
    wc = torch.bmm(q.unsqueeze(1), context).squeeze(1)
    return wc, p


def soft_dot_attention(sd, pre, h, context, mask=None):
    """SoftDotAttention.forward, output_tilde=True, output_prob=True (model.py:268-296)."""
    target = F.linear(h, sd[pre + "linear_in.weight"])
    logit = torch.bmm(context, target.unsqueeze(2)).squeeze(2)
    if mask is not None:
        logit = logit.masked_fill(mask.bool(), -float("inf"))
    alpha = torch.softmax(logit, 1)
    wc = torch.bmm(alpha.unsqueeze(1), context).squeeze(1)
    h_tilde = torch.tanh(F.linear(torch.cat((wc, h), 1), sd[pre + "linear_out.weight"]))
    return h_tilde, alpha


def candidate_logits(sd, pre, h, cand_feat):
    """SoftDotAttention as candidate_att_layer with output_prob=False (model.py:559): only the raw logits
    survive; softmax / weighted sum / linear_out are dead work (model.py:285-294)."""
    target = F.linear(h, sd[pre + "linear_in.weight"])
    return torch.bmm(cand_feat, target.unsqueeze(2)).squeeze(2)


# --------------------------------------------------------------------------------------------- LSTM cell (a6)
def lstm_cell(w_ih, w_hh, b_ih, b_hh, x, h, c):
    """nn.LSTMCell: gate order i,f,g,o."""
    gates = F.linear(x, w_ih, b_ih) + F.linear(h, w_hh, b_hh)
    i, f, g, o = gates.chunk(4, 1)
    c1 = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
    h1 = torch.sigmoid(o) * torch.tanh(c1)
    return h1, c1


# -------------------------------------------------------------------------------------------- decoder (a7)
def decoder_step(sd, cfg, action, feature, cand_feat, prev_h1, c_0, ctx, ctx_mask, drops=NoDrop(),
                 already_dropfeat=False):
    """BAttnDecoderLSTM.forward (model.py:472-574). `feature` / `cand_feat` are NOT mutated here; the dropped
    tensors the reference writes back in place (model.py:508,557) are returned in aux."""
    A, p = cfg.angle_size, cfg.dropout
    emb = torch.tanh(F.linear(action, sd["embedding.0.weight"], sd["embedding.0.bias"]))
    emb = drops(emb, p, "dec.act")
    if not already_dropfeat:
        feature = torch.cat([drops(feature[..., :-A], cfg.featdropout, "dec.feat"), feature[..., -A:]], -1)
    h_prev_drop = drops(prev_h1, p, "dec.h_prev")
    attn_feat, view_attn = shift_soft_dot_attention(sd, "feat_att_layer.", h_prev_drop, feature, cfg.headings)
    x = torch.cat((emb, attn_feat), 1)
    h_1, c_1 = lstm_cell(sd["lstm.weight_ih"], sd["lstm.weight_hh"], sd["lstm.bias_ih"], sd["lstm.bias_hh"],
                         x, prev_h1, c_0)
    h_1_drop = drops(h_1, p, "dec.h1")
    h_tilde, alpha = soft_dot_attention(sd, "attention_layer.", h_1_drop, ctx, ctx_mask)
    h_tilde_drop = drops(h_tilde, p, "dec.htilde")
    if not already_dropfeat:
        cand_feat = torch.cat([drops(cand_feat[..., :-A], cfg.featdropout, "dec.cand"), cand_feat[..., -A:]], -1)
    logit = candidate_logits(sd, "candidate_att_layer.", h_tilde_drop, cand_feat)
    aux = {"feature": feature, "cand_feat": cand_feat, "view_attn": view_attn, "alpha": alpha}
    return h_1, c_1, logit, h_tilde, aux


def critic(sd, state, drops=NoDrop(), p=0.5):
    """Critic.forward (model.py:974-982)."""
    x = torch.relu(F.linear(state, sd["state2value.0.weight"], sd["state2value.0.bias"]))
    x = drops(x, p, "critic")
    return F.linear(x, sd["state2value.3.weight"], sd["state2value.3.bias"]).squeeze()


# -------------------------------------------------------------------------------------------- encoder (a9)
def gelu_erf(x):
    """vilmodel.gelu (vilmodel.py:125-131): exact erf form."""
    return x * 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))


def _mha(sd, pre, names, x_q, x_kv, add_mask, heads, drops, p, tag):
    """BertSelfAttention / BertOutAttention (vilmodel.py:203-236, 479-506)."""
    B, Lq, Hb = x_q.shape
    dh = Hb // heads
    q = F.linear(x_q, sd[pre + names[0] + ".weight"], sd[pre + names[0] + ".bias"])
    k = F.linear(x_kv, sd[pre + names[1] + ".weight"], sd[pre + names[1] + ".bias"])
    v = F.linear(x_kv, sd[pre + names[2] + ".weight"], sd[pre + names[2] + ".bias"])
    q = q.view(B, Lq, heads, dh).permute(0, 2, 1, 3)
    k = k.view(B, -1, heads, dh).permute(0, 2, 1, 3)
    v = v.view(B, -1, heads, dh).permute(0, 2, 1, 3)
    s = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(dh)
    if add_mask is not None:
        s = s + add_mask
    pr = drops(torch.softmax(s, -1), p, tag + ".probs")
    o = torch.matmul(pr, v).permute(0, 2, 1, 3).contiguous().view(B, Lq, Hb)
    return o


def _out_ln(sd, pre, x, resid, eps, drops, p, tag):
    """BertSelfOutput / BertOutput (vilmodel.py:246-250, 305-309): LN(drop(dense(x)) + resid)."""
    y = drops(F.linear(x, sd[pre + "dense.weight"], sd[pre + "dense.bias"]), p, tag)
    return F.layer_norm(y + resid, (y.shape[-1],), sd[pre + "LayerNorm.weight"], sd[pre + "LayerNorm.bias"], eps)


def _self_att_block(sd, pre, x, add_mask, cfg, drops, tag):
    """BertAttention (vilmodel.py:276-280)."""
    o = _mha(sd, pre + "self.", ("query", "key", "value"), x, x, add_mask, cfg.bert_heads, drops,
             cfg.bert_dropout, tag)
    return _out_ln(sd, pre + "output.", o, x, cfg.bert_eps, drops, cfg.bert_dropout, tag + ".out")


def _ffn_block(sd, inter, out, x, cfg, drops, tag):
    """BertIntermediate + BertOutput (vilmodel.py:292-295, 305-309)."""
    y = gelu_erf(F.linear(x, sd[inter + "dense.weight"], sd[inter + "dense.bias"]))
    return _out_ln(sd, out, y, x, cfg.bert_eps, drops, cfg.bert_dropout, tag)


def bert_layer(sd, pre, x, add_mask, cfg, drops, tag):
    """BertLayer (vilmodel.py:319-325)."""
    a = _self_att_block(sd, pre + "attention.", x, add_mask, cfg, drops, tag + ".att")
    return _ffn_block(sd, pre + "intermediate.", pre + "output.", a, cfg, drops, tag + ".ffn")


def lxrt_layer(sd, pre, lang, lang_mask, visn, visn_mask, cfg, drops, tag):
    """LXRTXLayer.forward (vilmodel.py:1031-1064): ONE shared visual_attention for both directions."""
    xa = pre + "visual_attention."

    def cross(x, ctx, m, t):
        o = _mha(sd, xa + "att.", ("query", "key", "value"), x, ctx, m, cfg.bert_heads, drops,
                 cfg.bert_dropout, t)
        return _out_ln(sd, xa + "output.", o, x, cfg.bert_eps, drops, cfg.bert_dropout, t + ".out")

    l1 = cross(lang, visn, visn_mask, tag + ".x_lv")
    v1 = cross(visn, lang, lang_mask, tag + ".x_vl")
    l2 = _self_att_block(sd, pre + "lang_self_att.", l1, lang_mask, cfg, drops, tag + ".ls")
    v2 = _self_att_block(sd, pre + "visn_self_att.", v1, visn_mask, cfg, drops, tag + ".vs")
    l3 = _ffn_block(sd, pre + "lang_inter.", pre + "lang_output.", l2, cfg, drops, tag + ".lo")
    v3 = _ffn_block(sd, pre + "visn_inter.", pre + "visn_output.", v2, cfg, drops, tag + ".vo")
    return l3, v3


def language_stack(sd, cfg, seq, mask, drops=NoDrop()):
    """BertEmbeddings + la_layers x BertLayer (vilmodel.py:161-176, 1366-1378). seq [B,L] ids, mask True=pad.
    Step-invariant and gradient-free (detach at vilmodel.py:1377-1378)."""
    B, L = seq.shape
    pos = torch.arange(L).unsqueeze(0).expand(B, L)
    e = sd["bert.embeddings.word_embeddings.weight"][seq] + sd["bert.embeddings.position_embeddings.weight"][pos] \
        + sd["bert.embeddings.token_type_embeddings.weight"][torch.zeros_like(seq)]
    x = F.layer_norm(e, (e.shape[-1],), sd["bert.embeddings.LayerNorm.weight"],
                     sd["bert.embeddings.LayerNorm.bias"], cfg.bert_eps)
    x = drops(x, cfg.bert_dropout, "enc.emb")
    add_mask = (mask.to(x.dtype) * -10000.0)[:, None, None, :]          # vilmodel.py:1339-1347
    for i in range(cfg.la_layers):
        x = bert_layer(sd, "bert.lalayer.%d." % i, x, add_mask, cfg, drops, "enc.la%d" % i)
    return x.detach(), add_mask


def vision_encoder(sd, cfg, feats, drops=NoDrop()):
    """VisionEncoder (vilmodel.py:1083-1095): dropout(LN(Linear(feats))), eps=1e-12."""
    x = F.linear(feats, sd["bert.vision_encoder.visn_fc.weight"], sd["bert.vision_encoder.visn_fc.bias"])
    x = F.layer_norm(x, (x.shape[-1],), sd["bert.vision_encoder.visn_layer_norm.weight"],
                     sd["bert.vision_encoder.visn_layer_norm.bias"], 1e-12)
    return drops(x, cfg.bert_dropout, "enc.visn")


def reverse_tokens(x, lengths):
    """rev[b,i] = x[b, len_b-1-i] for i < len_b else 0 (r2rmodel.py:2326-2330)."""
    B, L, _ = x.shape
    idx = lengths.view(B, 1).to(torch.int64) - 1 - torch.arange(L).view(1, L)
    valid = idx >= 0
    out = torch.gather(x, 1, idx.clamp(min=0).unsqueeze(-1).expand_as(x))
    return out * valid.unsqueeze(-1).to(x.dtype)


def bilstm(sd, x, lengths, He):
    """Packed 1-layer bidirectional nn.LSTM (r2rmodel.py:2339-2357). Returns ctx [B,L,2He] with exact zeros in
    padded rows, final (h, c) per direction."""
    B, L, _ = x.shape
    outs, finals = [], []
    for sfx, rev in (("", False), ("_reverse", True)):
        w_ih, w_hh = sd["lstm.weight_ih_l0" + sfx], sd["lstm.weight_hh_l0" + sfx]
        b_ih, b_hh = sd["lstm.bias_ih_l0" + sfx], sd["lstm.bias_hh_l0" + sfx]
        h = x.new_zeros(B, He)
        c = x.new_zeros(B, He)
        out = x.new_zeros(B, L, He)
        steps = range(L - 1, -1, -1) if rev else range(L)
        for l in steps:
            act = (l < lengths).view(B, 1).to(x.dtype)
            h1, c1 = lstm_cell(w_ih, w_hh, b_ih, b_hh, x[:, l], h, c)
            h = act * h1 + (1 - act) * h
            c = act * c1 + (1 - act) * c
            out[:, l] = act * h1
        outs.append(out)
        finals.append((h, c))
    return torch.cat(outs, -1), finals


def encoder_forward(sd, cfg, seq, mask, lengths, f_t_all, drops=NoDrop(), lang_cache=None):
    """DicEncoder.forward (r2rmodel.py:2272-2365) + DicModel.forward (vilmodel.py:1327-1423), train config
    (update_lang_bert=False; update_add_layer per cfg). Returns (ctx, decoder_init, c_t, vision_outputs)."""
    L = mask.shape[1]
    if lang_cache is None:
        lang, lang_mask = language_stack(sd, cfg, seq[:, :L], mask, drops)
    else:
        lang, lang_mask = lang_cache
    visn = vision_encoder(sd, cfg, f_t_all, drops)
    for i in range(cfg.vl_layers):
        lang, visn = lxrt_layer(sd, "bert.addlayer.%d." % i, lang, lang_mask, visn, None, cfg, drops, "enc.vl%d" % i)
    if not cfg.update_add_layer:
        lang, visn = lang.detach(), visn.detach()
    rev = reverse_tokens(lang, lengths)
    ctx, ((h_f, c_f), (h_b, c_b)) = bilstm(sd, rev, lengths, cfg.enc_hidden)
    # h_t = cat(enc_h_t[-1], enc_h_t[-2]) : reverse-direction final state first (r2rmodel.py:2345-2346)
    h_cat, c_cat = torch.cat((h_b, h_f), 1), torch.cat((c_b, c_f), 1)
    decoder_init = torch.tanh(F.linear(h_cat, sd["encoder_lstm2decoder_ht.weight"], sd["encoder_lstm2decoder_ht.bias"]))
    c_t = F.linear(c_cat, sd["encoder_lstm2decoder_ct.weight"], sd["encoder_lstm2decoder_ct.bias"])
    ctx = drops(ctx, cfg.enc_dropout, "enc.ctx")
    return ctx, decoder_init, c_t, visn


# ------------------------------------------------------------------------------- rollout (a10, a11), teacher
def length2mask(length, size):
    """utils.length2mask (utils.py:503-508): True where index >= length."""
    return torch.arange(size).unsqueeze(0) >= length.view(-1, 1).to(torch.int64)


class _Prefixed:
    """Per-nav-step view of a dropout source: prefixes tags with 't<step>.'."""

    def __init__(self, base, t):
        self.base, self.t, self.training = base, t, base.training

    def __call__(self, x, p, tag):
        return self.base(x, p, "t%d.%s" % (self.t, tag))


def policy_step(state, cfg, seq, mask, lengths, step_inputs, carry, drops=NoDrop(), adain="channel",
                lang_cache=None, noise=None):
    """One iteration of the vl_rollout loop body up to the masked logits (agent_dg.py:727-841).
    carry = None at t==0 (decoder starts from the encoder state, :812-815) else (h1, c_t).
    Note: the encoder sees the RAW f_t while the decoder sees the AdaIN'd copy (agent_dg.py:728,764-768,793)."""
    input_a_t, f_t, d_t, cand_feat, cand_dfeat, cand_leng = step_inputs[:6]
    C = cfg.rgb_size
    if adain == "channel":
        gate = lambda f, d: adain_channel_gate(state["adaIn"], f, d)
        df_rgb = gate(f_t[..., :C], d_t[..., :C])
        cand_rgb = gate(cand_feat[..., :C], cand_dfeat[..., :C])
    elif adain == "stat":   # depth_stat_channel (agent_dg.py:755-760): candidates use the VIEW depth stats
        df_rgb = adain_stat_channel(state["adaIn"], f_t[..., :C], d_t[..., :C])
        cand_rgb = adain_stat_channel(state["adaIn"], cand_feat[..., :C], d_t[..., :C])
    elif adain == "default":
        df_rgb = adain_default(f_t[..., :C], d_t[..., :C])
        cand_rgb = adain_default(cand_feat[..., :C], cand_dfeat[..., :C])
    else:
        raise ValueError(adain)
    df_t = torch.cat([df_rgb, f_t[..., C:]], -1)
    cand = torch.cat([cand_rgb, cand_feat[..., C:]], -1)
    if noise is not None:
        # consistent_drop with --env_drop_stage after_adain --depth_drop (agent_dg.py:780-785): one [C] mask, shared by the
        # batch, the views and all steps, multiplies the AdaIN'd candidates, the RAW f_t (the encoder's input) and the
        # AdaIN'd views; the decoder is then called with already_dropfeat=True (agent_dg.py:812-820)
        cand = torch.cat([cand[..., :C] * noise, cand[..., C:]], -1)
        f_t = torch.cat([f_t[..., :C] * noise, f_t[..., C:]], -1)
        df_t = torch.cat([df_t[..., :C] * noise, df_t[..., C:]], -1)
    ctx, en_h, en_c, _ = encoder_forward(state["encoder"], cfg, seq, mask, lengths, f_t, drops, lang_cache)
    prev_h1, c_0 = (en_h, en_c) if carry is None else carry
    h_t, c_t, logit, h1, aux = decoder_step(state["decoder"], cfg, input_a_t, df_t, cand, prev_h1, c_0, ctx, mask, drops,
                                            already_dropfeat=noise is not None)
    logit = logit.masked_fill(length2mask(cand_leng, logit.shape[1]), -float("inf"))
    aux["ctx"] = ctx
    return logit, h_t, (h1, c_t), aux


def teacher_rollout(state, cfg, episodes, T=None, ml_weight=0.4, drops=NoDrop(), adain="channel",
                    cache_language=False, noise=None):
    """Teacher-forced vl_rollout (agent_dg.py:725-936, feedback='teacher', train_rl=False) followed by the loss
    assembly (agent_dg.py:1006-1024): loss = sum_t CE_sum(logit_t, target_t, ignore -100) * ml_weight / B.
    Returns (loss, list of masked logits, list of argmax actions)."""
    T = episodes.T if T is None else T
    seq, mask, lengths = episodes.seq, episodes.seq_mask, episodes.seq_lengths
    lang_cache = language_stack(state["encoder"], cfg, seq[:, :mask.shape[1]], mask) if cache_language else None
    carry, total, logits, actions = None, 0.0, [], []
    for t in range(T):
        step = episodes.step(t)
        sdrops = _Prefixed(drops, t) if drops.training else drops
        logit, h_t, carry, _ = policy_step(state, cfg, seq, mask, lengths, step, carry, sdrops, adain, lang_cache, noise)
        total = total + F.cross_entropy(logit, step[6], ignore_index=cfg.ignore_id, reduction="sum")
        logits.append(logit)
        actions.append(logit.argmax(1))
    loss = total * ml_weight / episodes.B
    return loss, logits, actions


# ------------------------------------------------------------------------------- sampled feedback + A2C (a10, a11)
class _Tagged:
    """Dropout source view that prefixes every tag (e.g. 'last.' for the extra decoder / critic call of the A2C epilogue)."""

    def __init__(self, base, prefix):
        self.base, self.prefix, self.training = base, prefix, base.training

    def __call__(self, x, p, tag):
        return self.base(x, p, self.prefix + tag)


def nav_reward(a_t, cand_leng, ignore_id, dist, last_dist, ended):
    """The reward / mask / ended bookkeeping after one action (agent_dg.py:890-930), loop for loop. `dist` is the distance
    to the goal after the action. Returns (reward [B] f32, mask [B] f32, ended [B] bool). The reference raises when a
    non-END action leaves the distance unchanged; here that case yields reward 0 (the synthetic walks never produce it)."""
    B = a_t.shape[0]
    cpu_a_t = a_t.clone()
    for i in range(B):
        if int(cpu_a_t[i]) == int(cand_leng[i]) - 1 or int(cpu_a_t[i]) == ignore_id:
            cpu_a_t[i] = -1
    reward, mask = torch.zeros(B), torch.ones(B)
    for i in range(B):
        if bool(ended[i]):
            reward[i], mask[i] = 0.0, 0.0
        elif int(cpu_a_t[i]) == -1:
            reward[i] = 2.0 if float(dist[i]) < 3 else -2.0
        else:
            r = -(float(dist[i]) - float(last_dist[i]))
            reward[i] = 1.0 if r > 0 else (-1.0 if r < 0 else 0.0)
    return reward, mask, ended | (cpu_a_t == -1)


def a2c_epilogue(policy_log_probs, entropys, values, last_value, rewards, masks, ended, gamma=0.9, ent_coef=0.01,
                 normalize="total"):
    """agent_dg.py:959-999 for lists of T per-step [B] tensors. values[t] = critic(hidden_states[t]); last_value is detached.
    Returns (rl_loss, total)."""
    B = last_value.shape[0]
    discount_reward = torch.where(ended, torch.zeros(B), last_value.detach().float())
    rl_loss, total = 0.0, 0.0
    for t in range(len(rewards) - 1, -1, -1):
        discount_reward = discount_reward * gamma + rewards[t]
        mask_, r_, v_ = masks[t], discount_reward.clone(), values[t]
        a_ = (r_ - v_).detach()
        rl_loss = rl_loss + (-policy_log_probs[t] * a_ * mask_).sum()
        rl_loss = rl_loss + (((r_ - v_) ** 2) * mask_).sum() * 0.5
        if entropys is not None:
            rl_loss = rl_loss + (-ent_coef * entropys[t] * mask_).sum()
        total = total + float(masks[t].sum())
    if normalize == "total":
        rl_loss = rl_loss / total
    elif normalize == "batch":
        rl_loss = rl_loss / B
    return rl_loss, total


def sample_rollout(state, cfg, episodes, T, actions, drops=NoDrop(), gamma=0.9, ent_coef=0.01, normalize="total",
                   adain="channel"):
    """vl_rollout with feedback='sample', train_rl=True, train_ml=None (agent_dg.py:725-999) over a pre-generated observation
    stream (episodes must hold T+1 observations and `dist` [T+1,B]); `actions[t]` are the sampled actions (injected, so that
    the CUDA path and this oracle follow the same trajectory). Returns (loss, dict of intermediates)."""
    seq, mask, lengths = episodes.seq, episodes.seq_mask, episodes.seq_lengths
    B = episodes.B
    ended = torch.zeros(B, dtype=torch.bool)
    last_dist = episodes.dist[0]
    carry, rewards, masks, hidden, logps, ents, logits = None, [], [], [], [], [], []
    for t in range(T):
        step = episodes.step(t)
        sdrops = _Prefixed(drops, t) if drops.training else drops
        logit, h_t, carry, aux = policy_step(state, cfg, seq, mask, lengths, step, carry, sdrops, adain)
        hidden.append(h_t)
        logits.append(logit)
        c = torch.distributions.Categorical(F.softmax(logit, 1))            # agent_dg.py:877-883
        ents.append(c.entropy())
        a_t = actions[t]
        logps.append(c.log_prob(a_t))
        dist = episodes.dist[t + 1]
        r, m, ended = nav_reward(a_t, step[5], cfg.ignore_id, dist, last_dist, ended)
        rewards.append(r)
        masks.append(m)
        last_dist = dist
    # last action in A2C (agent_dg.py:945-957): RAW features of the next observation, decoder applies its own drop_env
    nxt = episodes.step(T)
    ldrops = _Tagged(drops, "last.") if drops.training else drops
    h1, c_t = carry
    last_h, _, _, _, _ = decoder_step(state["decoder"], cfg, nxt[0], nxt[1], nxt[3], h1, c_t, aux["ctx"], mask, ldrops,
                                      already_dropfeat=False)
    last_value = critic(state["critic"], last_h, ldrops, cfg.dropout).detach()
    values = [critic(state["critic"], hidden[t], _Prefixed(drops, t) if drops.training else drops, cfg.dropout)
              for t in range(T)]
    loss, total = a2c_epilogue(logps, ents, values, last_value, rewards, masks, ended, gamma, ent_coef, normalize)
    return loss, {"logits": logits, "logps": logps, "ents": ents, "values": values, "last_value": last_value,
                  "rewards": rewards, "masks": masks, "ended": ended, "total": total}


# ------------------------------------------------------------------------------------ speaker inference
def _lstm_dir(sd, pre, sfx, x, h0=None, c0=None, reverse=False):
    """One direction of a one-layer batch_first nn.LSTM over full-length sequences, step by step with lstm_cell."""
    B, L, _ = x.shape
    H = sd[pre + "weight_hh_l0" + sfx].shape[1]
    h = torch.zeros(B, H) if h0 is None else h0
    c = torch.zeros(B, H) if c0 is None else c0
    out = [None] * L
    for l in (range(L - 1, -1, -1) if reverse else range(L)):
        h, c = lstm_cell(sd[pre + "weight_ih_l0" + sfx], sd[pre + "weight_hh_l0" + sfx], sd[pre + "bias_ih_l0" + sfx],
                         sd[pre + "bias_hh_l0" + sfx], x[:, l], h, c)
        out[l] = h
    return torch.stack(out, 1), h, c


def speaker_encoder(sd, action_embeds, feature):
    """model.SpeakerEncoder.forward in eval mode (model.py:1005-1036): bi-LSTM over the taken-candidate features, soft-dot
    attention of every step's state over that step's 36 views, post bi-LSTM. action_embeds [B, L, F], feature [B, L, 36, F]."""
    B, L, Fd = action_embeds.shape
    ctx = torch.cat((_lstm_dir(sd, "lstm.", "", action_embeds)[0], _lstm_dir(sd, "lstm.", "_reverse", action_embeds, reverse=True)[0]), 2)
    hidden = ctx.shape[2]
    x, _ = soft_dot_attention(sd, "attention_layer.", ctx.reshape(B * L, hidden), feature.reshape(B * L, -1, Fd))
    x = x.view(B, L, hidden)
    return torch.cat((_lstm_dir(sd, "post_lstm.", "", x)[0], _lstm_dir(sd, "post_lstm.", "_reverse", x, reverse=True)[0]), 2)


def speaker_decoder_step(sd, word, ctx, ctx_mask, h0, c0):
    """model.SpeakerDecoder.forward for one word per sequence in eval mode (model.py:1054-1078): word [B] int64,
    h0 / c0 [B, H] -> (logit [B, V], h1, c1)."""
    embeds = sd["embedding.weight"][word]
    h1, c1 = lstm_cell(sd["lstm.weight_ih_l0"], sd["lstm.weight_hh_l0"], sd["lstm.bias_ih_l0"], sd["lstm.bias_hh_l0"], embeds, h0, c0)
    x, _ = soft_dot_attention(sd, "attention_layer.", h1, ctx, ctx_mask)
    return F.linear(x, sd["projection.weight"], sd["projection.bias"]), h1, c1


def speaker_infer_greedy(enc_sd, dec_sd, can_feats, img_feats, lengths, bos, eos, pad, unk, max_decode=120):
    """Speaker.infer_batch with sampling=False (speaker.py:265-350) after from_shortest_path(): returns (words [B, n] int64,
    ctx, list of per-step logits)."""
    ctx = speaker_encoder(enc_sd, can_feats, img_feats)
    B = ctx.shape[0]
    ctx_mask = length2mask(torch.as_tensor(lengths), ctx.shape[1])
    H = dec_sd["lstm.weight_hh_l0"].shape[1]
    h_t, c_t = torch.zeros(B, H), torch.zeros(B, H)
    ended = torch.zeros(B, dtype=torch.bool)
    word = torch.full((B,), bos, dtype=torch.int64)
    words, all_logits = [], []
    for i in range(max_decode):
        logits, h_t, c_t = speaker_decoder_step(dec_sd, word, ctx, ctx_mask, h_t, c_t)
        logits = logits.clone()
        logits[:, unk] = -float("inf")
        all_logits.append(logits)
        _, word = logits.max(1)
        cpu_word = word.clone()
        cpu_word[ended] = pad
        words.append(cpu_word)
        ended = ended | (cpu_word == eos)
        if bool(ended.all()):
            break
    return torch.stack(words, 1), ctx, all_logits
