"""Deterministic synthetic R2R-shaped inputs and seeded weights.

There is no simulator, dataset or checkpoint offline, so both the CUDA path and the CPU oracle are fed from
this generator: weights carry the reference's state_dict keys/shapes so the same dict loads
into the reference modules (oracle/make_golden.py), the oracle restatement and the drop-in modules.
Everything is generated on the CPU with torch.Generator so CPU and GPU sides see identical bits.

Layout facts mirrored from the reference:
  * a view row is [2048 RGB | 128 angle]; angle = [sin h, cos h, sin e, cos e] x 32 (utils.py:361-368),
    h = (v % 12)*30deg - agent heading, e = (v // 12 - 1)*30deg (utils.py:386-405, env.py:332);
  * a candidate row is the view row at its pointId with the candidate-relative angle feature
    (env.py:273-288); the END row is all zeros and sits at index len(candidates) (agent_dg.py:300-311);
  * instruction ids: [CLS]=101 ... [SEP]=102, pad 0, batch sorted by length, longest first
    (utils.py:604-616, agent_dg.py:262-281); mask True = padding.
"""
import math
from collections import OrderedDict

import torch

from .config import PolicyConfig, FULL


def _gen(seed):
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    return g


# ----------------------------------------------------------------------------------------------- weights
def _lin(sd, g, name, out_f, in_f, bias=True, gain=1.0):
    sd[name + ".weight"] = torch.randn(out_f, in_f, generator=g) * (gain / math.sqrt(in_f))
    if bias:
        sd[name + ".bias"] = torch.randn(out_f, generator=g) * 0.02


def _ln(sd, g, name, n):
    sd[name + ".weight"] = 1.0 + 0.05 * torch.randn(n, generator=g)
    sd[name + ".bias"] = 0.02 * torch.randn(n, generator=g)


def adain_state(cfg: PolicyConfig = FULL, seed=0, kind="channel"):
    """DGAdaChannel (ab_type=a): a_fc (agent_dg.py:1516-1523); Stat: a_fc/b_fc over 4C (:1642-1646);
    Mean: a_fc/b_fc over C (:1623-1627)."""
    g, sd, C = _gen(1000 + seed), OrderedDict(), cfg.rgb_size
    if kind == "channel":
        _lin(sd, g, "a_fc", C, C)
    elif kind == "stat":
        _lin(sd, g, "a_fc", C, 4 * C)
        _lin(sd, g, "b_fc", C, 4 * C)
    elif kind == "mean":
        _lin(sd, g, "a_fc", C, C)
        _lin(sd, g, "b_fc", C, C)
    else:
        raise ValueError(kind)
    return sd


def decoder_state(cfg: PolicyConfig = FULL, seed=0):
    """BAttnDecoderLSTM parameters (model.py:425-443)."""
    g, sd = _gen(2000 + seed), OrderedDict()
    H, F, E, A, k = cfg.hidden, cfg.feat, cfg.action_emb, cfg.angle_size, cfg.shift_kernel
    _lin(sd, g, "embedding.0", E, A)
    sd["lstm.weight_ih"] = torch.randn(4 * H, E + F, generator=g) / math.sqrt(E + F)
    sd["lstm.weight_hh"] = torch.randn(4 * H, H, generator=g) / math.sqrt(H)
    sd["lstm.bias_ih"] = torch.randn(4 * H, generator=g) * 0.02
    sd["lstm.bias_hh"] = torch.randn(4 * H, generator=g) * 0.02
    _lin(sd, g, "feat_att_layer.linear_in", F, H, bias=False, gain=0.5)
    _lin(sd, g, "feat_att_layer.linear_shift", k, H, bias=True, gain=2.0)
    _lin(sd, g, "feat_att_layer.linear_out", H, H + F, bias=False)      # never used (model.py:511)
    _lin(sd, g, "attention_layer.linear_in", 2 * cfg.enc_hidden, H, bias=False, gain=0.5)
    _lin(sd, g, "attention_layer.linear_out", H, H + 2 * cfg.enc_hidden, bias=False)
    _lin(sd, g, "candidate_att_layer.linear_in", F, H, bias=False, gain=0.5)
    _lin(sd, g, "candidate_att_layer.linear_out", H, H + F, bias=False)  # dead work (model.py:285-294)
    return sd


def critic_state(cfg: PolicyConfig = FULL, seed=0):
    g, sd, D = _gen(3000 + seed), OrderedDict(), cfg.critic_dim
    _lin(sd, g, "state2value.0", D, D)
    _lin(sd, g, "state2value.3", 1, D)
    return sd


def _bert_attention(sd, g, p, Hb, self_style=True):
    a = p + (".self" if self_style else ".att")
    for n in ("query", "key", "value"):
        _lin(sd, g, a + "." + n, Hb, Hb, gain=1.5)
    _lin(sd, g, p + ".output.dense", Hb, Hb)
    _ln(sd, g, p + ".output.LayerNorm", Hb)


def _bert_ffn(sd, g, inter, out, Hb, I):
    _lin(sd, g, inter + ".dense", I, Hb)
    _lin(sd, g, out + ".dense", Hb, I)
    _ln(sd, g, out + ".LayerNorm", Hb)


def encoder_state(cfg: PolicyConfig = FULL, seed=0):
    """DicEncoder parameters (r2rmodel.py:2204-2250; vilmodel.py:1277-1296). Key order follows module
    registration order of the reference so state_dict() listings line up."""
    g, sd = _gen(4000 + seed), OrderedDict()
    Hb, I, He = cfg.bert_hidden, cfg.bert_inter, cfg.enc_hidden
    sd["bert.embeddings.word_embeddings.weight"] = torch.randn(cfg.vocab, Hb, generator=g) * 0.1
    sd["bert.embeddings.word_embeddings.weight"][0].zero_()          # padding_idx=0 (vilmodel.py:151)
    sd["bert.embeddings.position_embeddings.weight"] = torch.randn(cfg.max_pos, Hb, generator=g) * 0.05
    sd["bert.embeddings.token_type_embeddings.weight"] = torch.randn(cfg.type_vocab, Hb, generator=g) * 0.05
    _ln(sd, g, "bert.embeddings.LayerNorm", Hb)
    _lin(sd, g, "bert.pooler.dense", Hb, Hb)
    for i in range(cfg.la_layers):
        p = "bert.lalayer.%d" % i
        _bert_attention(sd, g, p + ".attention", Hb)
        _bert_ffn(sd, g, p + ".intermediate", p + ".output", Hb, I)
    for i in range(cfg.vl_layers):
        p = "bert.addlayer.%d" % i
        _bert_attention(sd, g, p + ".lang_self_att", Hb)
        _bert_ffn(sd, g, p + ".lang_inter", p + ".lang_output", Hb, I)
        _bert_attention(sd, g, p + ".visn_self_att", Hb)
        _bert_ffn(sd, g, p + ".visn_inter", p + ".visn_output", Hb, I)
        _bert_attention(sd, g, p + ".visual_attention", Hb, self_style=False)
    _lin(sd, g, "bert.vision_encoder.visn_fc", Hb, cfg.feat)
    _ln(sd, g, "bert.vision_encoder.visn_layer_norm", Hb)
    for sfx in ("", "_reverse"):
        sd["lstm.weight_ih_l0" + sfx] = torch.randn(4 * He, Hb, generator=g) / math.sqrt(Hb)
        sd["lstm.weight_hh_l0" + sfx] = torch.randn(4 * He, He, generator=g) / math.sqrt(He)
        sd["lstm.bias_ih_l0" + sfx] = torch.randn(4 * He, generator=g) * 0.02
        sd["lstm.bias_hh_l0" + sfx] = torch.randn(4 * He, generator=g) * 0.02
    _lin(sd, g, "encoder2decoder_ht", cfg.hidden, 2 * He)          # unused when top_lstm (r2rmodel.py:2333)
    _lin(sd, g, "encoder2decoder_ct", cfg.hidden, 2 * He)
    _lin(sd, g, "encoder_lstm2decoder_ht", cfg.hidden, 2 * He)
    _lin(sd, g, "encoder_lstm2decoder_ct", cfg.hidden, 2 * He)
    return sd


def speaker_state(feat, rnn_dim=512, wemb=256, vocab=992, seed=0):
    """SpeakerEncoder / SpeakerDecoder parameters (model.py:984-1052) -> (encoder state_dict, decoder state_dict)."""
    g, enc, dec = _gen(7000 + seed), OrderedDict(), OrderedDict()
    Hh = rnn_dim // 2

    def lstm(sd, pre, n_in, H, sfxs):
        for sfx in sfxs:
            sd[pre + "weight_ih_l0" + sfx] = torch.randn(4 * H, n_in, generator=g) / math.sqrt(n_in)
            sd[pre + "weight_hh_l0" + sfx] = torch.randn(4 * H, H, generator=g) / math.sqrt(H)
            sd[pre + "bias_ih_l0" + sfx] = torch.randn(4 * H, generator=g) * 0.02
            sd[pre + "bias_hh_l0" + sfx] = torch.randn(4 * H, generator=g) * 0.02
    lstm(enc, "lstm.", feat, Hh, ("", "_reverse"))
    _lin(enc, g, "attention_layer.linear_in", feat, rnn_dim, bias=False, gain=0.5)
    _lin(enc, g, "attention_layer.linear_out", rnn_dim, rnn_dim + feat, bias=False)
    lstm(enc, "post_lstm.", rnn_dim, Hh, ("", "_reverse"))
    dec["embedding.weight"] = torch.randn(vocab, wemb, generator=g) * 0.3
    dec["embedding.weight"][0].zero_()                      # padding_idx = <PAD> = 0
    lstm(dec, "lstm.", wemb, rnn_dim, ("",))
    _lin(dec, g, "attention_layer.linear_in", rnn_dim, rnn_dim, bias=False, gain=1.0)
    _lin(dec, g, "attention_layer.linear_out", rnn_dim, 2 * rnn_dim, bias=False)
    _lin(dec, g, "projection", vocab, rnn_dim, gain=3.0)
    _lin(dec, g, "baseline_projection.0", 128, rnn_dim)
    _lin(dec, g, "baseline_projection.3", 1, 128)
    return enc, dec


def speaker_inputs(B, L, cfg: PolicyConfig = FULL, seed=0):
    """Shortest-path features as Speaker.from_shortest_path returns them (speaker.py:150-190): can_feats [B, L, F] (the view
    row of the taken candidate), img_feats [B, L, 36, F], path lengths [B] (sorted like the instruction batch is not required)."""
    g = _gen(8000 + seed)
    C, A, V, F = cfg.rgb_size, cfg.angle_size, cfg.views, cfg.feat
    img = torch.empty(B, L, V, F)
    img[..., :C] = resnet_like((B, L, V, C), g)
    base = torch.randint(0, V, (B * L,), generator=g)
    img[..., C:] = view_angle_features(base, cfg).view(B, L, V, A)
    pick = torch.randint(0, V, (B, L), generator=g)
    can = torch.gather(img, 2, pick[:, :, None, None].expand(B, L, 1, F)).squeeze(2).clone()
    lengths = torch.randint(max(2, L // 2), L + 1, (B,), generator=g)
    lengths[0] = L
    for b in range(B):                                       # padded steps are zeros (speaker.py:181-186)
        can[b, int(lengths[b]):] = 0
        img[b, int(lengths[b]):] = 0
    return can, img, lengths


def policy_state(cfg: PolicyConfig = FULL, seed=0, adain_kind="channel"):
    return {"adaIn": adain_state(cfg, seed, adain_kind), "decoder": decoder_state(cfg, seed),
            "critic": critic_state(cfg, seed), "encoder": encoder_state(cfg, seed)}


# ------------------------------------------------------------------------------------------------ inputs
def angle_feature(heading, elevation, size):
    """utils.angle_feature (utils.py:361-368) for tensors of headings/elevations -> [..., size]."""
    base = torch.stack([torch.sin(heading), torch.cos(heading), torch.sin(elevation), torch.cos(elevation)], -1)
    return base.repeat(*([1] * heading.dim()), size // 4).float()


def view_angle_features(base_view, cfg: PolicyConfig = FULL):
    """[B] agent view index -> [B, 36, A] relative angle features (utils.get_point_angle_feature)."""
    v = torch.arange(cfg.views)
    head = (v % cfg.headings).double() * math.radians(30)
    elev = ((v // cfg.headings) - 1).double() * math.radians(30)
    base_heading = (base_view % cfg.headings).double() * math.radians(30)
    h = head[None, :] - base_heading[:, None]
    return angle_feature(h, elev[None, :].expand_as(h), cfg.angle_size)


def resnet_like(shape, g):
    """ResNet-152 pool5-like activations: non-negative, about half zeros."""
    return torch.relu(torch.randn(*shape, generator=g)) * 0.5


def instructions(B, cfg: PolicyConfig = FULL, seed=0, full_length=False):
    """-> seq [B, max_input] int64, mask [B, Lmax] bool (True = pad), lengths [B] int64 sorted desc."""
    g = _gen(5000 + seed)
    Lm = cfg.max_input
    if full_length:
        lens = torch.full((B,), Lm, dtype=torch.int64)
    else:
        lens = torch.clamp((torch.randn(B, generator=g) * 11 + 29).round().long(), min(8, Lm), Lm)
        if Lm < 40:
            lens = torch.randint(max(4, Lm // 3), Lm + 1, (B,), generator=g)
    lens, _ = lens.sort(descending=True)
    lo = min(1000, cfg.vocab // 2)
    seq = torch.zeros(B, Lm, dtype=torch.int64)
    for b in range(B):
        n = int(lens[b])
        seq[b, :n] = torch.randint(lo, cfg.vocab, (n,), generator=g)
        seq[b, 0], seq[b, n - 1] = min(101, cfg.vocab - 2), min(102, cfg.vocab - 1)
    Lmax = int(lens[0])
    mask = (seq == 0)[:, :Lmax]
    return seq, mask, lens


class Episodes:
    """A batch of teacher-forced synthetic episodes, fully materialised per step (host tensors).

    Per step t the fields match get_input_feat / _teacher_action (agent_dg.py:313-344):
      input_a_t [T,B,A], f_t/d_t [T,B,36,F], cand_feat/cand_dfeat [T,B,Nc,F], cand_leng [T,B] (incl. END),
      target [T,B] (teacher action, ignore_id once ended).
    """

    def __init__(self, B, T, cfg: PolicyConfig = FULL, seed=0, nc_max=14, bank=48, pin=False,
                 full_length=False, stress=False):
        g = _gen(6000 + seed)
        self.B, self.T, self.cfg, self.nc_max = B, T, cfg, nc_max
        C, A, V, F = cfg.rgb_size, cfg.angle_size, cfg.views, cfg.feat
        draw = (lambda s: torch.randn(*s, generator=g)) if stress else (lambda s: resnet_like(s, g))
        rgb_bank, dep_bank = draw((bank, V, C)), draw((bank, V, C))
        self.seq, self.seq_mask, self.seq_lengths = instructions(B, cfg, seed, full_length)
        # walk: viewpoint id per (t, b), agent view index, degree, episode length
        vp = torch.randint(0, bank, (T, B), generator=g)
        base_view = torch.randint(0, V, (T, B), generator=g)
        # candidate count without END: R2R degree distribution is mean ~4, max 13
        deg = torch.clamp((torch.randn(T, B, generator=g).abs() * 3.0 + 1.5).round().long(), 1, nc_max - 1)
        ep_len = torch.randint(max(1, T // 2), T + 1, (B,), generator=g)
        ep_len[torch.randint(0, B, (1,), generator=g)] = T            # no early exit of the whole batch
        self.ep_len = ep_len
        ncm = int(deg.max()) + 1
        self.cand_leng = (deg + 1).int()
        self.input_a_t = torch.empty(T, B, A)
        self.f_t = torch.empty(T, B, V, F)
        self.d_t = torch.empty(T, B, V, F)
        self.cand_feat = torch.zeros(T, B, ncm, F)
        self.cand_dfeat = torch.zeros(T, B, ncm, F)
        self.target = torch.empty(T, B, dtype=torch.int64)
        for t in range(T):
            ang = view_angle_features(base_view[t], cfg)                       # [B,36,A]
            self.f_t[t, :, :, :C], self.f_t[t, :, :, C:] = rgb_bank[vp[t]], ang
            self.d_t[t, :, :, :C], self.d_t[t, :, :, C:] = dep_bank[vp[t]], ang
            head = (base_view[t] % cfg.headings).double() * math.radians(30)
            elev = ((base_view[t] // cfg.headings) - 1).double() * math.radians(30)
            self.input_a_t[t] = angle_feature(head, elev, A)
            for b in range(B):
                n = int(deg[t, b])
                pts = torch.randint(0, V, (n,), generator=g)
                rel_h = (torch.rand(n, generator=g).double() - 0.5) * math.radians(30) + \
                        (pts % cfg.headings).double() * math.radians(30) - head[b]
                rel_e = (torch.rand(n, generator=g).double() - 0.5) * math.radians(20) + \
                        ((pts // cfg.headings) - 1).double() * math.radians(30)
                cang = angle_feature(rel_h, rel_e, A)
                self.cand_feat[t, b, :n, :C], self.cand_feat[t, b, :n, C:] = rgb_bank[vp[t, b], pts], cang
                self.cand_dfeat[t, b, :n, :C], self.cand_dfeat[t, b, :n, C:] = dep_bank[vp[t, b], pts], cang
                if t >= int(ep_len[b]):
                    self.target[t, b] = cfg.ignore_id
                elif t == int(ep_len[b]) - 1:
                    self.target[t, b] = n                                       # STOP = END row
                else:
                    self.target[t, b] = int(torch.randint(0, n, (1,), generator=g))
        # distance to the goal before action t (dist[t]) and after it (dist[t+1]) for the sampled-feedback reward
        # (agent_dg.py:906-925): a random walk that never stands still; drawn last so earlier fields keep their bits
        steps = (torch.rand(T, B, generator=g) * 2.0 + 0.5) * torch.where(torch.rand(T, B, generator=g) < 0.6, -1.0, 1.0)
        self.dist = torch.cat([torch.rand(1, B, generator=g) * 10.0 + 4.0, steps], 0).cumsum(0).abs().float()
        if pin and torch.cuda.is_available():
            for k in ("input_a_t", "f_t", "d_t", "cand_feat", "cand_dfeat", "cand_leng", "target"):
                setattr(self, k, getattr(self, k).pin_memory())

    def step(self, t):
        return (self.input_a_t[t], self.f_t[t], self.d_t[t], self.cand_feat[t], self.cand_dfeat[t],
                self.cand_leng[t], self.target[t])

    def h2d_bytes_per_step(self):
        n = 0
        for k in ("input_a_t", "f_t", "d_t", "cand_feat", "cand_dfeat", "cand_leng", "target"):
            x = getattr(self, k)
            n += x[0].numel() * x.element_size()
        return n


def dropout_mask(shape, p, g):
    """Keep-mask scaled by 1/(1-p) (what nn.Dropout multiplies by in train mode)."""
    return (torch.rand(*shape, generator=g) >= p).float() / (1.0 - p)
