"""Drop-in torch.nn modules for the agent_dg hot path: same constructor arguments, forward signatures, return values
and state_dict keys/shapes as the reference, computed by the sm_100a kernels of libdasa_b200.

Reference classes mirrored here (paths relative to /root/reference/r2r_src):
  DGAdaChannel / DGAdaStatChannel / DGAdaMeanChannel   agent_dg.py:1513-1661
  adaptive_instance_normalization                      model.py:1822-1840
  SoftDotAttention / ShiftSoftDotAttention             model.py:253-353
  BAttnDecoderLSTM                                     model.py:422-574
  Critic                                               model.py:970-982
  DicEncoder (+ DicModel, LXRTXLayer, BertLayer ...)   r2rmodel.py:2199-2365, vilmodel.py:1245-1423

The reference reads a module-global `args`; here the same fields are explicit keyword arguments whose defaults are the
README "train" command's values (README.md:82-96). nn.Linear / nn.LayerNorm / nn.Embedding / nn.LSTM / nn.LSTMCell
objects are used ONLY as parameter containers (identical state_dict keys and default initialisation); their forward()
is never called. Modules raise if fed CPU tensors: there is no fallback path.

Dropout: the reference uses nn.Dropout with torch's Philox stream, which custom kernels cannot reproduce. Every dropout
site therefore draws an explicit keep mask from the ambient `DropoutSource` (see below): generated on the device by a
counter-based RNG kernel in normal training, or injected by tag for parity tests (same tags as oracle/restated.py).
"""
import contextlib
import math

import torch
import torch.nn as nn

from . import functions as Fn
from . import ops
from .config import FULL, PolicyConfig, flag


# ------------------------------------------------------------------------------------------------- dropout source
class DropoutSource:
    """Provides keep masks for dropout sites. tag -> uint8 keep mask (or None in eval)."""

    def __init__(self, seed=0, injected=None, prefix="", device_seed=False, device="cuda"):
        self.seed, self.injected, self.prefix, self.counter = int(seed), injected, prefix, 0
        self._pools = {}
        # device_seed: the RNG seed lives in device memory and is advanced by a kernel (`advance()`), so a captured CUDA graph
        # draws different masks on every replay
        self.seed_dev = torch.tensor([int(seed)], dtype=torch.int64, device=device) if device_seed else None

    def mask_steps(self, tag, shape, p, training, device, steps):
        """Keep mask for `steps` consecutive rollout steps stacked along dim 0 (shape = per-step shape). With injected masks
        the per-step tags 't<i>.<tag>' are concatenated, so batched and per-step evaluation see identical masks."""
        if not training or p <= 0.0:
            return None, 1.0
        if self.injected is not None:
            parts = [self.injected.get("%st%d.%s" % (self.prefix, i, tag)) for i in range(steps)]
            if any(m is None for m in parts):
                return None, 1.0
            m = ops.as_keep_mask(torch.cat([x.to(device) for x in parts], 0))
            return m, 1.0 / (1.0 - p)
        return self.mask(tag, (shape[0] * steps,) + tuple(shape[1:]), p, training, device)

    def advance(self):
        """Start a new iteration: fresh masks (device-seed mode) and recycled pools."""
        self._pools = {}
        if self.seed_dev is not None:
            self.counter = 0
            ops.bump_counter(self.seed_dev)

    def mask(self, tag, shape, p, training, device):
        if not training or p <= 0.0:
            return None, 1.0
        scale = 1.0 / (1.0 - p)
        if self.injected is not None:
            m = self.injected.get(self.prefix + tag)
            if m is None:
                return None, 1.0
            m = ops.as_keep_mask(m.to(device))
            assert tuple(m.shape) == tuple(shape), (tag, m.shape, shape)
            return m, scale
        n = 1
        for s in shape:
            n *= s
        # sub-allocate from a per-probability pool filled by ONE RNG launch (dozens of tiny launches per action otherwise)
        pool = self._pools.get(p)
        if pool is None or pool[1] + n > pool[0].numel():
            size = max(self.POOL_BYTES, n)
            if self.seed_dev is not None:
                buf = ops.dropout_mask_dev((size,), p, self.seed_dev, self.counter, device)
            else:
                buf = ops.dropout_mask((size,), p, self.seed, self.counter, device)
            self.counter += size
            pool = [buf, 0]
            self._pools[p] = pool
        off = pool[1]
        pool[1] = (off + n + 15) & ~15
        return pool[0][off:off + n].view(tuple(shape)), scale

    POOL_BYTES = 16 << 20

    def stream(self, nbytes, p, training):
        """In-place draws for a forward-only dropout site (nothing re-reads the mask): reserves `nbytes` stream bytes and returns
        (ops.DropStream, scale); (None, 1.0) in eval mode. Only without injected masks - tests that inject masks get tensors
        from mask() / mask_steps()."""
        assert self.injected is None
        if not training or p <= 0.0:
            return None, 1.0
        base = self.counter
        self.counter += (int(nbytes) + 3) // 4
        return ops.DropStream(self.seed_dev, self.seed, base, p), 1.0 / (1.0 - p)


_source = DropoutSource()


def dropout_source():
    return _source


@contextlib.contextmanager
def use_dropout_source(src):
    global _source
    old, _source = _source, src
    try:
        yield src
    finally:
        _source = old


def _drop(x, tag, p, training):
    m, scale = _source.mask(tag, x.shape, p, training, x.device)
    return Fn.dropout(x, m, scale)


# -------------------------------------------------------------------------------------------------- AdaIN family
class DGAdaChannel(nn.Module):
    """agent_dg.py:1513-1547 with ab_type in {a}, a_type='sigmoid' (the README configuration)."""

    def __init__(self, channel, eps=1e-6, ab_type=None, a_type=None):
        """DGAdaChannel(channel) reads args.ab_type / args.a_type (agent_dg.py:1518-1546)."""
        super().__init__()
        ab_type, a_type = flag("ab_type", ab_type, "a"), flag("a_type", a_type, "sigmoid")
        if ab_type != "a" or a_type != "sigmoid":
            raise NotImplementedError("only --ab_type a --a_type sigmoid is on the hot path (README.md:86)")
        self.a_fc = nn.Linear(channel, channel)
        self.eps, self.channel = eps, channel

    def forward(self, f_t, d_t):
        """f_t, d_t: [N, L, C] (strided slices of [N, L, C+A] buffers are read in place) -> [N, L, C]."""
        return Fn.AdaINGateFn.apply(f_t, d_t, self.a_fc.weight, self.a_fc.bias, None, 1.0, self.channel)

    def gate_features(self, feat, dfeat, drop_mask=None, drop_scale=1.0):
        """Fused form used by the rollout: feat/dfeat [N, L, C+A] -> AdaIN'd copy with the angle part carried over and
        the decoder's drop_env mask folded into the GEMM epilogue (agent_dg.py:764-768 + model.py:506-508)."""
        return Fn.AdaINGateFn.apply(feat, dfeat, self.a_fc.weight, self.a_fc.bias, drop_mask, drop_scale, self.channel)


def _view_stats_nograd(d_t):
    """The depth features are environment data in the agent (agent_dg.py:742-777: built from obs, never a graph output), so
    the statistics carry no gradient; asking for one is outside the path and fails loudly."""
    if d_t.requires_grad:
        raise NotImplementedError("gradients w.r.t. the depth features d_t are not on the agent_dg path (d_t is env data)")
    return ops.view_stats(d_t)


class DGAdaStatChannel(nn.Module):
    """agent_dg.py:1639-1661, forward + backward (a_fc / b_fc gradients through ChannelModulateFn)."""

    def __init__(self, channel, eps=1e-6):
        super().__init__()
        self.a_fc = nn.Linear(4 * channel, channel)
        self.b_fc = nn.Linear(4 * channel, channel)
        self.eps = eps

    def forward(self, f_t, d_t):
        stats = _view_stats_nograd(d_t)
        a = Fn.linear(stats, self.a_fc.weight, self.a_fc.bias)
        b = Fn.linear(stats, self.b_fc.weight, self.b_fc.bias)
        return Fn.ChannelModulateFn.apply(f_t, a, b)


class DGAdaMeanChannel(nn.Module):
    """agent_dg.py:1620-1636, forward + backward."""

    def __init__(self, channel, eps=1e-6):
        super().__init__()
        self.a_fc = nn.Linear(channel, channel)
        self.b_fc = nn.Linear(channel, channel)
        self.eps, self.channel = eps, channel

    def forward(self, f_t, d_t):
        mean = _view_stats_nograd(d_t)[:, :self.channel].contiguous()
        a = Fn.linear(mean, self.a_fc.weight, self.a_fc.bias)
        b = Fn.linear(mean, self.b_fc.weight, self.b_fc.bias)
        return Fn.ChannelModulateFn.apply(f_t, a, b)


def adaptive_instance_normalization(content_feat, style_feat):
    """model.py:1831-1840."""
    assert content_feat.size() == style_feat.size()
    return ops.adain_rows(content_feat, style_feat, 1e-5)


# ------------------------------------------------------------------------------------------------------ attention
class SoftDotAttention(nn.Module):
    """model.py:253-296."""

    def __init__(self, query_dim, ctx_dim):
        super().__init__()
        self.linear_in = nn.Linear(query_dim, ctx_dim, bias=False)
        self.linear_out = nn.Linear(query_dim + ctx_dim, query_dim, bias=False)

    def forward(self, h, context, mask=None, output_tilde=True, output_prob=True, cand_leng=None, rgb_channels=None):
        if output_tilde and output_prob:
            return Fn.SoftDotAttnFn.apply(h, context, mask, self.linear_in.weight, self.linear_out.weight)
        if not output_prob:
            # candidate_att_layer call (model.py:559): only the raw logits are consumed; the softmax / weighted sum /
            # linear_out tail of the reference is dead work and is skipped.
            leng = cand_leng
            if leng is None and mask is not None:
                leng = (~mask.bool()).sum(1).to(torch.int32)
            rgb = context.shape[2] if rgb_channels is None else rgb_channels
            logit = Fn.CandLogitsFn.apply(h, context, leng, self.linear_in.weight, rgb)
            return None, logit
        raise NotImplementedError("SoftDotAttention(output_tilde=False, output_prob=True) is not used by agent_dg")


class ShiftSoftDotAttention(nn.Module):
    """model.py:300-353 (used with mask=None, output_tilde=False: model.py:511)."""

    def __init__(self, query_dim, ctx_dim, kernel_size=3, headings=12):
        super().__init__()
        self.linear_in = nn.Linear(query_dim, ctx_dim, bias=False)
        self.linear_shift = nn.Linear(query_dim, kernel_size)
        self.linear_out = nn.Linear(query_dim + ctx_dim, query_dim, bias=False)   # unused by the decoder; kept for state_dict
        self.kernel_size, self.padding_size, self.headings = kernel_size, kernel_size // 2, headings

    def forward(self, h, context, mask=None, output_tilde=True, output_prob=True):
        if mask is not None or output_tilde or not output_prob:
            raise NotImplementedError("agent_dg calls feat_att_layer(h, feature, output_tilde=False) only (model.py:511)")
        return Fn.ShiftAttnFn.apply(h, context, self.linear_in.weight, self.linear_shift.weight, self.linear_shift.bias,
                                    self.headings)


# -------------------------------------------------------------------------------------------------------- decoder
class BAttnDecoderLSTM(nn.Module):
    """model.py:422-574."""

    def __init__(self, embedding_size, hidden_size, dropout_ratio, feature_size=2048 + 4, pred_back=False,
                 angle_feat_size=None, featdropout=None, use_shift=None, shift_kernel_size=None):
        """The first five arguments are the reference's (model.py:425). The rest are the flags it reads from the global `args`
        (model.py:432-439): given explicitly, or taken from param.args when r2r_src's param module is loaded, or the README
        values (--angle_feat_size 128 --featdropout 0.4 --use_shift --shift_kernel_size 5)."""
        super().__init__()
        angle_feat_size = flag("angle_feat_size", angle_feat_size, 128)
        featdropout = flag("featdropout", featdropout, 0.4)
        use_shift = flag("use_shift", use_shift, True)
        shift_kernel_size = flag("shift_kernel_size", shift_kernel_size, 5)
        if pred_back:
            raise NotImplementedError("--pred_back is not part of the agent_dg README configuration")
        self.embedding_size, self.feature_size, self.hidden_size = embedding_size, feature_size, hidden_size
        self.angle_feat_size, self.dropout_ratio, self.featdropout = angle_feat_size, dropout_ratio, featdropout
        self.embedding = nn.Sequential(nn.Linear(angle_feat_size, embedding_size), nn.Tanh())
        self.drop = nn.Dropout(p=dropout_ratio)          # attribute kept: the agent pokes decoder.drop_env (agent_dg.py:657)
        self.drop_env = nn.Dropout(p=featdropout)
        self.lstm = nn.LSTMCell(embedding_size + feature_size, hidden_size)
        if use_shift:
            self.feat_att_layer = ShiftSoftDotAttention(hidden_size, feature_size, shift_kernel_size)
        else:
            raise NotImplementedError("--use_shift is part of the hot-path configuration")
        self.attention_layer = SoftDotAttention(hidden_size, hidden_size * 2)
        self.candidate_att_layer = SoftDotAttention(hidden_size, feature_size)
        self.pred_back = pred_back

    def embed_actions(self, actions_all, steps):
        """The action embeddings of `steps` actions stacked along dim 0 ([steps*B, A] -> [steps*B, E]) in one call: they do
        not depend on the recurrent state (model.py:504-505), so a teacher-forced rollout hoists them out of its loop.
        Dropout masks follow the per-step tags 't<i>.dec.act'."""
        p, tr = self.dropout_ratio, self.training
        emb = Fn.linear(actions_all, self.embedding[0].weight, self.embedding[0].bias, "tanh")
        m, s = _source.mask_steps("dec.act", (actions_all.shape[0] // steps, emb.shape[1]), p, tr, emb.device, steps)
        return Fn.dropout(emb, m, s)

    def candidate_logits_steps(self, h_tilde_all, cand_all, cand_leng_all, steps):
        """drop(h_tilde) -> candidate_att_layer for `steps` actions at once (model.py:555-559): the logits feed the loss only,
        never the recurrence, so a teacher-forced rollout evaluates them after its loop as ONE projection over steps*B rows
        instead of `steps` weight-streaming ones. Dropout masks follow the per-step tags 't<i>.dec.htilde'."""
        p, tr = self.dropout_ratio, self.training
        m, s = _source.mask_steps("dec.htilde", (h_tilde_all.shape[0] // steps, h_tilde_all.shape[1]), p, tr,
                                  h_tilde_all.device, steps)
        _, logit = self.candidate_att_layer(Fn.dropout(h_tilde_all, m, s), cand_all, output_prob=False,
                                            cand_leng=cand_leng_all, rgb_channels=self.feature_size - self.angle_feat_size)
        return logit

    # ------------------------------------------------------------------ persistent whole-rollout kernel (decoder_persist.cu)
    def _persistent_geometry(self, B, V, L, D):
        H, E, F = self.hidden_size, self.embedding_size, self.feature_size
        k = self.feat_att_layer.kernel_size
        NK = (F + k + 31) // 32 * 32
        return (B, H, E, F, V, L, D, NK, k)

    def _rollout_fn(self, emb, feature, ctx, ctx_mask, h0, c0, m_hprev, m_h1, scale):
        fa, al, l = self.feat_att_layer, self.attention_layer, self.lstm
        if ctx_mask is not None and ctx_mask.dtype != torch.uint8:
            ctx_mask = ctx_mask.to(torch.uint8)
        return Fn.DecoderRolloutFn.apply(emb, feature, ctx, ctx_mask, h0, c0, m_hprev, m_h1, scale, fa.linear_in.weight,
                                         fa.linear_shift.weight, fa.linear_shift.bias, l.weight_ih, l.weight_hh, l.bias_ih,
                                         l.bias_hh, al.linear_in.weight, al.linear_out.weight, fa.headings)

    def rollout_steps(self, emb_all, feat_all, ctx_all, ctx_mask, h0, c0, steps):
        """model.py:504-554 for `steps` consecutive teacher-forced actions in one cooperative launch (forward; one more for the
        backward): emb_all [steps*B, E] (from embed_actions), feat_all [steps*B, V, F], ctx_all [steps*B, L, D] -> h_tilde of
        every action [steps*B, H], or None when the kernel does not take this geometry / precision (the caller then loops over
        forward()). Dropout masks follow the per-step tags 't<i>.dec.h_prev' / 't<i>.dec.h1'."""
        TB, V, _ = feat_all.shape
        B = TB // steps
        L, D = ctx_all.shape[1], ctx_all.shape[2]
        if not (feat_all.is_contiguous() and ctx_all.is_contiguous() and
                ops.decoder_rollout_supported(*self._persistent_geometry(B, V, L, D))):
            return None
        p, tr, H = self.dropout_ratio, self.training, self.hidden_size
        m_hp, s = _source.mask_steps("dec.h_prev", (B, H), p, tr, feat_all.device, steps)
        m_h1, _ = _source.mask_steps("dec.h1", (B, H), p, tr, feat_all.device, steps)
        ht, _, _ = self._rollout_fn(emb_all.view(steps, B, -1), feat_all.view(steps, B, V, -1), ctx_all.view(steps, B, L, D),
                                    ctx_mask, h0, c0, m_hp, m_h1, s)
        return ht.view(steps * B, H)

    def forward(self, action, feature, cand_feat, h_0, prev_h1, c_0, ctx, ctx_mask=None, already_dropfeat=False,
                cand_leng=None, emb=None, want_logit=True):
        """Same contract as the reference; h_0 is ignored there too (model.py:472-474, 514). When not already_dropfeat
        and in training mode the dropped features are written back into the caller's tensors like the reference does
        (model.py:508, 557). Optional extensions: `cand_leng` (int32 [B]) lets the logits come out already -inf-masked;
        `emb` = this action's slice of embed_actions(); want_logit=False skips the candidate branch (the caller batches it
        with candidate_logits_steps) and returns logit = None."""
        A, p, tr = self.angle_feat_size, self.dropout_ratio, self.training
        if emb is None:
            emb = Fn.linear(action, self.embedding[0].weight, self.embedding[0].bias, "tanh")
            emb = _drop(emb, "dec.act", p, tr)
        if not already_dropfeat and tr:
            m, s = _source.mask("dec.feat", feature[..., :-A].shape, self.featdropout, tr, feature.device)
            if m is not None:
                dropped = torch.cat([Fn.dropout(feature[..., :-A], m, s), feature[..., -A:]], -1)
                with torch.no_grad():
                    feature.copy_(dropped)
                feature = dropped
        if (feature.dim() == 3 and feature.is_contiguous() and ctx.is_contiguous() and ops.decoder_rollout_supported(
                *self._persistent_geometry(feature.shape[0], feature.shape[1], ctx.shape[1], ctx.shape[2]))):
            # one cooperative launch for the whole step up to h_tilde (8 device-wide barriers instead of ~28 dependent launches)
            B, H = feature.shape[0], self.hidden_size
            m_hp, s = _source.mask("dec.h_prev", (B, H), p, tr, feature.device)
            m_h1, _ = _source.mask("dec.h1", (B, H), p, tr, feature.device)
            ht, h1s, cs = self._rollout_fn(emb.unsqueeze(0), feature.unsqueeze(0), ctx.unsqueeze(0), ctx_mask, prev_h1, c_0,
                                           m_hp, m_h1, s)
            h_1, c_1, h_tilde = h1s[0], cs[0], ht[0]
            return self._logit_tail(h_1, c_1, h_tilde, cand_feat, already_dropfeat, cand_leng, want_logit)
        h_prev_drop = _drop(prev_h1, "dec.h_prev", p, tr)
        attn_feat, _ = self.feat_att_layer(h_prev_drop, feature, output_tilde=False)
        xh = torch.cat((emb, attn_feat, prev_h1), 1)        # [x ; h]: one gate GEMM against [W_ih | W_hh]
        m_h1, s_h1 = _source.mask("dec.h1", (xh.shape[0], self.hidden_size), p, tr, xh.device)
        if m_h1 is not None:       # drop(h_1) comes out of the cell's pointwise kernel
            h_1, c_1, h_1_drop = Fn.LSTMCellFn.apply(xh, c_0, self.lstm.weight_ih, self.lstm.weight_hh, self.lstm.bias_ih,
                                                     self.lstm.bias_hh, m_h1, s_h1)
        else:
            h_1, c_1 = Fn.LSTMCellFn.apply(xh, c_0, self.lstm.weight_ih, self.lstm.weight_hh, self.lstm.bias_ih, self.lstm.bias_hh)
            h_1_drop = h_1
        h_tilde, alpha = self.attention_layer(h_1_drop, ctx, ctx_mask)
        return self._logit_tail(h_1, c_1, h_tilde, cand_feat, already_dropfeat, cand_leng, want_logit)

    def _logit_tail(self, h_1, c_1, h_tilde, cand_feat, already_dropfeat, cand_leng, want_logit):
        """drop(h_tilde) -> candidate logits (model.py:555-559)."""
        A, p, tr = self.angle_feat_size, self.dropout_ratio, self.training
        if not want_logit:
            return h_1, c_1, None, h_tilde, {}
        h_tilde_drop = _drop(h_tilde, "dec.htilde", p, tr)
        if not already_dropfeat and tr:
            m, s = _source.mask("dec.cand", cand_feat[..., :-A].shape, self.featdropout, tr, cand_feat.device)
            if m is not None:
                dropped = torch.cat([Fn.dropout(cand_feat[..., :-A], m, s), cand_feat[..., -A:]], -1)
                with torch.no_grad():
                    cand_feat.copy_(dropped)
                cand_feat = dropped
        _, logit = self.candidate_att_layer(h_tilde_drop, cand_feat, output_prob=False, cand_leng=cand_leng,
                                            rgb_channels=self.feature_size - A)
        return h_1, c_1, logit, h_tilde, {}


class Critic(nn.Module):
    """model.py:970-982."""

    def __init__(self, critic_dim=None, dropout=None):
        """model.Critic() takes no arguments and reads args.critic_dim / args.dropout (model.py:973-977)."""
        super().__init__()
        critic_dim, dropout = flag("critic_dim", critic_dim, 1024), flag("dropout", dropout, 0.5)
        self.dim, self.p = critic_dim, dropout
        self.state2value = nn.Sequential(nn.Linear(critic_dim, critic_dim), nn.ReLU(), nn.Dropout(dropout),
                                         nn.Linear(critic_dim, 1))

    def forward(self, state):
        x = Fn.linear(state, self.state2value[0].weight, self.state2value[0].bias, "relu")
        x = _drop(x, "critic", self.p, self.training)
        return Fn.linear(x, self.state2value[3].weight, self.state2value[3].bias).squeeze()

    def forward_steps(self, states, steps):
        """The critic over the hidden states of `steps` actions stacked along dim 0 ([steps*B, dim]) in one batched call
        (the reference calls it once per action, agent_dg.py:977); dropout masks follow the per-step tags 't<i>.critic'."""
        x = Fn.linear(states, self.state2value[0].weight, self.state2value[0].bias, "relu")
        m, scale = _source.mask_steps("critic", (states.shape[0] // steps, self.dim), self.p, self.training, states.device, steps)
        x = Fn.dropout(x, m, scale)
        return Fn.linear(x, self.state2value[3].weight, self.state2value[3].bias).reshape(-1)


# -------------------------------------------------------------------------------------------------------- encoder
class _BertSelfAttention(nn.Module):
    def __init__(self, hid):
        super().__init__()
        self.query, self.key, self.value = nn.Linear(hid, hid), nn.Linear(hid, hid), nn.Linear(hid, hid)


class _BertSelfOutput(nn.Module):
    def __init__(self, hid, eps, in_dim=None):
        super().__init__()
        self.dense = nn.Linear(in_dim or hid, hid)
        self.LayerNorm = nn.LayerNorm(hid, eps=eps)


class _BertAttention(nn.Module):
    def __init__(self, hid, eps):
        super().__init__()
        self.self = _BertSelfAttention(hid)
        self.output = _BertSelfOutput(hid, eps)


class _BertXAttention(nn.Module):
    def __init__(self, hid, eps):
        super().__init__()
        self.att = _BertSelfAttention(hid)
        self.output = _BertSelfOutput(hid, eps)


class _BertIntermediate(nn.Module):
    def __init__(self, hid, inter):
        super().__init__()
        self.dense = nn.Linear(hid, inter)


class _BertLayer(nn.Module):
    def __init__(self, hid, inter, eps):
        super().__init__()
        self.attention = _BertAttention(hid, eps)
        self.intermediate = _BertIntermediate(hid, inter)
        self.output = _BertSelfOutput(hid, eps, in_dim=inter)


class _LXRTXLayer(nn.Module):
    def __init__(self, hid, inter, eps):
        super().__init__()
        self.lang_self_att = _BertAttention(hid, eps)
        self.lang_inter = _BertIntermediate(hid, inter)
        self.lang_output = _BertSelfOutput(hid, eps, in_dim=inter)
        self.visn_self_att = _BertAttention(hid, eps)
        self.visn_inter = _BertIntermediate(hid, inter)
        self.visn_output = _BertSelfOutput(hid, eps, in_dim=inter)
        self.visual_attention = _BertXAttention(hid, eps)


class _BertEmbeddings(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.word_embeddings = nn.Embedding(cfg.vocab, cfg.bert_hidden, padding_idx=0)
        self.position_embeddings = nn.Embedding(cfg.max_pos, cfg.bert_hidden)
        self.token_type_embeddings = nn.Embedding(cfg.type_vocab, cfg.bert_hidden)
        self.LayerNorm = nn.LayerNorm(cfg.bert_hidden, eps=cfg.bert_eps)


class _BertPooler(nn.Module):
    def __init__(self, hid):
        super().__init__()
        self.dense = nn.Linear(hid, hid)


class _VisionEncoder(nn.Module):
    def __init__(self, vision_size, hid):
        super().__init__()
        self.visn_fc = nn.Linear(vision_size, hid)
        self.visn_layer_norm = nn.LayerNorm(hid, eps=1e-12)


class PackInfo:
    """Row bookkeeping for the packed (padding-free) evaluation of the frozen language / cross-modal stack: the valid tokens
    of `steps` x B sequences stored back to back. Padded keys get an additive -10000 in the reference (vilmodel.py:1339-1347),
    i.e. exactly 0 after the fp32 softmax, and padded query rows are never read downstream (r2rmodel.py:2326-2357), so
    dropping them changes no output."""

    def __init__(self, lengths_host, L, steps, device):
        lens = [int(x) for x in lengths_host] * steps
        self.L, self.steps, self.nseq = L, steps, len(lens)
        off, rows, o = [], [], 0
        for b, n in enumerate(lens):
            off.append(o)
            rows.extend(range(b * L, b * L + n))
            o += n
        self.ntok = o
        self.max_len = max(lens)
        self.off = torch.tensor(off, dtype=torch.int32, device=device)
        self.len = torch.tensor(lens, dtype=torch.int32, device=device)
        self.rows = torch.tensor(rows, dtype=torch.int32, device=device)      # packed row -> row of the padded [nseq*L, .] layout
        self.rows64 = self.rows.to(torch.int64)
        self.pair = (self.off, self.len)
        self._lens_host, self._off_host, self._bl, self._device = lens, off, None, device

    def bilstm_plan(self):
        """Bookkeeping of the padding-free bi-LSTM (csrc/bilstm_packed.cu): sequences ranked by length (descending, stable), tokens
        of the REVERSED sequences (r2rmodel.py:2326-2330) in position-block order. Built once per length pattern."""
        if self._bl is None:
            import ctypes
            lens, offs, L = self._lens_host, self._off_host, self.L
            R = len(lens)
            order = sorted(range(R), key=lambda i: -lens[i])
            n_rows = [0] * L
            for n in lens:
                for p in range(min(n, L)):
                    n_rows[p] += 1
            off = [0]
            for p in range(L):
                off.append(off[-1] + n_rows[p])
            src = []
            for p in range(L):
                for r in range(n_rows[p]):
                    i = order[r]
                    src.append(offs[i] + lens[i] - 1 - p)       # position p of the reversed sequence = original token len-1-p
            rank_of = [0] * R
            for r, i in enumerate(order):
                rank_of[i] = r
            dev = self._device

            class Plan:
                pass
            pl = Plan()
            pl.R, pl.L, pl.N = R, L, off[-1]
            pl.n_rows = (ctypes.c_int32 * L)(*n_rows)
            pl.off = (ctypes.c_int64 * (L + 1))(*off)
            pl.n_rows_list, pl.off_list = n_rows, off
            pl.perm = torch.tensor(order, dtype=torch.int32, device=dev)
            pl.perm64 = pl.perm.to(torch.int64)
            pl.rank_of = torch.tensor(rank_of, dtype=torch.int64, device=dev)
            pl.src = torch.tensor(src, dtype=torch.int64, device=dev)
            pl.steps_used = max(lens)
            self._bl = pl
        return self._bl

    def step_slice(self, t):
        """Rows of rollout step t inside a packed tensor built with steps > 1."""
        n1 = self.ntok // self.steps
        return slice(t * n1, (t + 1) * n1)


class DicModel(nn.Module):
    """vilmodel.py:1245-1423, forward-only kernels (train config: every output is detached, vilmodel.py:1377-1410)."""

    def __init__(self, cfg: PolicyConfig, vision_size):
        super().__init__()
        self.cfg = cfg
        hid, inter, eps = cfg.bert_hidden, cfg.bert_inter, cfg.bert_eps
        self.embeddings = _BertEmbeddings(cfg)
        self.pooler = _BertPooler(hid)
        self.lalayer = nn.ModuleList([_BertLayer(hid, inter, eps) for _ in range(cfg.la_layers)])
        self.addlayer = nn.ModuleList([_LXRTXLayer(hid, inter, eps) for _ in range(cfg.vl_layers)])
        self.vision_encoder = _VisionEncoder(vision_size, hid)
        self.update_lang_bert, self.update_add_layer = False, cfg.update_add_layer
        self._qkv_cache = {}

    # fused [3*hid, hid] QKV weights (one GEMM instead of three); rebuilt when the parameters change version
    def _qkv(self, att, which="qkv"):
        key = (id(att), which)
        # the raw-pointer optimizer (which does not bump tensor versions) only ever touches these weights in the finetune
        # configuration; in the train configuration the stack is frozen and the fused copies (and their fp16 twins) live on
        epoch = Fn.weights_epoch() if (self.update_add_layer and att.query.weight.requires_grad) else -1
        ver = (att.query.weight._version, att.key.weight._version, att.value.weight._version,
               att.query.weight.data_ptr(), epoch)
        hit = self._qkv_cache.get(key)
        if hit is not None and hit[0] == ver:
            return hit[1], hit[2]
        mods = {"qkv": (att.query, att.key, att.value), "kv": (att.key, att.value)}[which]
        with torch.no_grad():
            w = torch.cat([m.weight for m in mods], 0).contiguous()
            b = torch.cat([m.bias for m in mods], 0).contiguous()
        self._qkv_cache[key] = (ver, w, b)
        return w, b

    def _row_mask(self, tag, y, pack, training, stream_ok=False):
        """Dropout mask for a [.., hid] activation: padded per-step shape, or one row per packed token. stream_ok: the consumer
        can draw the flags itself (ops.DropStream) - used unless a test injects masks."""
        if stream_ok and _source.injected is None and ops.stream_dropout:
            return _source.stream(y.numel(), self.cfg.bert_dropout, training)
        if pack is None:
            return self._mask(tag, (y.shape[0] // self._steps,) + tuple(y.shape[1:]), training, y.device)
        p = self.cfg.bert_dropout
        if not training or p <= 0.0:
            return None, 1.0
        if _source.injected is None:
            return _source.mask(tag, (pack.ntok, y.shape[-1]), p, training, y.device)
        m, s = self._mask(tag, (pack.nseq // self._steps, pack.L, y.shape[-1]), training, y.device)   # tests: padded masks
        return (None, 1.0) if m is None else (m.reshape(-1, y.shape[-1])[pack.rows64].contiguous(), s)

    # ---- forward-only stack (train configuration: every output is detached, vilmodel.py:1377-1410). An activation travels as
    # (fp32 tensor, fp16 copy or None): where the token count gives the CTA-pair kernel enough tiles, the GEMMs take the fp16 copy
    # and fp16 weights (ops.linear_f16: tcgen05 kind::f16, twice the TF32 rate, the same 10-bit mantissa); the residual stream,
    # LayerNorm, softmax and every accumulation stay fp32.
    @staticmethod
    def _rows_of(t):
        return t.numel() // t.shape[-1]

    def _lin(self, act, w, b, epi=None, out_half=False):
        x, x16 = act
        src = x16 if x16 is not None else x
        M = self._rows_of(src)
        if x16 is not None and (x is None or ops.gemm_f16_supported(M, w.shape[0], w.shape[1])):
            y = ops.linear_f16(x16.reshape(M, x16.shape[-1]), ops.half_weight(w), b, epi, out_half)
            return y.view(tuple(x16.shape[:-1]) + (w.shape[0],))
        assert not out_half
        return ops.linear_fwd(x, w, b) if epi is None else ops.linear_fwd(x, w, b, epi)

    def _half_ok(self, M, n_out, n_in):
        return ops.gemm_f16_supported(M, n_out, n_in)

    def _out_ln(self, out_mod, act, resid, tag, training, pack=None):
        """dense -> dropout -> + resid -> LayerNorm (BertSelfOutput / BertOutput). Returns (out fp32, fp16 copy or None)."""
        w = out_mod.dense.weight
        src = act[1] if act[1] is not None else act[0]
        y16 = ops.half_ln_input and act[1] is not None and (act[0] is None or self._half_ok(self._rows_of(src), w.shape[0], w.shape[1]))
        y = self._lin(act, w, out_mod.dense.bias, out_half=bool(y16))      # fp16 branch output: read once by the LayerNorm below
        m, s = self._row_mask(tag, y, pack, training, stream_ok=True)
        hid = y.shape[-1]
        want16 = self._half_ok(self._rows_of(y), hid, hid)
        r = ops.dropout_residual_layernorm_fwd(y, resid, out_mod.LayerNorm.weight, out_mod.LayerNorm.bias, out_mod.LayerNorm.eps,
                                               m, s, half_copy=want16)
        return r if want16 else (r, None)

    def _probs_drop(self, tag, B, Lq, Lk, training, device, h16):
        """Keep flags of the attention-probability dropout: drawn inside the fp16 attention kernel (ops.DropStream) unless a
        test injects masks; a [B, heads, Lq, Lk] uint8 mask otherwise."""
        heads = self.cfg.bert_heads
        if h16 and _source.injected is None and ops.stream_dropout:
            return _source.stream(ops.mha_h16_stream_bytes(B, heads, Lq, Lk), self.cfg.bert_dropout, training)
        return self._mask(tag, (B // self._steps, heads, Lq, Lk), training, device)

    def _self_att(self, att_mod, act, key_pad, tag, training, pack=None):
        cfg = self.cfg
        hid = cfg.bert_hidden
        x = act[0]
        w, b = self._qkv(att_mod.self)
        f16 = act[1] is not None and self._half_ok(self._rows_of(x), hid, hid)      # context in fp16 for the output projection
        # fp16 Q / K / V straight from the projection's epilogue -> the fp16 attention kernel (dh = 64)
        h16 = f16 and ops.half_attention and self._half_ok(self._rows_of(x), 3 * hid, hid) and hid // cfg.bert_heads == 64
        qkv = self._lin(act, w, b, out_half=h16)
        if pack is not None:                     # x [ntok, hid]: only valid tokens, no key padding left to mask
            m, s = self._probs_drop(tag + ".probs", pack.nseq, pack.L, pack.L, training, x.device, h16)
            if h16:
                o = ops.mha_fwd_h16(qkv[:, :hid], qkv[:, hid:2 * hid], qkv[:, 2 * hid:], cfg.bert_heads, pack.pair, pack.pair,
                                    pack.L, pack.L, None, m, s)
            else:
                o = ops.mha_fwd_varlen(qkv[:, :hid], qkv[:, hid:2 * hid], qkv[:, 2 * hid:], cfg.bert_heads, pack.pair, pack.pair,
                                       pack.L, pack.L, m, s, out_half=f16)
            return self._out_ln(att_mod.output, (None, o) if f16 else (o, None), x, tag + ".out", training, pack)
        B, L = x.shape[0], x.shape[1]
        m, s = self._probs_drop(tag + ".probs", B, L, L, training, x.device, h16)
        if h16:
            o = ops.mha_fwd_h16(qkv[..., :hid], qkv[..., hid:2 * hid], qkv[..., 2 * hid:], cfg.bert_heads, key_pad=key_pad,
                                drop=m, drop_scale=s)
        else:
            o = ops.mha_fwd(qkv[..., :hid], qkv[..., hid:2 * hid], qkv[..., 2 * hid:], cfg.bert_heads, key_pad, m, s, out_half=f16)
        return self._out_ln(att_mod.output, (None, o) if f16 else (o, None), x, tag + ".out", training)

    def _cross_att(self, xatt, act, act_ctx, key_pad, tag, training, q_pack=None, k_pack=None):
        """q_pack: x is the packed language stream (ctx dense); k_pack: ctx is the packed language stream (x dense)."""
        cfg = self.cfg
        hid = cfg.bert_hidden
        x, ctx = act[0], act_ctx[0]
        w, b = self._qkv(xatt.att, "kv")
        f16 = act[1] is not None and self._half_ok(self._rows_of(x), hid, hid)
        h16 = (f16 and ops.half_attention and act_ctx[1] is not None and self._half_ok(self._rows_of(ctx), 2 * hid, hid)
               and hid // cfg.bert_heads == 64)
        q = self._lin(act, xatt.att.query.weight, xatt.att.query.bias, out_half=h16)
        kv = self._lin(act_ctx, w, b, out_half=h16)
        if q_pack is not None or k_pack is not None:
            pk = q_pack or k_pack
            Lq = pk.L if q_pack is not None else x.shape[1]
            Lk = pk.L if k_pack is not None else ctx.shape[1]
            m, s = self._probs_drop(tag + ".probs", pk.nseq, Lq, Lk, training, x.device, h16)
            if h16:
                o = ops.mha_fwd_h16(q, kv[..., :hid], kv[..., hid:], cfg.bert_heads, q_pack.pair if q_pack else None,
                                    k_pack.pair if k_pack else None, Lq, Lk, None, m, s)
            else:
                o = ops.mha_fwd_varlen(q, kv[..., :hid], kv[..., hid:], cfg.bert_heads, q_pack.pair if q_pack else None,
                                       k_pack.pair if k_pack else None, Lq, Lk, m, s, out_half=f16)
            return self._out_ln(xatt.output, (None, o) if f16 else (o, None), x, tag + ".out", training, q_pack)
        B, Lq, Lk = x.shape[0], x.shape[1], ctx.shape[1]
        m, s = self._probs_drop(tag + ".probs", B, Lq, Lk, training, x.device, h16)
        if h16:
            o = ops.mha_fwd_h16(q, kv[..., :hid], kv[..., hid:], cfg.bert_heads, key_pad=key_pad, drop=m, drop_scale=s)
        else:
            o = ops.mha_fwd(q, kv[..., :hid], kv[..., hid:], cfg.bert_heads, key_pad, m, s, out_half=f16)
        return self._out_ln(xatt.output, (None, o) if f16 else (o, None), x, tag + ".out", training)

    def _ffn(self, inter_mod, out_mod, act, tag, training, pack=None):
        x = act[0]
        w1 = inter_mod.dense.weight
        M = self._rows_of(x)
        f16 = act[1] is not None and self._half_ok(M, w1.shape[0], w1.shape[1]) and self._half_ok(M, w1.shape[1], w1.shape[0])
        if f16:     # the [M, 3072] GELU intermediate only ever exists as fp16 (half the bytes written and re-read)
            y16 = self._lin((None, act[1]), w1, inter_mod.dense.bias, ops.EPI_BIAS_GELU, out_half=True)
            return self._out_ln(out_mod, (None, y16), x, tag, training, pack)
        y = ops.linear_fwd(x, w1, inter_mod.dense.bias, ops.EPI_BIAS_GELU)
        return self._out_ln(out_mod, (y, None), x, tag, training, pack)

    def _act(self, x):
        """(fp32, fp16 copy) of a stream entering the stack."""
        hid = x.shape[-1]
        return (x, ops.to_half(x)) if self._half_ok(self._rows_of(x), hid, hid) else (x, None)

    @torch.no_grad()
    def language_stack(self, input_ids, pad_mask, training, steps=1, pack=None):
        """BertEmbeddings + la_layers x BertLayer (vilmodel.py:1366-1378). pad_mask: uint8 [B, L], 1 = padding.
        steps > 1 evaluates the stack for `steps` rollout actions at once: the instruction does not depend on the action
        taken, so the T per-action evaluations of the reference (agent_dg.py:789-797, each with its own dropout masks) are
        batched along dim 0 -> [steps*B, L, hid]. Same arithmetic per (action, episode); nothing is cached or skipped."""
        cfg, e = self.cfg, self.embeddings
        if steps > 1:
            input_ids = input_ids.repeat(steps, 1)
            pad_mask = pad_mask.repeat(steps, 1)
        B, L = input_ids.shape
        self._steps = steps
        m, s = self._mask("enc.emb", (B // steps, L, cfg.bert_hidden), training, input_ids.device)
        x = ops.embed_layernorm(input_ids, e.word_embeddings.weight, e.position_embeddings.weight,
                                e.token_type_embeddings.weight[0], e.LayerNorm.weight, e.LayerNorm.bias, e.LayerNorm.eps, m, s)
        if pack is not None:                     # drop the padding rows: everything below runs on valid tokens only
            x = ops.gather_rows(x.view(B * L, -1), pack.rows)
        act = self._act(x)
        for i, layer in enumerate(self.lalayer):
            a = self._self_att(layer.attention, act, pad_mask, "enc.la%d.att" % i, training, pack)
            act = self._ffn(layer.intermediate, layer.output, a, "enc.la%d.ffn" % i, training, pack)
        self._steps = 1
        return act[0]

    _steps = 1

    def _mask(self, tag, shape, training, device):
        """Dropout mask for one step, or for `self._steps` stacked steps while the batched language stack runs."""
        if self._steps > 1:
            return _source.mask_steps(tag, shape, self.cfg.bert_dropout, training, device, self._steps)
        return _source.mask(tag, shape, self.cfg.bert_dropout, training, device)

    @torch.no_grad()
    def cross_modal(self, lang, pad_mask, img_feats, training, steps=1, pack=None):
        """VisionEncoder + vl_layers x LXRTXLayer (vilmodel.py:1383-1410). steps > 1: the inputs hold `steps` rollout actions
        stacked along dim 0 (teacher-forced rollouts know every observation up front); per-action dropout masks are kept."""
        cfg, ve = self.cfg, self.vision_encoder
        self._steps = steps
        try:
            return self._cross_modal(lang, pad_mask, img_feats, training, pack)
        finally:
            self._steps = 1

    def _cross_modal(self, lang, pad_mask, img_feats, training, pack=None):
        cfg, ve = self.cfg, self.vision_encoder
        v = ops.linear_fwd(img_feats, ve.visn_fc.weight, ve.visn_fc.bias)
        m, s = self._mask("enc.visn", (v.shape[0] // self._steps,) + tuple(v.shape[1:]), training, v.device)
        visn = ops.dropout_residual_layernorm(v, None, ve.visn_layer_norm.weight, ve.visn_layer_norm.bias, 1e-12,
                                              None, 1.0, m, s)
        lang, visn = self._act(lang), self._act(visn)
        for i, layer in enumerate(self.addlayer):
            t = "enc.vl%d" % i
            l1 = self._cross_att(layer.visual_attention, lang, visn, None, t + ".x_lv", training, q_pack=pack)
            v1 = self._cross_att(layer.visual_attention, visn, lang, pad_mask, t + ".x_vl", training, k_pack=pack)
            l2 = self._self_att(layer.lang_self_att, l1, pad_mask, t + ".ls", training, pack)
            v2 = self._self_att(layer.visn_self_att, v1, None, t + ".vs", training)
            lang = self._ffn(layer.lang_inter, layer.lang_output, l2, t + ".lo", training, pack)
            visn = self._ffn(layer.visn_inter, layer.visn_output, v2, t + ".vo", training)
        return lang[0], visn[0]


def _dicmodel_grad_methods():
    """Differentiable twin of DicModel._cross_modal for the finetune config (--d_update_add_layer True, train.py:179-180):
    same kernels in the forward, autograd Functions (functions.MHAFn / DropResLNFn / LinearFn) so that the three
    cross-modal layers and the vision encoder receive gradients (vilmodel.py:1383-1410 without the detach at :1408-1410)."""

    def _out_ln_g(self, out_mod, x, resid, tag, training):
        y = Fn.linear(x, out_mod.dense.weight, out_mod.dense.bias)
        m, s = self._mask(tag, (y.shape[0] // self._steps,) + tuple(y.shape[1:]), training, y.device)
        return Fn.DropResLNFn.apply(y, resid, out_mod.LayerNorm.weight, out_mod.LayerNorm.bias, out_mod.LayerNorm.eps,
                                    m, s, None, 1.0)

    def _att_g(self, att, out_mod, x, ctx, key_pad, tag, training):
        cfg = self.cfg
        q = Fn.linear(x, att.query.weight, att.query.bias)
        k = Fn.linear(ctx, att.key.weight, att.key.bias)
        v = Fn.linear(ctx, att.value.weight, att.value.bias)
        B, Lq, Lk = x.shape[0], x.shape[1], ctx.shape[1]
        m, s = self._mask(tag + ".probs", (B // self._steps, cfg.bert_heads, Lq, Lk), training, x.device)
        o = Fn.MHAFn.apply(q, k, v, cfg.bert_heads, key_pad, m, s)
        return self._out_ln_g(out_mod, o, x, tag + ".out", training)

    def _ffn_g(self, inter_mod, out_mod, x, tag, training):
        y = Fn.linear(x, inter_mod.dense.weight, inter_mod.dense.bias, "gelu")
        return self._out_ln_g(out_mod, y, x, tag, training)

    def cross_modal_grad(self, lang, pad_mask, img_feats, training, steps=1):
        cfg, ve = self.cfg, self.vision_encoder
        self._steps = steps
        try:
            v = Fn.linear(img_feats, ve.visn_fc.weight, ve.visn_fc.bias)
            m, s = self._mask("enc.visn", (v.shape[0] // steps,) + tuple(v.shape[1:]), training, v.device)
            visn = Fn.DropResLNFn.apply(v, None, ve.visn_layer_norm.weight, ve.visn_layer_norm.bias, 1e-12, None, 1.0, m, s)
            for i, layer in enumerate(self.addlayer):
                t = "enc.vl%d" % i
                xa = layer.visual_attention
                l1 = self._att_g(xa.att, xa.output, lang, visn, None, t + ".x_lv", training)
                v1 = self._att_g(xa.att, xa.output, visn, lang, pad_mask, t + ".x_vl", training)
                l2 = self._att_g(layer.lang_self_att.self, layer.lang_self_att.output, l1, l1, pad_mask, t + ".ls", training)
                v2 = self._att_g(layer.visn_self_att.self, layer.visn_self_att.output, v1, v1, None, t + ".vs", training)
                lang = self._ffn_g(layer.lang_inter, layer.lang_output, l2, t + ".lo", training)
                visn = self._ffn_g(layer.visn_inter, layer.visn_output, v2, t + ".vo", training)
            return lang, visn
        finally:
            self._steps = 1

    DicModel._out_ln_g, DicModel._att_g, DicModel._ffn_g, DicModel.cross_modal_grad = _out_ln_g, _att_g, _ffn_g, cross_modal_grad


_dicmodel_grad_methods()


class DicEncoder(nn.Module):
    """r2rmodel.py:2199-2365. Constructor signature as the reference; `cfg` (optional) shrinks the transformer for tests."""

    lstm_num_layers = 1

    def __init__(self, vision_size, hidden_size, dec_hidden_size, dropout_ratio, bidirectional, update, bert_n_layers,
                 reverse_input, top_lstm, vl_layers, la_layers, bert_type="small", update_add_layer=True, cfg=None):
        super().__init__()
        if not (bidirectional and reverse_input and top_lstm and bert_n_layers == 1 and bert_type == "small" and not update):
            raise NotImplementedError("agent_dg constructs DicEncoder(.., True, False, 1, True, True, ..) (agent_dg.py:161)")
        from dataclasses import replace
        cfg = replace(cfg or FULL, vl_layers=vl_layers, la_layers=la_layers, enc_hidden=hidden_size, hidden=dec_hidden_size,
                      enc_dropout=dropout_ratio, update_add_layer=bool(update_add_layer))
        self.cfg = cfg
        self.hidden_size, self.dec_hidden_size, self.dropout_ratio = hidden_size, dec_hidden_size, dropout_ratio
        self.drop = nn.Dropout(p=dropout_ratio)
        self.num_directions = 2
        self.bert = DicModel(cfg, vision_size)
        self.lstm = nn.LSTM(cfg.bert_hidden, hidden_size, 1, batch_first=True, bidirectional=True)
        n_in = hidden_size * 2
        self.encoder2decoder_ht = nn.Linear(n_in, dec_hidden_size)     # unused when top_lstm; kept for state_dict
        self.encoder2decoder_ct = nn.Linear(n_in, dec_hidden_size)
        self.encoder_lstm2decoder_ht = nn.Linear(n_in, dec_hidden_size)
        self.encoder_lstm2decoder_ct = nn.Linear(n_in, dec_hidden_size)
        self.cache_language = False     # exact in eval mode; opt-in
        self._lang_cache = None
        self.pack_tokens = True         # evaluate the frozen transformer stack on valid tokens only (see PackInfo)
        self._packs = {}

    packed_lstm = True              # large batches on the tensor-core precision: padding-free recurrence (csrc/bilstm_packed.cu)

    def _packed_lstm(self, pack):
        return (self.packed_lstm and pack is not None and ops._precision == ops.PREC_TF32 and self.hidden_size % 32 == 0 and
                pack.nseq > ops.lib.load().dasa_bilstm_max_batch())

    def _pack(self, lengths, lengths_host, L, steps, device):
        """PackInfo for this batch, or None (lengths only known on the device, finetune config, or packing switched off)."""
        if not self.pack_tokens or self.cfg.update_add_layer:
            return None
        if lengths_host is None:
            if isinstance(lengths, (list, tuple)):
                lengths_host = lengths
            elif torch.is_tensor(lengths) and not lengths.is_cuda:
                lengths_host = lengths.tolist()
            else:
                return None
        key = (tuple(int(x) for x in lengths_host), L, steps, str(device))
        hit = self._packs.get(key)
        if hit is None:
            if len(self._packs) > 16:
                self._packs.clear()
            hit = self._packs[key] = PackInfo(key[0], L, steps, device)
        return hit

    class RolloutLanguage:
        """Language-stack output of all `steps` actions of a rollout: padded [steps, B, L, hid] or packed rows."""

        def __init__(self, x, pack, steps, B, L):
            self.x, self.pack, self.steps, self.B, self.L = x, pack, steps, B, L

        def __getitem__(self, t):
            if self.pack is None:
                return self.x.view(self.steps, self.B, self.L, -1)[t]
            return self.x[self.pack.step_slice(t)]

        def all(self):
            return self.x if self.pack is not None else self.x.reshape(self.steps * self.B, self.L, -1)

    def language_for_rollout(self, inputs, mask, steps, lengths_host=None):
        """The language stack for `steps` actions in one batched pass (see DicModel.language_stack); index the result with the
        action number to get that action's slice for forward(lang_out=...)."""
        L = mask.size(1)
        pad = mask.to(torch.uint8).contiguous()
        pack = self._pack(None, lengths_host, L, steps, inputs.device)
        out = self.bert.language_stack(inputs[:, :L].contiguous(), pad, self.training, steps, pack)
        return DicEncoder.RolloutLanguage(out, pack, steps, inputs.shape[0], L)

    def encode_rollout(self, inputs, mask, lengths, f_all, steps, lang_all=None, lengths_host=None):
        """Encoder for `steps` actions of a TEACHER-FORCED rollout in one batch (the trajectory, hence every panorama, is
        independent of the policy's outputs): f_all [steps*B, 36, F] -> (ctx [steps*B, L, 2H], decoder_init [B, Hd], c_t [B, Hd])
        where the decoder initial state comes from action 0 (agent_dg.py:812-815). Per-action dropout masks are preserved."""
        tr = self.training
        B, L = inputs.shape[0], mask.size(1)
        pad = mask.to(torch.uint8).contiguous()
        pack = self._pack(lengths, lengths_host, L, steps, f_all.device)
        if lang_all is None:
            lang_all = self.bert.language_stack(inputs[:, :L].contiguous(), pad, tr, steps, pack)
        elif isinstance(lang_all, DicEncoder.RolloutLanguage):
            assert (lang_all.pack is None) == (pack is None)
            lang_all = lang_all.all()
        if pack is None:
            lang_all = lang_all.reshape(steps * B, L, -1)
        pad_all = pad.repeat(steps, 1)
        len32 = torch.as_tensor(lengths, device=lang_all.device).to(torch.int32).repeat(steps)
        if self.cfg.update_add_layer:                    # finetune: gradients flow into the cross-modal layers
            lang, visn = self.bert.cross_modal_grad(lang_all, pad_all, f_all, tr, steps)
            rev = Fn.ReverseTokensFn.apply(lang, len32)
        elif pack is not None:
            lang, visn = self.bert.cross_modal(lang_all, pad_all, f_all, tr, steps, pack)
            rev = None if self._packed_lstm(pack) else ops.reverse_tokens_packed(lang, pack.off, pack.len, L)
        else:
            lang, visn = self.bert.cross_modal(lang_all, pad_all, f_all, tr, steps)
            rev = ops.reverse_tokens(lang, len32)
        l = self.lstm
        lstm_args = (l.weight_ih_l0, l.weight_hh_l0, l.bias_ih_l0, l.bias_hh_l0, l.weight_ih_l0_reverse, l.weight_hh_l0_reverse,
                     l.bias_ih_l0_reverse, l.bias_hh_l0_reverse)
        ctx_dropped = False
        if rev is None:      # padding-free recurrence straight from the packed tokens (reversal folded into its gather index)
            # `ctx = self.drop(ctx)` (r2rmodel.py:2357) rides on the kernels that write ctx / read its gradient: flags drawn in
            # place unless a test injects masks
            shape = (steps * B, L, 2 * self.hidden_size)
            if _source.injected is None and ops.stream_dropout:
                m, sc = _source.stream(shape[0] * shape[1] * shape[2], self.dropout_ratio, tr)
            else:
                m, sc = _source.mask_steps("enc.ctx", (B,) + shape[1:], self.dropout_ratio, tr, lang.device, steps)
            ctx, h_fin, c_fin = Fn.PackedBiLSTMFn.apply(lang, pack.bilstm_plan(), *lstm_args, m, sc)
            ctx_dropped = True
        else:
            ctx, h_fin, c_fin = Fn.BiLSTMFn.apply(rev, len32, *lstm_args)
        h_cat = torch.cat((h_fin[1, :B], h_fin[0, :B]), 1)
        c_cat = torch.cat((c_fin[1, :B], c_fin[0, :B]), 1)
        decoder_init = Fn.linear(h_cat, self.encoder_lstm2decoder_ht.weight, self.encoder_lstm2decoder_ht.bias, "tanh")
        c_t = Fn.linear(c_cat, self.encoder_lstm2decoder_ct.weight, self.encoder_lstm2decoder_ct.bias)
        if not ctx_dropped:
            m, sc = _source.mask_steps("enc.ctx", (B,) + tuple(ctx.shape[1:]), self.dropout_ratio, tr, ctx.device, steps)
            ctx = Fn.dropout(ctx, m, sc)
        return ctx, decoder_init, c_t

    def forward(self, inputs, mask, lengths, f_t_all=None, lang_out=None, lengths_host=None):
        """inputs [B, maxInput] int64, mask [B, Lmax] bool (True = pad), lengths [B] (sorted desc), f_t_all [B, 36, F]
        -> (ctx [B, Lmax, 2H], decoder_init [B, Hd], c_t [B, Hd], mask, vision_outputs [B, 36, 768]).
        lang_out (optional extension): this action's language-stack output from language_for_rollout()."""
        tr = self.training
        L = mask.size(1)
        pad = mask.to(torch.uint8).contiguous()
        ids = inputs[:, :L]
        pack = self._pack(lengths, lengths_host, L, 1, inputs.device)
        key = (inputs.data_ptr(), L, inputs._version, pack is not None)
        if lang_out is not None:
            lang0 = lang_out                             # this action's slice of language_for_rollout() (same packing)
            assert (lang0.dim() == 2) == (pack is not None), "lang_out layout does not match this call's packing"
        elif self.cache_language and not tr and self._lang_cache is not None and self._lang_cache[0] == key:
            lang0 = self._lang_cache[1]
        else:
            lang0 = self.bert.language_stack(ids, pad, tr, 1, pack)
            if self.cache_language and not tr:
                self._lang_cache = (key, lang0)
        len32 = torch.as_tensor(lengths, device=lang0.device).to(torch.int32)
        if self.cfg.update_add_layer:                    # finetune: gradients flow into the cross-modal layers
            lang, visn = self.bert.cross_modal_grad(lang0, pad, f_t_all, tr)
            rev = Fn.ReverseTokensFn.apply(lang, len32)
        elif pack is not None:
            lang, visn = self.bert.cross_modal(lang0, pad, f_t_all, tr, 1, pack)
            rev = None if self._packed_lstm(pack) else ops.reverse_tokens_packed(lang, pack.off, pack.len, L)
        else:
            lang, visn = self.bert.cross_modal(lang0, pad, f_t_all, tr)
            rev = ops.reverse_tokens(lang, len32)
        l = self.lstm
        lstm_args = (l.weight_ih_l0, l.weight_hh_l0, l.bias_ih_l0, l.bias_hh_l0, l.weight_ih_l0_reverse, l.weight_hh_l0_reverse,
                     l.bias_ih_l0_reverse, l.bias_hh_l0_reverse)
        if rev is None:
            ctx, h_fin, c_fin = Fn.PackedBiLSTMFn.apply(lang, pack.bilstm_plan(), *lstm_args)
        else:
            ctx, h_fin, c_fin = Fn.BiLSTMFn.apply(rev, len32, *lstm_args)
        # h_t = cat(enc_h_t[-1], enc_h_t[-2]): reverse-direction final state first (r2rmodel.py:2345-2346)
        h_cat = torch.cat((h_fin[1], h_fin[0]), 1)
        c_cat = torch.cat((c_fin[1], c_fin[0]), 1)
        decoder_init = Fn.linear(h_cat, self.encoder_lstm2decoder_ht.weight, self.encoder_lstm2decoder_ht.bias, "tanh")
        c_t = Fn.linear(c_cat, self.encoder_lstm2decoder_ct.weight, self.encoder_lstm2decoder_ct.bias)
        ctx = _drop(ctx, "enc.ctx", self.dropout_ratio, tr)
        return ctx, decoder_init, c_t, mask, visn


def build_policy(cfg: PolicyConfig = FULL, state=None, device="cuda"):
    """Construct (encoder, decoder, critic, adaIn) the way Seq2SeqAgent.__init__ does (agent_dg.py:161-201) and
    optionally load seeded / checkpoint state dicts (keys as in the reference)."""
    enc = DicEncoder(cfg.feat, cfg.enc_hidden, cfg.hidden, cfg.enc_dropout, True, False, 1, True, True, cfg.vl_layers,
                     cfg.la_layers, "small", cfg.update_add_layer, cfg=cfg)
    dec = BAttnDecoderLSTM(cfg.action_emb, cfg.hidden, cfg.dropout, feature_size=cfg.feat, angle_feat_size=cfg.angle_size,
                           featdropout=cfg.featdropout, shift_kernel_size=cfg.shift_kernel)
    cri = Critic(cfg.critic_dim, cfg.dropout)
    ada = DGAdaChannel(cfg.rgb_size)
    if state is not None:
        enc.load_state_dict(state["encoder"], strict=True)
        dec.load_state_dict(state["decoder"], strict=True)
        cri.load_state_dict(state["critic"], strict=True)
        ada.load_state_dict(state["adaIn"], strict=True)
    return tuple(m.to(device) for m in (enc, dec, cri, ada))
