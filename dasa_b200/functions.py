"""torch.autograd.Function wrappers: forward AND backward of every trainable op of the hot path run on the
hand-written kernels (closed forms: SURVEY.md Appendix A). Weight gradients are accumulated in place into
`param.grad` by beta=1 GEMMs (no per-step temporary), which is why the Functions return None for parameters.
"""
import torch

from . import ops
from .ops import (EPI_BIAS, EPI_BIAS_RELU, EPI_BIAS_TANH, EPI_GATE, EPI_NONE, EPI_TANH, _acc_grad)


def _zeros_like_grad(p):
    if p.grad is None:
        p.grad = torch.zeros_like(p)
    return p.grad


# ---------------------------------------------------------------------------------------- deferred weight gradients
# dW = dY^T X is a reduction over rows. Inside a rollout every step contributes only B (=20) rows to each decoder weight,
# i.e. a rank-20 update that re-reads and re-writes the whole weight gradient. With deferral the (dY, X) row blocks of all
# T steps are queued and ONE GEMM per weight with reduction length T*B (or T*L*B for the bi-LSTM) runs at flush time.
_defer = False
_queue = {}


def defer_weight_grads(flag=True):
    global _defer
    _defer = bool(flag)


def _bias_grads(dY, bias, bias2, half=None):
    """db += colsum(dY); a second bias fed by the same dY (nn.LSTM's b_ih / b_hh) reuses the column sums instead of a second
    pass over dY. half = (dY16, alpha): take the sums from the scaled fp16 copy instead (dY may then be None)."""
    if bias is None or not bias.requires_grad:
        bias, bias2 = bias2, None
    if bias is None or not bias.requires_grad:
        return

    def csum(out, acc):
        if dY is not None:
            ops.colsum(dY, out, acc)
        else:
            ops.colsum_h(half[0], half[1], out, acc)
    if bias2 is None or not bias2.requires_grad:
        csum(_zeros_like_grad(bias), True)
        return
    s = torch.empty_like(bias)
    csum(s, False)
    ops.axpy2d(1.0, s, _zeros_like_grad(bias), accumulate=True)
    ops.axpy2d(1.0, s, _zeros_like_grad(bias2), accumulate=True)


def _wgrad(weight, dy, x, bias=None, bias2=None):
    """dW += dy^T x ; db += colsum(dy) — immediately, or queued until flush_weight_grads()."""
    if not weight.requires_grad:
        return
    if not _defer:
        ops.linear_bwd_weight(dy, x, _zeros_like_grad(weight), True)
        _bias_grads(dy, bias, bias2)
        return
    ent = _queue.get(id(weight))
    if ent is None:
        ent = _queue[id(weight)] = [weight, bias, [], [], bias2]
    d2, x2 = ops._rows(dy)[0], ops._rows(x)[0]
    ent[2].append(d2)
    ent[3].append(x2)


def _wgrad16(weight, dy16, x16, alpha, dy, bias=None, bias2=None):
    """dW += alpha * dy16^T x16 on the fp16-operand tensor-core kernel (operand copies the producer already keeps: the bi-LSTM's
    scaled dgates and state rows); db += colsum(dy) from the fp32 gradient. Immediately, or queued like _wgrad."""
    if not weight.requires_grad:
        return
    if not _defer:
        ops.linear_bwd_weight_f16(dy16, x16, _zeros_like_grad(weight), alpha, True)
        _bias_grads(dy, bias, bias2, (dy16, alpha))
        return
    _queue16.append((weight, dy16, x16, alpha, dy, bias, bias2))


_queue16 = []

# Called (once, then cleared by the caller) when the backward pass reaches the encoder's packed bi-LSTM: in the batched
# teacher-forced rollout every decoder / critic node was created after the encoder, so autograd has finished all of them by then.
# trainer.RolloutTrainer uses it to flush + all-reduce the decoder group under the bi-LSTM's latency-bound backward chain.
pre_encoder_backward = None


def discard_weight_grads():
    """Drop every queued weight-gradient reduction (profiling helpers that run a backward pass without an optimizer step)."""
    _queue.clear()
    _queue16.clear()


def pending_weight_grads(owner):
    """Number of queued weight-gradient reductions whose target lives inside the flat buffer `owner`."""
    n = sum(1 for (weight, _, _, _, _) in _queue.values() if _inside(weight.grad, owner))
    return n + sum(1 for ent in _queue16 if _inside(ent[0].grad, owner))


def _inside(t, owner):
    """True when tensor t lives inside the flat buffer `owner`."""
    if t is None:
        return False
    lo = owner.data_ptr()
    return lo <= t.data_ptr() < lo + owner.numel() * owner.element_size()


def flush_weight_grads(owner=None):
    """Run the queued weight-gradient reductions (call once after loss.backward()). owner: a flat gradient buffer — only the
    weights whose .grad is a view of it are flushed (trainer.py reduces one optimizer group while the next one's GEMMs run)."""
    done = []
    for key, (weight, bias, dys, xs, bias2) in _queue.items():
        if owner is not None and not _inside(weight.grad, owner):
            continue
        # pieces with thousands of rows (the AdaIN gate sees the 25 200 view rows and the 8 400 candidate rows of a rollout) each
        # get their own accumulating long-K GEMM: concatenating them copied 550 MB per step; the many 20-row pieces of a
        # per-action path are still stacked into one operand first
        big = [i for i, d in enumerate(dys) if d.shape[0] >= 2048]
        small = [i for i in range(len(dys)) if i not in big]
        pieces = [(dys[i], xs[i]) for i in big]
        if small:
            pieces.append((dys[small[0]], xs[small[0]]) if len(small) == 1 else
                          (torch.cat([dys[i] for i in small], 0), torch.cat([xs[i] for i in small], 0)))
        for dY, X in pieces:
            ops.linear_bwd_weight(dY, X, _zeros_like_grad(weight), True)
            _bias_grads(dY, bias, bias2)
        done.append(key)
    for key in done:
        del _queue[key]
    keep = []
    for ent in _queue16:
        weight, dy16, x16, alpha, dy, bias, bias2 = ent
        if owner is not None and not _inside(weight.grad, owner):
            keep.append(ent)
            continue
        ops.linear_bwd_weight_f16(dy16, x16, _zeros_like_grad(weight), alpha, True)
        _bias_grads(dy, bias, bias2, (dy16, alpha))
    _queue16[:] = keep


def _rowmajor(t):
    """Row-strided 2-D gradients (column slices of a wider buffer) are passed to the kernels in place: every kernel takes a
    leading dimension. Only tensors without a unit inner stride are copied."""
    return t if (t.dim() == 2 and t.stride(1) == 1) else t.contiguous()


class LinearFn(torch.autograd.Function):
    """y = act(x W^T + b) with act in {none, tanh, relu} fused in the GEMM epilogue; act = gelu (finetune config) keeps the
    pre-activation for the backward and applies the erf-GELU in a second pass."""

    @staticmethod
    def forward(ctx, x, weight, bias, act):
        epi = {None: EPI_BIAS if bias is not None else EPI_NONE,
               "tanh": EPI_BIAS_TANH if bias is not None else EPI_TANH,
               "relu": EPI_BIAS_RELU, "gelu": EPI_BIAS}[act]
        y = ops.linear_fwd(x, weight, bias, epi)
        ctx.act, ctx.has_bias = act, bias is not None
        ctx.save_for_backward(x, weight, bias if bias is not None else x.new_empty(0), y if act else x.new_empty(0))
        return ops.gelu_fwd(y) if act == "gelu" else y

    @staticmethod
    def backward(ctx, dy):
        x, weight, bias, y = ctx.saved_tensors
        dy = dy.contiguous()
        if ctx.act:
            dy = ops.act_backward(ctx.act, dy, y)
        _wgrad(weight, dy, x, bias if ctx.has_bias else None)
        dx = ops.linear_bwd_input(dy, weight) if ctx.needs_input_grad[0] else None
        return dx, None, None, None


def linear(x, weight, bias=None, act=None):
    return LinearFn.apply(x, weight, bias, act)


class DropoutFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, mask, scale):
        ctx.scale = scale
        ctx.save_for_backward(mask)
        return ops.dropout_apply(x, mask, scale)

    @staticmethod
    def backward(ctx, dy):
        (mask,) = ctx.saved_tensors
        return ops.dropout_apply(_rowmajor(dy), mask, ctx.scale), None, None


def dropout(x, mask, scale):
    return x if mask is None else DropoutFn.apply(x, mask, scale)


class AdaINGateFn(torch.autograd.Function):
    """DGAdaChannel (ab_type=a, sigmoid) over the RGB slice of a feature tensor, fused with the decoder's drop_env mask:
    out[..., :C] = sigmoid(d[..., :C] W^T + b) * f[..., :C] (* mask*scale);  out[..., C:] = f[..., C:]  (angle part).
    One GEMM with the gate epilogue reading f / writing out in place at row stride F (no clones, no copy-backs)."""

    @staticmethod
    def forward(ctx, f, d, weight, bias, mask, scale, C):
        F_all = f.shape[-1]
        out = torch.empty(f.shape, device=f.device, dtype=torch.float32)
        f2, R, _, ldf = ops._rows(f)
        d2, _, _, ldd = ops._rows(d)
        o2, _, _, ldo = ops._rows_out(out)
        s = torch.empty(R, C, device=f.device, dtype=torch.float32)
        d16 = None
        if ops.half_gate and C % 8 == 0 and ops.gemm_f16_supported(R, C, C):
            # depth features are O(1..10): their fp16 copy keeps the 11 significant bits the TF32 kernel would use; the GEMM runs
            # at the fp16 tensor rate and the copy is the X operand of the weight gradient as well
            d16 = ops.to_half_rows(d2, ldd, R, C)
            ops.linear_f16_gate(d16, ops.half_weight(weight), bias, o2, ldo, f2, ldf, s, mask, scale)
        else:
            ops.gemm(d2, ldd, 1, weight, C, 1, o2, ldo, R, C, C, epilogue=EPI_GATE, bias=bias, gate_src=f2, ld_gate=ldf,
                     gate_out=s, ld_gate_out=C, drop_mask=mask, drop_scale=scale)
        ctx.d16 = d16
        if F_all > C:
            ops.axpy2d(1.0, f2[:, C:], o2[:, C:], accumulate=False)
        ctx.C, ctx.scale = C, scale
        ctx.save_for_backward(f, d, weight, bias, s, mask if mask is not None else f.new_empty(0))
        return out

    @staticmethod
    def backward(ctx, dout):
        f, d, weight, bias, s, mask = ctx.saved_tensors
        C = ctx.C
        mask = mask if mask.numel() else None
        dout = dout.contiguous()
        f2, R, _, ldf = ops._rows(f)
        d2, _, _, ldd = ops._rows(d)
        g2, _, _, ldg = ops._rows(dout)
        if ctx.d16 is not None and R >= 64:
            dg16 = torch.empty(R, C, device=f.device, dtype=torch.float16)     # only the scaled fp16 gradient is materialised
            ops.call("dasa_gate_backward_h", g2.data_ptr(), ldg, f2.data_ptr(), ldf, s.data_ptr(), C,
                     None if mask is None else mask.data_ptr(), float(ctx.scale), None, C, dg16.data_ptr(), 256.0, R, C,
                     ops._stream())
            _wgrad16(weight, dg16, ctx.d16, 1.0 / 256.0, None, bias)    # dW[C,C] += 2^-8 dg16^T d16 ; db += 2^-8 colsum(dg16)
        else:
            dg = torch.empty(R, C, device=f.device, dtype=torch.float32)
            ops.call("dasa_gate_backward", g2.data_ptr(), ldg, f2.data_ptr(), ldf, s.data_ptr(), C,
                     None if mask is None else mask.data_ptr(), float(ctx.scale), dg.data_ptr(), C, R, C, ops._stream())
            _wgrad(weight, dg, d2[:, :C], bias)        # dW[C,C] += dg^T d ; db += colsum(dg)
        return None, None, None, None, None, None, None


class ChannelModulateFn(torch.autograd.Function):
    """DGAdaStatChannel / DGAdaMeanChannel modulation out = a[:, None, :] * f + b[:, None, :] (agent_dg.py:1636, 1661);
    backward = one pass over (dout, f): da, db per (sample, channel), df only when f requires grad (env data does not)."""

    @staticmethod
    def forward(ctx, f, a, b):
        ctx.save_for_backward(f, a)
        return ops.channel_modulate(f, a, b)

    @staticmethod
    def backward(ctx, dout):
        f, a = ctx.saved_tensors
        if dout.stride(-1) != 1:
            dout = dout.contiguous()
        df, da, db = ops.channel_modulate_bwd(dout, f, a, want_df=ctx.needs_input_grad[0], want_db=ctx.needs_input_grad[2])
        return df, (da if ctx.needs_input_grad[1] else None), db


class ShiftAttnFn(torch.autograd.Function):
    """ShiftSoftDotAttention with output_tilde=False (model.py:318-353): returns (weighted_context, pre-shift softmax).
    linear_in and linear_shift read the same h, so their weights are stacked ([F + k (padded to 32), H], cached until the
    parameters change) and ONE projection produces the attention target and the shift-kernel logits; the backward writes dt
    and dkappa into one buffer and one GEMM returns dh."""

    @staticmethod
    def forward(ctx, h, context, w_in, w_shift, b_shift, headings):
        F_all, k = w_in.shape[0], w_shift.shape[0]
        Wc, bc = ops.stacked_weights((w_in, w_shift), 0, 32, (None, b_shift))
        tk = ops.linear_fwd(h, Wc, bc)                                    # [B, F + k + pad]
        t, kl = tk[:, :F_all], tk[:, F_all:F_all + k]
        wc, p, q, kappa = ops.row_attention_fwd(context, t, None, k, headings, kl)
        ctx.k, ctx.headings = k, headings
        ctx.save_for_backward(h, context, w_in, w_shift, b_shift, tk, p, q, kappa)
        ctx.mark_non_differentiable(p)
        return wc, p

    @staticmethod
    def backward(ctx, dwc, _dp):
        h, context, w_in, w_shift, b_shift, tk, p, q, kappa = ctx.saved_tensors
        F_all, k = w_in.shape[0], ctx.k
        need_dctx = ctx.needs_input_grad[1]
        dtk = torch.zeros(tk.shape, device=tk.device, dtype=torch.float32)       # padding columns must be exact zeros
        dt, dkl = dtk[:, :F_all], dtk[:, F_all:F_all + k]
        dctx, _, _ = ops.row_attention_bwd(context, tk[:, :F_all], p, q, kappa, _rowmajor(dwc), k, ctx.headings, need_dctx,
                                           dt=dt, dkl=dkl)
        _wgrad(w_in, dt, h)
        _wgrad(w_shift, dkl, h, b_shift)
        dh = None
        if ctx.needs_input_grad[0]:
            Wc, _ = ops.stacked_weights((w_in, w_shift), 0, 32, (None, b_shift))
            dh = ops.linear_bwd_input(dtk, Wc)
        return dh, dctx, None, None, None, None


class SoftDotAttnFn(torch.autograd.Function):
    """SoftDotAttention with output_tilde=True (model.py:268-296): returns (h_tilde, alpha)."""

    @staticmethod
    def forward(ctx, h, context, mask, w_in, w_out):
        B, Hq = h.shape
        D = context.shape[2]
        t = ops.linear_fwd(h, w_in)
        cat = torch.empty(B, D + Hq, device=h.device, dtype=torch.float32)     # [wc ; h]  (model.py:291)
        _, alpha, _, _ = ops.row_attention_fwd(context, t, mask, 0, 1, None, wc=cat)
        ops.axpy2d(1.0, h, cat[:, D:], accumulate=False)
        h_tilde = ops.linear_fwd(cat, w_out, None, EPI_TANH)
        ctx.D = D
        ctx.save_for_backward(h, context, w_in, w_out, t, alpha, cat, h_tilde)
        ctx.mark_non_differentiable(alpha)
        return h_tilde, alpha

    @staticmethod
    def backward(ctx, dht, _da):
        h, context, w_in, w_out, t, alpha, cat, h_tilde = ctx.saved_tensors
        D = ctx.D
        du = ops.act_backward("tanh", _rowmajor(dht), h_tilde)
        _wgrad(w_out, du, cat)
        dcat = ops.linear_bwd_input(du, w_out)
        need_dctx = ctx.needs_input_grad[1]
        dctx, dt, _ = ops.row_attention_bwd(context, t, alpha, alpha, None, dcat[:, :D], 0, 1, need_dctx)
        _wgrad(w_in, dt, h)
        dh = None
        if ctx.needs_input_grad[0]:
            dh = dcat[:, D:]                                   # accumulate in place in the [dwc ; dh] buffer (row stride D + H)
            ops.linear_bwd_input(dt, w_in, out=dh, beta=1.0)
        return dh, dctx, None, None, None


class CandLogitsFn(torch.autograd.Function):
    """candidate_att_layer with output_prob=False (model.py:559) + the agent's -inf masking (agent_dg.py:832-841)."""

    @staticmethod
    def forward(ctx, h, cand, leng, w_in, rgb_channels):
        t = ops.linear_fwd(h, w_in)
        logit = ops.cand_logits_fwd(cand, t, leng)
        ctx.rgb = rgb_channels
        ctx.save_for_backward(h, cand, leng if leng is not None else h.new_empty(0), w_in, t)
        return logit

    @staticmethod
    def backward(ctx, dlogit):
        h, cand, leng, w_in, t = ctx.saved_tensors
        leng = leng if leng.numel() else None
        need_dcand = ctx.needs_input_grad[1]
        dcand_rgb, dt = ops.cand_logits_bwd(cand, t, leng, dlogit, ctx.rgb, need_dcand)
        dcand = None
        if need_dcand:
            # only the AdaIN'd RGB slice carries gradient; the angle part is environment data
            dcand = torch.zeros(cand.shape, device=cand.device, dtype=torch.float32)
            ops.axpy2d(1.0, dcand_rgb, dcand[..., :ctx.rgb], accumulate=False)
        _wgrad(w_in, dt, h)
        dh = ops.linear_bwd_input(dt, w_in) if ctx.needs_input_grad[0] else None
        return dh, dcand, None, None, None


class LSTMCellFn(torch.autograd.Function):
    """nn.LSTMCell (model.py:437,514) on the concatenated input xh = [x ; h] (the caller builds it with the one torch.cat it
    needs anyway) against the column-stacked weight [W_ih | W_hh] (cached until the parameters change): ONE gate GEMM forward
    and ONE for d[x ; h] backward instead of two each, + the fused pointwise kernels. With a keep mask the decoder's
    drop(h_1) (model.py:515-516) is produced by the same pointwise kernel: returns (h1, c1, h1_dropped)."""

    @staticmethod
    def forward(ctx, xh, c, w_ih, w_hh, b_ih, b_hh, drop_mask=None, drop_scale=1.0):
        B, H = c.shape
        Wc, _ = ops.stacked_weights((w_ih, w_hh), 1)
        gates = ops.linear_fwd(xh, Wc)
        h1 = torch.empty(B, H, device=c.device, dtype=torch.float32)
        c1 = torch.empty(B, H, device=c.device, dtype=torch.float32)
        acts = torch.empty(B, 4 * H, device=c.device, dtype=torch.float32)
        h1d = torch.empty(B, H, device=c.device, dtype=torch.float32) if drop_mask is not None else None
        ops.lstm_pointwise_fwd(gates, None, b_ih, b_hh, c.contiguous(), None, h1, c1, h1d, acts, seq_mask=drop_mask,
                               seq_scale=drop_scale)
        ctx.scale = drop_scale
        ctx.save_for_backward(xh, c, w_ih, w_hh, b_ih, b_hh, acts, c1, drop_mask if drop_mask is not None else c.new_empty(0))
        if drop_mask is None:
            return h1, c1
        return h1, c1, h1d

    @staticmethod
    def backward(ctx, dh1, dc1, dh1d=None):
        xh, c, w_ih, w_hh, b_ih, b_hh, acts, c1, mask = ctx.saved_tensors
        B, H = c.shape
        n_x = w_ih.shape[1]
        dgates = torch.empty(B, 4 * H, device=c.device, dtype=torch.float32)
        dc0 = torch.empty(B, H, device=c.device, dtype=torch.float32)
        ops.lstm_pointwise_bwd(None if dh1 is None else _rowmajor(dh1), None if dh1d is None else _rowmajor(dh1d),
                               None if dc1 is None else _rowmajor(dc1), acts, c.contiguous(), c1, dgates, dc0,
                               dh2_mask=mask if (mask.numel() and dh1d is not None) else None, dh2_scale=ctx.scale)
        _wgrad(w_ih, dgates, xh[:, :n_x], b_ih, b_hh)         # db_ih == db_hh: one column sum
        _wgrad(w_hh, dgates, xh[:, n_x:])
        dxh = None
        if ctx.needs_input_grad[0]:
            Wc, _ = ops.stacked_weights((w_ih, w_hh), 1)
            dxh = ops.linear_bwd_input(dgates, Wc)
        return dxh, (dc0 if ctx.needs_input_grad[1] else None), None, None, None, None, None, None


class DecoderRolloutFn(torch.autograd.Function):
    """BAttnDecoderLSTM.forward up to h_tilde (model.py:504-554) for T consecutive actions in ONE cooperative launch each way
    (csrc/decoder_persist.cu). emb [T,B,E] (embedded, dropped actions), feat [T,B,V,F] (AdaIN'd, dropped views), ctx [T,B,L,D],
    (h0, c0) = (h_tilde, c) before action 0. Returns (h_tilde [T,B,H], h_1 [T,B,H], c [T,B,H]); T = 1 is the per-action form.
    Weight gradients: the kernel leaves the dY of every projection in [T*B, .] buffers and one long-K GEMM per weight follows
    (immediately, or at flush_weight_grads() when deferred)."""

    @staticmethod
    def forward(ctx_, emb, feat, ctx, ctx_mask, h0, c0, m_hprev, m_h1, scale, w_in, w_shift, b_shift, w_ih, w_hh, b_ih, b_hh,
                w_att_in, w_att_out, headings):
        k = w_shift.shape[0]
        Wf, bf = ops.stacked_weights((w_in, w_shift), 0, 32, (None, b_shift))
        Wl, _ = ops.stacked_weights((w_ih, w_hh), 1)
        o = ops.decoder_rollout_fwd(emb, feat, ctx, ctx_mask, h0, c0, m_hprev, m_h1, scale, Wf, bf, Wl, b_ih, b_hh, w_att_in,
                                    w_att_out, headings, k)
        ctx_.headings, ctx_.k, ctx_.scale = headings, k, scale
        ctx_.saved = o
        e = emb.new_empty(0)
        ctx_.save_for_backward(feat, ctx, ctx_mask if ctx_mask is not None else e, m_hprev if m_hprev is not None else e,
                               m_h1 if m_h1 is not None else e, w_in, w_shift, b_shift, w_ih, w_hh, b_ih, b_hh, w_att_in, w_att_out)
        return o["htilde"], o["h1"], o["c"][1:]

    @staticmethod
    def backward(ctx_, d_htilde, d_h1, d_c):
        (feat, ctx, ctx_mask, m_hprev, m_h1, w_in, w_shift, b_shift, w_ih, w_hh, b_ih, b_hh, w_att_in,
         w_att_out) = ctx_.saved_tensors
        o = ctx_.saved
        T, B, H = o["htilde"].shape
        F_all, k = w_in.shape[0], ctx_.k
        n_x = w_ih.shape[1]
        D = ctx.shape[3]
        if d_htilde is None:
            d_htilde = torch.zeros_like(o["htilde"])
        Wf, _ = ops.stacked_weights((w_in, w_shift), 0, 32, (None, b_shift))
        Wl, _ = ops.stacked_weights((w_ih, w_hh), 1)
        d_c_last = None
        if d_c is not None:
            # only the last cell state leaves the kernel as a recurrent carry (per-action use); earlier slots are internal
            d_c_last = d_c[T - 1]
        g = ops.decoder_rollout_bwd(o, feat, ctx, ctx_mask if ctx_mask.numel() else None, m_hprev if m_hprev.numel() else None,
                                    m_h1 if m_h1.numel() else None, ctx_.scale, ops.transposed_weight(Wf), ops.transposed_weight(Wl),
                                    ops.transposed_weight(w_att_in), ops.transposed_weight(w_att_out), ctx_.headings, k, d_htilde,
                                    d_h1, d_c_last)
        R = T * B
        du, dt2, dg, dtk = g["du"].view(R, H), g["dt2"].view(R, D), g["dgates"].view(R, 4 * H), g["dtk"].view(R, -1)
        cat, xh, hpd = o["cat"].view(R, -1), o["xh"].view(R, -1), o["hprev_drop"].view(R, H)
        _wgrad(w_att_out, du, cat)
        _wgrad(w_att_in, dt2, cat[:, D:])
        _wgrad(w_ih, dg, xh[:, :n_x], b_ih, b_hh)            # db_ih == db_hh: one column sum
        _wgrad(w_hh, dg, xh[:, n_x:])
        _wgrad(w_in, dtk[:, :F_all], hpd)
        _wgrad(w_shift, dtk[:, F_all:F_all + k], hpd, b_shift)
        ctx_.saved = None
        need = ctx_.needs_input_grad
        return (g["demb"] if need[0] else None, g["dfeat"] if need[1] else None, g["dctx"] if need[2] else None, None,
                g["dh0"] if need[4] else None, g["dc0"] if need[5] else None) + (None,) * 13


def invalidate_weight_caches():
    """Call after parameters were updated through raw pointers (the fused RMSprop kernel does not bump tensor versions)."""
    ops.weights_epoch += 1


def weights_epoch():
    return ops.weights_epoch


def _transposed(w):
    return ops.transposed_weight(w)


class BiLSTMFn(torch.autograd.Function):
    """Packed one-layer bidirectional nn.LSTM over the reversed token sequence (r2rmodel.py:2339-2357).
    x [B, L, In]; lengths int32 [B]. Returns ctx [B, L, 2H] (zero rows past each length), h_fin [2,B,H], c_fin [2,B,H]
    (index 0 = forward direction, 1 = reverse direction). Small batches run the fused per-step kernels (one C call for
    the whole sequence); larger ones the GEMM + pointwise path."""

    @staticmethod
    def forward(ctx, x, lengths, w_ih_f, w_hh_f, b_ih_f, b_hh_f, w_ih_r, w_hh_r, b_ih_r, b_hh_r):
        B, L, In = x.shape
        H = w_hh_f.shape[1]
        dev = x.device
        x = x.contiguous()
        out = torch.empty(B, L, 2 * H, device=dev, dtype=torch.float32)
        hs = torch.empty(2, L + 1, B, H, device=dev, dtype=torch.float32)     # state BEFORE step s at index s
        cs = torch.empty(2, L + 1, B, H, device=dev, dtype=torch.float32)
        hs[:, 0].zero_()
        cs[:, 0].zero_()
        acts = torch.empty(2, L, B, 4 * H, device=dev, dtype=torch.float32)
        params = ((w_ih_f, w_hh_f, b_ih_f, b_hh_f), (w_ih_r, w_hh_r, b_ih_r, b_hh_r))
        fused = B <= ops.lib.load().dasa_bilstm_max_batch() and H % 64 == 0
        # large batches on the tensor-core precision: grouped CTA-pair GEMM of both directions + one pointwise launch per step
        paired = (not fused) and ops._precision == ops.PREC_TF32 and H % 32 == 0
        if fused or paired:
            xp = [ops.linear_fwd(x, w_ih) for (w_ih, _, _, _) in params]       # [B, L, 4H] all time steps at once
            P2 = ops.lib.P * 2
            a = ops.lib.BiLstmFwd(P2(xp[0].data_ptr(), xp[1].data_ptr()), P2(w_hh_f.data_ptr(), w_hh_r.data_ptr()),
                                  P2(b_ih_f.data_ptr(), b_ih_r.data_ptr()), P2(b_hh_f.data_ptr(), b_hh_r.data_ptr()),
                                  P2(hs[0].data_ptr(), hs[1].data_ptr()), P2(cs[0].data_ptr(), cs[1].data_ptr()),
                                  P2(acts[0].data_ptr(), acts[1].data_ptr()), out.data_ptr(), lengths.data_ptr(), B, L, H)
            if fused:
                ops.call("dasa_bilstm_seq_fwd", ops.ctypes.byref(a), ops._precision, ops._stream())
            else:
                nb = ops.lib.load().dasa_bilstm_seq_gemm_workspace(B, H, 0)
                ws = ops.workspace(nb)
                ops.call("dasa_bilstm_seq_gemm_fwd", ops.ctypes.byref(a), ops._p(ws), ws.numel(), ops._stream())
        else:
            gh = torch.empty(B, 4 * H, device=dev, dtype=torch.float32)
            for d, (w_ih, w_hh, b_ih, b_hh) in enumerate(params):
                xp = ops.linear_fwd(x, w_ih)
                order = range(L) if d == 0 else range(L - 1, -1, -1)
                for s, l in enumerate(order):
                    ops.linear_fwd(hs[d, s], w_hh, out=gh)
                    ops.lstm_pointwise_fwd(xp[:, l], gh, b_ih, b_hh, cs[d, s], hs[d, s], hs[d, s + 1], cs[d, s + 1],
                                           out[:, l, d * H:(d + 1) * H], acts[d, s], lengths, l)
        h_fin = torch.stack((hs[0, L], hs[1, L]))
        c_fin = torch.stack((cs[0, L], cs[1, L]))
        ctx.fused = fused
        ctx.paired = paired
        ctx.save_for_backward(x, lengths, w_ih_f, w_hh_f, b_ih_f, b_hh_f, w_ih_r, w_hh_r, b_ih_r, b_hh_r, hs, cs, acts)
        return out, h_fin, c_fin

    @staticmethod
    def backward(ctx, dout, dh_fin, dc_fin):
        (x, lengths, w_ih_f, w_hh_f, b_ih_f, b_hh_f, w_ih_r, w_hh_r, b_ih_r, b_hh_r, hs, cs, acts) = ctx.saved_tensors
        B, L, In = x.shape
        H = w_hh_f.shape[1]
        dev = x.device
        dout = dout.contiguous() if dout is not None else torch.zeros(B, L, 2 * H, device=dev)
        params = ((w_ih_f, w_hh_f, b_ih_f, b_hh_f), (w_ih_r, w_hh_r, b_ih_r, b_hh_r))
        need_dx = ctx.needs_input_grad[0]
        dx = torch.zeros(B, L, In, device=dev, dtype=torch.float32) if need_dx else None
        dgates_all = torch.empty(2, L, B, 4 * H, device=dev, dtype=torch.float32)     # indexed by step s
        if ctx.fused or ctx.paired:
            dhf = dh_fin.contiguous() if dh_fin is not None else None
            dcf = dc_fin.contiguous() if dc_fin is not None else None
            wt = (_transposed(w_hh_f), _transposed(w_hh_r))
            work = torch.empty(2, 2, 2, B, H, device=dev, dtype=torch.float32)
            P2 = ops.lib.P * 2

            def pp(t, d):
                return None if t is None else t[d].data_ptr()
            a = ops.lib.BiLstmBwd(P2(wt[0].data_ptr(), wt[1].data_ptr()), P2(acts[0].data_ptr(), acts[1].data_ptr()),
                                  P2(cs[0].data_ptr(), cs[1].data_ptr()), dout.data_ptr(), P2(pp(dhf, 0), pp(dhf, 1)),
                                  P2(pp(dcf, 0), pp(dcf, 1)), P2(dgates_all[0].data_ptr(), dgates_all[1].data_ptr()),
                                  P2(work[0, 0].data_ptr(), work[0, 1].data_ptr()),
                                  P2(work[1, 0].data_ptr(), work[1, 1].data_ptr()), lengths.data_ptr(), B, L, H)
            if ctx.fused:
                ops.call("dasa_bilstm_seq_bwd", ops.ctypes.byref(a), ops._precision, ops._stream())
            else:
                nb = ops.lib.load().dasa_bilstm_seq_gemm_workspace(B, H, 1)
                ws = ops.workspace(nb)
                ops.call("dasa_bilstm_seq_gemm_bwd", ops.ctypes.byref(a), ops._p(ws), ws.numel(), ops._stream())
        for d, (w_ih, w_hh, b_ih, b_hh) in enumerate(params):
            dgates = dgates_all[d]
            order = list(range(L)) if d == 0 else list(range(L - 1, -1, -1))
            if not (ctx.fused or ctx.paired):
                dh = dh_fin[d].contiguous() if dh_fin is not None else torch.zeros(B, H, device=dev)
                dc = dc_fin[d].contiguous() if dc_fin is not None else torch.zeros(B, H, device=dev)
                dh_rec = torch.empty(B, H, device=dev, dtype=torch.float32)
                dh_pass = torch.empty(B, H, device=dev, dtype=torch.float32)
                dc_prev = torch.empty(B, H, device=dev, dtype=torch.float32)
                for s in range(L - 1, -1, -1):
                    l = order[s]
                    # dh (carried) + dout[:, l] -> dgates ; inactive rows pass dh/dc through untouched
                    ops.lstm_pointwise_bwd(dh, dout[:, l, d * H:(d + 1) * H], dc, acts[d, s], cs[d, s], cs[d, s + 1],
                                           dgates[s], dc_prev, dh_pass, lengths, l)
                    ops.linear_bwd_input(dgates[s], w_hh, out=dh_rec)
                    ops.axpy2d(1.0, dh_pass, dh_rec, accumulate=True)
                    dh, dh_rec = dh_rec, dh
                    dc, dc_prev = dc_prev, dc
            # weight gradients: one GEMM each over all L*B rows; x rows for step s are x[:, order[s]]
            xs = x if d == 0 else x.flip(1)
            xs = xs.transpose(0, 1).contiguous()                       # [L, B, In] in step order
            _wgrad(w_ih, dgates.view(L * B, 4 * H), xs.view(L * B, In), b_ih, b_hh)    # db_ih == db_hh: one column sum
            _wgrad(w_hh, dgates.view(L * B, 4 * H), hs[d, :L].reshape(L * B, H))
            if need_dx:
                dxs = ops.linear_bwd_input(dgates.view(L * B, 4 * H), w_ih).view(L, B, In).transpose(0, 1)
                dx += dxs if d == 0 else dxs.flip(1)
        return (dx, None) + (None,) * 8


class PackedBiLSTMFn(torch.autograd.Function):
    """The same bidirectional LSTM, padding-free (csrc/bilstm_packed.cu): x_packed [N_tokens, In] = the valid tokens of R
    sequences back to back in ORIGINAL token order (PackInfo layout); plan = PackInfo.bilstm_plan() (length ranking, position
    blocks, the gather index that also performs the token reversal of r2rmodel.py:2326-2330). Returns ctx [R, L, 2H] (zero rows
    past each length), h_fin [2, R, H], c_fin [2, R, H] in original sequence order, like BiLSTMFn."""

    @staticmethod
    def forward(ctx, x_packed, plan, w_ih_f, w_hh_f, b_ih_f, b_hh_f, w_ih_r, w_hh_r, b_ih_r, b_hh_r, drop=None, drop_scale=1.0):
        """drop: dropout on the returned ctx (r2rmodel.py:2357) fused into the kernels that write it / read its gradient: a uint8
        keep mask [R, L, 2H], an ops.DropStream (flags drawn in the kernel, forward and backward draw the same ones) or None."""
        R, L, N = plan.R, plan.L, plan.N
        H = w_hh_f.shape[1]
        dev = x_packed.device
        xc = x_packed.detach().index_select(0, plan.src)                       # [N, In] position-block order
        params = ((w_ih_f, w_hh_f, b_ih_f, b_hh_f), (w_ih_r, w_hh_r, b_ih_r, b_hh_r))
        use16 = ops.fused_lstm_cell and H >= 64 and H % 64 == 0 and xc.shape[1] % 8 == 0 and xc.shape[1] >= 64
        xc16 = ops.to_half(xc) if use16 else None                              # LayerNorm output, O(1): fp16 keeps TF32's 11 bits
        if use16 and ops.gemm_f16_supported(N, 4 * H, xc.shape[1]):
            xp = [ops.linear_f16(xc16, ops.half_weight(w_ih)) for (w_ih, _, _, _) in params]
        else:
            xp = [ops.linear_fwd(xc, w_ih) for (w_ih, _, _, _) in params]      # [N, 4H]: valid tokens only
        hprev = torch.empty(2, N, H, device=dev, dtype=torch.float32)
        cs = torch.empty(2, L + 1, R, H, device=dev, dtype=torch.float32)
        acts = torch.empty(2, N, 4 * H, device=dev, dtype=torch.float32)
        out = torch.zeros(R, L, 2 * H, device=dev, dtype=torch.float32)
        fin = torch.empty(2, 2, R, H, device=dev, dtype=torch.float32)         # [h | c][direction] in rank order
        P2 = ops.lib.P * 2
        cast = ops.ctypes.cast
        h16 = None
        if ops.fused_lstm_cell and H >= 64 and H % 64 == 0:                    # cell update in the recurrent GEMM's epilogue
            w16 = (ops.lstm_whh_interleaved(w_hh_f), ops.lstm_whh_interleaved(w_hh_r))
            h16 = torch.empty(2, N, H, device=dev, dtype=torch.float16)
            fused = (P2(w16[0].data_ptr(), w16[1].data_ptr()), P2(h16[0].data_ptr(), h16[1].data_ptr()))
        else:
            fused = (P2(None, None), P2(None, None))
        a = ops.lib.BiLstmPackedFwd(R, L, H, cast(plan.n_rows, ops.lib.P), cast(plan.off, ops.lib.P), plan.perm.data_ptr(),
                                    P2(xp[0].data_ptr(), xp[1].data_ptr()), P2(w_hh_f.data_ptr(), w_hh_r.data_ptr()),
                                    P2(b_ih_f.data_ptr(), b_ih_r.data_ptr()), P2(b_hh_f.data_ptr(), b_hh_r.data_ptr()),
                                    P2(hprev[0].data_ptr(), hprev[1].data_ptr()), P2(cs[0].data_ptr(), cs[1].data_ptr()),
                                    P2(acts[0].data_ptr(), acts[1].data_ptr()), out.data_ptr(),
                                    P2(fin[0, 0].data_ptr(), fin[0, 1].data_ptr()), P2(fin[1, 0].data_ptr(), fin[1, 1].data_ptr()),
                                    *_drop_fields(drop, drop_scale, (R, L, 2 * H)), *fused)
        ws = ops.workspace(ops.lib.load().dasa_bilstm_packed_workspace(R, H, 0))
        ops.call("dasa_bilstm_packed_fwd", ops.ctypes.byref(a), ops._p(ws), ws.numel(), ops._stream())
        h_fin = fin[0].index_select(1, plan.rank_of)
        c_fin = fin[1].index_select(1, plan.rank_of)
        ctx.plan, ctx.drop, ctx.drop_scale = plan, drop, drop_scale
        ctx.half = (xc16, h16) if (use16 and h16 is not None) else None        # fp16 operand copies for the weight gradients
        ctx.save_for_backward(xc, w_ih_f, w_hh_f, b_ih_f, b_hh_f, w_ih_r, w_hh_r, b_ih_r, b_hh_r, hprev, cs, acts)
        return out, h_fin, c_fin

    @staticmethod
    def backward(ctx, dout, dh_fin, dc_fin):
        (xc, w_ih_f, w_hh_f, b_ih_f, b_hh_f, w_ih_r, w_hh_r, b_ih_r, b_hh_r, hprev, cs, acts) = ctx.saved_tensors
        global pre_encoder_backward
        if pre_encoder_backward is not None:
            hook, pre_encoder_backward = pre_encoder_backward, None
            hook()
        plan = ctx.plan
        R, L, N = plan.R, plan.L, plan.N
        H = w_hh_f.shape[1]
        dev = xc.device
        dout = dout.contiguous() if dout is not None else torch.zeros(R, L, 2 * H, device=dev)
        params = ((w_ih_f, w_hh_f, b_ih_f, b_hh_f), (w_ih_r, w_hh_r, b_ih_r, b_hh_r))
        dhf = dh_fin.index_select(1, plan.perm64).contiguous() if dh_fin is not None else None      # rank order
        dcf = dc_fin.index_select(1, plan.perm64).contiguous() if dc_fin is not None else None
        wt = (_transposed(w_hh_f), _transposed(w_hh_r))
        work = torch.empty(2, 2, R, H, device=dev, dtype=torch.float32)
        P2 = ops.lib.P * 2
        cast = ops.ctypes.cast

        def pp(t, d):
            return None if t is None else t[d].data_ptr()
        dg16 = None
        if ops.fused_lstm_cell and H >= 64 and H % 64 == 0:       # dh = dgates W_hh on fp16 operands (scaled dgates copy)
            wt16 = (ops.half_weight(wt[0]), ops.half_weight(wt[1]))
            dg16 = torch.empty(2, N, 4 * H, device=dev, dtype=torch.float16)
            half = (P2(wt16[0].data_ptr(), wt16[1].data_ptr()), P2(dg16[0].data_ptr(), dg16[1].data_ptr()))
        else:
            half = (P2(None, None), P2(None, None))
        half16 = ctx.half if (dg16 is not None and ctx.half is not None and N >= 64) else None
        # the fp32 dgates are only read by dX (finetune configuration) and by the TF32 weight-gradient path
        need32 = half16 is None or ctx.needs_input_grad[0]
        dgates = torch.empty(2, N, 4 * H, device=dev, dtype=torch.float32) if need32 else None
        a = ops.lib.BiLstmPackedBwd(R, L, H, cast(plan.n_rows, ops.lib.P), cast(plan.off, ops.lib.P), plan.perm.data_ptr(),
                                    P2(wt[0].data_ptr(), wt[1].data_ptr()), P2(acts[0].data_ptr(), acts[1].data_ptr()),
                                    P2(cs[0].data_ptr(), cs[1].data_ptr()), dout.data_ptr(), P2(pp(dhf, 0), pp(dhf, 1)),
                                    P2(pp(dcf, 0), pp(dcf, 1)), P2(pp(dgates, 0), pp(dgates, 1)),
                                    P2(work[0].data_ptr(), work[1].data_ptr()), *_drop_fields(ctx.drop, ctx.drop_scale, (R, L, 2 * H)),
                                    *half)
        ws = ops.workspace(ops.lib.load().dasa_bilstm_packed_workspace(R, H, 1))
        ops.call("dasa_bilstm_packed_bwd", ops.ctypes.byref(a), ops._p(ws), ws.numel(), ops._stream())
        dxc = None
        for d, (w_ih, w_hh, b_ih, b_hh) in enumerate(params):
            if half16 is not None:      # dW = 2^-8 (dgates * 2^8)^T x on kind::f16 with the copies the recurrence keeps
                _wgrad16(w_ih, dg16[d], half16[0], 1.0 / 256.0, None, b_ih, b_hh)
                _wgrad16(w_hh, dg16[d], half16[1][d], 1.0 / 256.0, None)
            else:
                _wgrad(w_ih, dgates[d], xc, b_ih, b_hh)                        # db_ih == db_hh: one column sum
                _wgrad(w_hh, dgates[d], hprev[d])
            if ctx.needs_input_grad[0]:
                g = ops.linear_bwd_input(dgates[d], w_ih)
                dxc = g if dxc is None else dxc + g
        dx = None
        if dxc is not None:
            dx = torch.empty_like(dxc)
            dx[plan.src] = dxc                                                 # the gather index is a bijection of the N tokens
        return (dx, None) + (None,) * 10


def _drop_fields(drop, scale, shape):
    """(out_mask, drop_seed_dev, drop_seed, drop_base, drop_p, drop_scale) of the fused-output-dropout ABI structs."""
    if drop is not None and not isinstance(drop, ops.DropStream):
        assert tuple(drop.shape) == tuple(shape) and drop.dtype == torch.uint8 and drop.is_contiguous(), (drop.shape, shape)
    mp, sp, seed, base, p = ops._drop_args(drop)
    return mp, sp, seed, base, p, float(scale)


class MaskedCEFn(torch.autograd.Function):
    """sum-reduced cross entropy with ignore_index over candidate logits that already carry -inf (agent_dg.py:850).
    Returns (loss[1], greedy action[B])."""

    @staticmethod
    def forward(ctx, logit, target, ignore_index):
        loss = torch.zeros(1, device=logit.device, dtype=torch.float32)
        dlogit, action, _, _ = ops.masked_ce(logit.contiguous(), target, ignore_index, 1.0, loss)
        ctx.save_for_backward(dlogit)
        ctx.mark_non_differentiable(action)
        return loss, action

    @staticmethod
    def backward(ctx, dloss, _da):
        (dlogit,) = ctx.saved_tensors
        return dlogit * dloss, None, None


class PolicySampleFn(torch.autograd.Function):
    """feedback='sample' action selection (agent_dg.py:876-882): returns (action[B], log_prob[B], entropy[B]).
    `u` = one uniform per episode (device RNG) or None with `action_in` (injected actions, tests) / argmax."""

    @staticmethod
    def forward(ctx, logit, u, action_in):
        action, lp, ent, probs = ops.policy_sample_fwd(logit.contiguous(), u, action_in)
        ctx.save_for_backward(probs, action, ent)
        ctx.mark_non_differentiable(action)
        return action, lp, ent

    @staticmethod
    def backward(ctx, _da, dlp, dent):
        probs, action, ent = ctx.saved_tensors
        dlp = dlp.contiguous() if dlp is not None else None
        dent = dent.contiguous() if dent is not None else None
        return ops.policy_sample_bwd(probs, action, dlp, dent, ent), None, None


class A2CLossFn(torch.autograd.Function):
    """The A2C epilogue of vl_rollout (agent_dg.py:943-999) as one fused kernel over [T,B] stacks: returns (loss[1], total[1])."""

    @staticmethod
    def forward(ctx, logp, ent, value, last_value, reward, mask, ended, gamma, ent_coef, normalize):
        loss, total, dlogp, dent, dvalue = ops.a2c_loss(logp.contiguous(), None if ent is None else ent.contiguous(),
                                                        value.contiguous(), last_value.contiguous(), reward, mask, ended,
                                                        gamma, ent_coef, normalize)
        ctx.save_for_backward(dlogp, dent, dvalue)
        ctx.mark_non_differentiable(total)
        return loss, total

    @staticmethod
    def backward(ctx, dloss, _dt):
        dlogp, dent, dvalue = ctx.saved_tensors
        return dlogp * dloss, (None if dent is None else dent * dloss), dvalue * dloss, None, None, None, None, None, None, None


# ------------------------------------------------------------------------------- finetune config (--d_update_add_layer)
class MHAFn(torch.autograd.Function):
    """BertSelfAttention / BertOutAttention core (vilmodel.py:203-236, 479-506): softmax(q k^T / sqrt(dh) + pad) (dropout) v."""

    @staticmethod
    def forward(ctx, q, k, v, heads, key_pad, drop_mask, drop_scale):
        out, probs = ops.mha_fwd(q, k, v, heads, key_pad, drop_mask, drop_scale, save_probs=True)
        ctx.heads, ctx.scale = heads, drop_scale
        ctx.save_for_backward(q, k, v, probs, drop_mask if drop_mask is not None else q.new_empty(0))
        return out

    @staticmethod
    def backward(ctx, dout):
        q, k, v, probs, mask = ctx.saved_tensors
        dq, dk, dv = ops.mha_bwd(q, k, v, probs, dout, ctx.heads, mask if mask.numel() else None, ctx.scale)
        return dq, dk, dv, None, None, None, None


class DropResLNFn(torch.autograd.Function):
    """BertSelfOutput / BertOutput / VisionEncoder tail: LN(x*mask*scale + resid) * gamma + beta (* post_mask*post_scale)."""

    @staticmethod
    def forward(ctx, x, resid, gamma, beta, eps, mask, scale, post_mask, post_scale):
        out, stats, z = ops.dropout_residual_layernorm(x, resid, gamma, beta, eps, mask, scale, post_mask, post_scale, save=True)
        ctx.scale, ctx.post_scale, ctx.has_resid = scale, post_scale, resid is not None
        e = x.new_empty(0)
        ctx.save_for_backward(z, stats, gamma, beta, mask if mask is not None else e, post_mask if post_mask is not None else e)
        return out

    @staticmethod
    def backward(ctx, dout):
        z, stats, gamma, beta, mask, post_mask = ctx.saved_tensors
        dg, db = _zeros_like_grad(gamma), _zeros_like_grad(beta)
        dx, dresid = ops.layernorm_bwd(dout.contiguous(), z, gamma, stats, dg, db, mask if mask.numel() else None, ctx.scale,
                                       post_mask if post_mask.numel() else None, ctx.post_scale)
        return dx, (dresid if ctx.has_resid else None), None, None, None, None, None, None, None


class ReverseTokensFn(torch.autograd.Function):
    """Per-sample token reversal (r2rmodel.py:2326-2330); self-inverse, so the backward is the same kernel."""

    @staticmethod
    def forward(ctx, x, lengths_i32):
        ctx.save_for_backward(lengths_i32)
        return ops.reverse_tokens(x, lengths_i32)

    @staticmethod
    def backward(ctx, dy):
        (lengths,) = ctx.saved_tensors
        return ops.reverse_tokens(dy.contiguous(), lengths), None
