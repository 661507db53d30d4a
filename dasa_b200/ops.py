"""Tensor-level wrappers over the C ABI (include/dasa_b200.h) and the autograd Functions built from them.

PyTorch is plumbing here (device memory, streams, the autograd tape); every numeric step of the hot path is one of
the hand-written sm_100a kernels in dasa_b200/csrc. There is no CPU / eager fallback: tensors must be CUDA fp32.
"""
import ctypes
import weakref

import torch

from . import lib
from .lib import Epilogue, call  # noqa: F401

EPI_NONE, EPI_BIAS, EPI_BIAS_TANH, EPI_BIAS_GELU, EPI_BIAS_RELU, EPI_GATE, EPI_TANH = range(7)
PREC_FP32, PREC_TF32 = 0, 1

_precision = PREC_FP32
_workspaces = {}


def set_precision(name):
    """'fp32' (FFMA, exact fp32 products) or 'tf32' (tcgen05 tensor cores for the dense projections)."""
    global _precision
    _precision = {"fp32": PREC_FP32, "tf32": PREC_TF32}[name]


def get_precision():
    return "tf32" if _precision == PREC_TF32 else "fp32"


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return None if t is None else t.data_ptr()


def _chk(*ts):
    for t in ts:
        if t is not None:
            if not t.is_cuda:
                raise lib.DasaError("dasa_b200 ops need CUDA tensors (no CPU fallback)")
            if t.dtype != torch.float32:
                raise lib.DasaError("expected float32, got %s" % t.dtype)


def _rows(t):
    """[.., C] tensor -> (tensor, R, C, ld): rows addressed as base + r*ld, unit inner stride. Strided slices such as
    feat[..., :2048] of a [B, V, 2176] buffer are passed through in place; anything irregular is made contiguous
    (inputs only — callers allocate outputs themselves, so outputs are always regular)."""
    if t.dim() == 1:
        t = t.unsqueeze(0)
    if t.stride(-1) != 1 and t.shape[-1] != 1:
        t = t.contiguous()
    if t.dim() == 2:
        return t, t.shape[0], t.shape[1], (t.stride(0) if t.shape[0] > 1 else max(t.stride(0), t.shape[1]))
    uniform = all(t.shape[i] == 1 or t.stride(i) == t.stride(i + 1) * t.shape[i + 1] for i in range(t.dim() - 2))
    if not uniform or t.shape[-2] == 1:
        t = t.contiguous()
    R, C, ld = t.numel() // t.shape[-1], t.shape[-1], t.stride(-2)
    return t.as_strided((R, C), (ld, 1)), R, C, ld          # genuine 2-D view over the same storage


def _rows_out(t):
    """Like _rows, for OUTPUT tensors: the kernel must write the caller's storage, so a copy is an error."""
    t2, R, C, ld = _rows(t)
    if t2.data_ptr() != t.data_ptr():
        raise lib.DasaError("output tensor is not row-regular (shape %s strides %s)" % (tuple(t.shape), t.stride()))
    return t2, R, C, ld


def workspace(nbytes):
    dev = torch.cuda.current_device()
    ws = _workspaces.get(dev)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(int(nbytes), 64 << 20), dtype=torch.uint8, device="cuda")
        _workspaces[dev] = ws
    return ws


# ----------------------------------------------------------------------------------------------------- GEMM
def gemm(A, lda, a_kmajor, Bm, ldb, b_kmajor, C, ldc, M, N, K, alpha=1.0, beta=0.0, epilogue=EPI_NONE, bias=None,
         gate_src=None, ld_gate=0, gate_out=None, ld_gate_out=0, drop_mask=None, drop_scale=1.0, precision=None):
    prec = _precision if precision is None else precision
    nbytes = lib.load().dasa_gemm_workspace_bytes(M, N, K, prec)
    ws = workspace(nbytes) if nbytes else None
    epi = Epilogue(_p(bias), _p(gate_src), ld_gate, _p(gate_out), ld_gate_out, _p(drop_mask), float(drop_scale))
    call("dasa_gemm", int(a_kmajor), int(b_kmajor), M, N, K, float(alpha), _p(A), lda, _p(Bm), ldb, float(beta), _p(C), ldc,
         epilogue, ctypes.byref(epi), prec, _p(ws), ws.numel() if ws is not None else 0, _stream())


def linear_fwd(x, w, bias=None, epilogue=None, out=None, beta=0.0, precision=None):
    """y[.., N] = epi(x[.., K] @ w[N, K]^T (+ bias)) ; x rows may be strided."""
    _chk(x, w, bias, out)
    x2, M, K, lda = _rows(x)
    N = w.shape[0]
    assert w.shape[1] == K and w.is_contiguous(), (w.shape, K)
    if out is None:
        out = torch.empty(*x.shape[:-1], N, device=x.device, dtype=torch.float32) if x.dim() > 1 else \
            torch.empty(N, device=x.device, dtype=torch.float32)
    o2, Mo, No, ldc = _rows_out(out)
    assert Mo == M and No == N
    if epilogue is None:
        epilogue = EPI_BIAS if bias is not None else EPI_NONE
    gemm(x2, lda, 1, w, K, 1, o2, ldc, M, N, K, beta=beta, epilogue=epilogue, bias=bias, precision=precision)
    return out


def transpose(x, out=None):
    """x [R, C] (rows may be strided) -> [C, Rp] with Rp = R rounded up to 4 (row stride TMA-legal); returns out[:, :R]."""
    x2, R, C, ld = _rows(x)
    Rp = (R + 3) & ~3
    if out is None:
        out = torch.empty(C, Rp, device=x.device, dtype=torch.float32)
    call("dasa_transpose", _p(x2), ld, R, C, _p(out), out.stride(0), _stream())
    return out[:, :R]


use_mn_major = True    # feed tcgen05 MN-major operands where the pair kernel takes the problem (else transposed copies)
_wt_cache = {}
weights_epoch = 0      # bumped by the optimizer step (raw-pointer updates do not bump tensor versions)


def transposed_weight(w):
    """W^T ([in, out], contiguous rows) cached until the parameter changes: lets dX = dY.W run as a K-major GEMM."""
    key = id(w)
    tag = (w._version, weights_epoch, w.data_ptr())
    hit = _wt_cache.get(key)
    if hit is not None and hit[0] == tag and hit[2]() is w:      # the weakref guards against id()/address reuse
        return hit[1]
    if len(_wt_cache) > 256:
        for k in [k for k, v in _wt_cache.items() if v[2]() is None]:
            del _wt_cache[k]
    wt = transpose(w.detach())
    _wt_cache[key] = (tag, wt, weakref.ref(w))
    return wt


half_stack = True      # frozen transformer stack (forward-only): fp16 operands on the kind::f16 tensor-core path where the shapes allow
half_attention = True  # ... and fp16 Q / K / V into the fp16 attention kernel (dasa_mha_fwd_h16)
half_ln_input = True   # ... and the fp16 output of the dense projections into the residual LayerNorm
stream_dropout = True  # forward-only dropout sites of that stack draw their keep flags in the consuming kernel (DropStream)


def gemm_f16_supported(M, N, K):
    """True when dasa_gemm_f16 takes the shape (enough 256 x 256 tiles for the CTA-pair kernel) under the tensor-core precision."""
    return half_stack and _precision == PREC_TF32 and bool(lib.load().dasa_gemm_f16_supported(int(M), int(N), int(K)))


def linear_f16(x16, w16, bias=None, epilogue=None, out_half=False):
    """y[M, N] = epilogue(x16[M, K] w16[N, K]^T + bias) on the fp16-operand tcgen05 kernel; y fp32, or fp16 with out_half."""
    assert x16.dtype == torch.float16 and w16.dtype == torch.float16 and x16.dim() == 2 and w16.dim() == 2
    assert x16.stride(1) == 1 and w16.stride(1) == 1
    M, K = x16.shape
    N = w16.shape[0]
    epi = epilogue if epilogue is not None else (EPI_BIAS if bias is not None else EPI_NONE)
    y = torch.empty(M, N, device=x16.device, dtype=torch.float16 if out_half else torch.float32)
    e = lib.Epilogue()
    e.bias = _p(bias)
    call("dasa_gemm_f16", M, N, K, _p(x16), x16.stride(0), _p(w16), w16.stride(0), _p(y), y.stride(0), int(out_half), int(epi),
         ctypes.byref(e), _stream())
    return y


def to_half(x):
    """fp16 copy of a contiguous fp32 tensor (dasa_f32_to_f16)."""
    x = x if x.is_contiguous() else x.contiguous()
    out = torch.empty(x.shape, device=x.device, dtype=torch.float16)
    call("dasa_f32_to_f16", _p(x), _p(out), x.numel(), _stream())
    return out


def to_half_rows(x2, ld, R, C):
    """fp16 copy [R, C] of the leading C columns of R rows with row stride ld (dasa_f32_to_f16_rows)."""
    out = torch.empty(R, C, device=x2.device, dtype=torch.float16)
    call("dasa_f32_to_f16_rows", _p(x2), ld, _p(out), C, R, C, _stream())
    return out


half_gate = True        # AdaIN gate GEMM (and its weight gradient) on fp16 operands in the tensor-core precision mode


def linear_f16_gate(d16, w16, bias, out2, ldo, f2, ldf, gate_out, mask, scale):
    """out2[r, :N] = sigmoid(d16 w16^T + bias) * f2[r, :N] (* mask * scale) on the fp16-operand kernel, fp32 in-place rows."""
    M, K = d16.shape
    N = w16.shape[0]
    e = Epilogue(_p(bias), _p(f2), ldf, _p(gate_out), gate_out.stride(0) if gate_out is not None else 0, _p(mask), float(scale))
    call("dasa_gemm_f16", M, N, K, _p(d16), d16.stride(0), _p(w16), w16.stride(0), _p(out2), ldo, 0, EPI_GATE, ctypes.byref(e),
         _stream())


_half_cache = {}


def half_weight(w):
    """fp16 copy of a (derived) weight tensor, cached while `w` itself is alive and unchanged. `w` is normally the output of
    stacked_weights / transposed_weight, which are rebuilt when a parameter changes, so the identity of `w` is the cache tag."""
    key = id(w)
    # frozen tensors (requires_grad False: the detached transformer stack and weights derived from it) are not touched by the
    # raw-pointer optimizer, so the per-step epoch does not invalidate their copies; an in-place load bumps _version
    tag = (w._version, weights_epoch if w.requires_grad else -1, w.data_ptr())
    hit = _half_cache.get(key)
    if hit is not None and hit[0] == tag and hit[2]() is w:
        return hit[1]
    if len(_half_cache) > 64:
        for k in [k for k, v in _half_cache.items() if v[2]() is None]:
            del _half_cache[k]
    src = w.detach()
    src = src if src.is_contiguous() else src.contiguous()
    out = torch.empty(src.shape, device=src.device, dtype=torch.float16)
    call("dasa_f32_to_f16", _p(src), _p(out), src.numel(), _stream())
    _half_cache[key] = (tag, out, weakref.ref(w))
    return out


def lstm_whh_interleaved(w_hh):
    """fp16 copy of a recurrent LSTM weight [4H, H] with its rows interleaved for the fused cell epilogue of the packed bi-LSTM
    (dasa_lstm_whh_interleave_f16), cached like half_weight (rebuilt after every optimizer step for a trainable weight)."""
    key = ("whh16", id(w_hh))
    tag = (w_hh._version, weights_epoch if w_hh.requires_grad else -1, w_hh.data_ptr())
    hit = _half_cache.get(key)
    if hit is not None and hit[0] == tag and hit[2]() is w_hh:
        return hit[1]
    src = w_hh.detach()
    src = src if src.is_contiguous() else src.contiguous()
    out = torch.empty(src.shape, device=src.device, dtype=torch.float16)
    call("dasa_lstm_whh_interleave_f16", _p(src), _p(out), src.shape[1], _stream())
    _half_cache[key] = (tag, out, weakref.ref(w_hh))
    return out


fused_lstm_cell = True      # packed bi-LSTM forward: cell update in the recurrent GEMM's epilogue (one launch per time step)

_stack_cache = {}


def stacked_weights(ws, dim, pad_to=1, biases=None):
    """Several weights concatenated into one GEMM operand, cached until any of them changes (like transposed_weight):
      dim = 0: rows stacked ([N1+N2(+pad), K]: one projection producing several outputs), zero rows up to a multiple of pad_to;
      dim = 1: columns stacked ([N, K1+K2]: one projection over concatenated inputs).
    `biases` (dim = 0 only): per-weight bias or None -> the matching stacked bias vector (zeros where None / padding).
    Returns (W, bias | None). Gradients are NOT routed through the stack: callers accumulate into the original parameters."""
    key = tuple(id(w) for w in ws) + (dim,)
    tag = tuple((w._version, w.data_ptr()) for w in ws) + (weights_epoch,) + \
        (tuple((b._version, b.data_ptr()) if b is not None else None for b in biases) if biases else ())
    hit = _stack_cache.get(key)
    if hit is not None and hit[0] == tag and all(r() is w for r, w in zip(hit[3], ws)):
        return hit[1], hit[2]
    if len(_stack_cache) > 64:
        for k in [k for k, v in _stack_cache.items() if any(r() is None for r in v[3])]:
            del _stack_cache[k]
    with torch.no_grad():
        if dim == 1:
            W = torch.cat([w.detach() for w in ws], 1).contiguous()
            bias = None
        else:
            n = sum(w.shape[0] for w in ws)
            npad = (n + pad_to - 1) // pad_to * pad_to
            W = torch.zeros(npad, ws[0].shape[1], device=ws[0].device, dtype=torch.float32)
            bias = torch.zeros(npad, device=ws[0].device, dtype=torch.float32) if biases else None
            off = 0
            for i, w in enumerate(ws):
                W[off:off + w.shape[0]].copy_(w.detach())
                if biases and biases[i] is not None:
                    bias[off:off + w.shape[0]].copy_(biases[i].detach())
                off += w.shape[0]
    _stack_cache[key] = (tag, W, bias, [weakref.ref(w) for w in ws])
    return W, bias


def _tc_ok(*dims):
    return _precision == PREC_TF32 and all(d >= 32 and d % 4 == 0 for d in dims)


def linear_bwd_input(dy, w, out=None, beta=0.0, precision=None):
    """dx[.., K] (+)= dy[.., N] @ w[N, K]"""
    _chk(dy, w, out)
    d2, M, N, lda = _rows(dy)
    K = w.shape[1]
    if out is None:
        out = torch.empty(*dy.shape[:-1], K, device=dy.device, dtype=torch.float32)
    o2, _, _, ldc = _rows_out(out)
    if precision is None and _tc_ok(N) and lda % 4 == 0 and d2.data_ptr() % 16 == 0:
        if use_mn_major and K % 4 == 0 and lib.load().dasa_gemm_layout_on_tensor_cores(1, 0, M, K, N):
            gemm(d2, lda, 1, w, K, 0, o2, ldc, M, K, N, beta=beta)  # W read in place as an MN-major B operand
            return out
        wt = transposed_weight(w)                                  # [K, N] rows of stride Np
        gemm(d2, lda, 1, wt, wt.stride(0), 1, o2, ldc, M, K, N, beta=beta)
        return out
    gemm(d2, lda, 1, w, K, 0, o2, ldc, M, K, N, beta=beta, precision=precision)
    return out


def linear_bwd_weight(dy, x, dw, accumulate=True, precision=None):
    """dw[N, K] (+)= dy[.., N]^T @ x[.., K]   (reduction over the rows)"""
    _chk(dy, x, dw)
    d2, M, N, ldd = _rows(dy)
    x2, Mx, K, ldx = _rows(x)
    assert M == Mx and dw.shape == (N, K) and dw.is_contiguous()
    if precision is None and _tc_ok(M) and M >= 64:
        if (use_mn_major and ldd % 4 == 0 and ldx % 4 == 0 and d2.data_ptr() % 16 == 0 and x2.data_ptr() % 16 == 0 and
                lib.load().dasa_gemm_layout_on_tensor_cores(0, 0, N, K, M)):
            # dY [M, N] and X [M, K] are MN-major operands of dW = dY^T X as they stand: no transposed copies
            gemm(d2, ldd, 0, x2, ldx, 0, dw, K, N, K, M, beta=1.0 if accumulate else 0.0)
            return dw
        dyt, xt = transpose(d2), transpose(x2)                    # [N, Mp], [K, Mp]: both K-major over the row index
        gemm(dyt, dyt.stride(0), 1, xt, xt.stride(0), 1, dw, K, N, K, M, beta=1.0 if accumulate else 0.0)
        return dw
    gemm(d2, ldd, 0, x2, ldx, 0, dw, K, N, K, M, beta=1.0 if accumulate else 0.0, precision=precision)
    return dw


def linear_bwd_weight_f16(dy16, x16, dw, alpha=1.0, accumulate=True):
    """dw[N, K] (+)= alpha * dy16[M, N]^T @ x16[M, K] with fp16 operands read where they lie (both MN-major: dasa_gemm_f16_mn)."""
    assert dy16.dtype == torch.float16 and x16.dtype == torch.float16 and dy16.dim() == 2 and x16.dim() == 2
    assert dy16.stride(1) == 1 and x16.stride(1) == 1 and dy16.shape[0] == x16.shape[0]
    M, N = dy16.shape
    K = x16.shape[1]
    assert dw.shape == (N, K) and dw.is_contiguous() and dw.dtype == torch.float32
    nbytes = lib.load().dasa_gemm_workspace_bytes(N, K, M, PREC_TF32)
    ws = workspace(nbytes) if nbytes else None
    call("dasa_gemm_f16_mn", N, K, M, float(alpha), _p(dy16), dy16.stride(0), _p(x16), x16.stride(0), 1.0 if accumulate else 0.0,
         _p(dw), K, _p(ws), ws.numel() if ws is not None else 0, _stream())
    return dw


def colsum(x, out, accumulate=True):
    x2, M, N, ld = _rows(x)
    call("dasa_colsum", _p(x2), ld, M, N, _p(out), int(accumulate), _stream())
    return out


def colsum_h(x16, scale, out, accumulate=True):
    """out[n] (+)= scale * sum_m x16[m, n] over an fp16 matrix (dasa_colsum_h)."""
    assert x16.dtype == torch.float16 and x16.dim() == 2 and x16.stride(1) == 1
    call("dasa_colsum_h", _p(x16), x16.stride(0), x16.shape[0], x16.shape[1], float(scale), _p(out), int(accumulate), _stream())
    return out


def _acc_grad(param, fn_new):
    """Accumulate a parameter gradient in place (param.grad is created zero-filled on first use)."""
    if param.grad is None:
        param.grad = torch.zeros_like(param)
    fn_new(param.grad)


# ------------------------------------------------------------------------------------------------ elementwise
def dropout_apply(x, mask, scale, out=None):
    x2, R, C, ldx = _rows(x)
    if out is None:
        out = torch.empty(x.shape, device=x.device, dtype=torch.float32)
    o2, _, _, ldo = _rows_out(out)
    call("dasa_dropout_apply", _p(x2), ldx, _p(mask), float(scale), _p(o2), ldo, R, C, _stream())
    return out


def act_backward(act, dy, y, mask=None, scale=1.0):
    d2, R, C, ldd = _rows(dy)
    y2, _, _, ldy = _rows(y)
    dx = torch.empty(dy.shape, device=dy.device, dtype=torch.float32)
    x2, _, _, ldx = _rows_out(dx)
    call("dasa_act_backward", {"tanh": 0, "relu": 1, "gelu": 2}[act], _p(d2), ldd, _p(y2), ldy, _p(mask), float(scale), _p(x2), ldx, R, C,
         _stream())
    return dx


def gelu_fwd(x):
    x = x.contiguous()
    y = torch.empty_like(x)
    call("dasa_gelu_fwd", _p(x), _p(y), x.numel(), _stream())
    return y


def axpy2d(a, x, y, accumulate=True):
    x2, R, C, ldx = _rows(x)
    y2, _, _, ldy = _rows_out(y)
    call("dasa_axpy2d", float(a), _p(x2), ldx, _p(y2), ldy, int(accumulate), R, C, _stream())
    return y


def dropout_mask(shape, p, seed, offset=0, device="cuda"):
    m = torch.empty(shape, dtype=torch.uint8, device=device)
    call("dasa_dropout_mask", _p(m), m.numel(), float(p), int(seed), int(offset), _stream())
    return m


def dropout_mask_dev(shape, p, seed_dev, offset=0, device="cuda"):
    """Keep mask whose seed is read from device memory at run time (fresh masks on every CUDA-graph replay)."""
    m = torch.empty(shape, dtype=torch.uint8, device=device)
    call("dasa_dropout_mask_dev", _p(m), m.numel(), float(p), _p(seed_dev), int(offset), _stream())
    return m


def bump_counter(counter, inc=0x9E3779B97F4A7C15):
    call("dasa_bump_counter", _p(counter), int(inc) & 0xFFFFFFFFFFFFFFFF, _stream())


def as_keep_mask(m):
    """Accept bool / float (pre-scaled) / uint8 masks from tests; return contiguous uint8 keep flags."""
    if m is None:
        return None
    if m.dtype != torch.uint8:
        m = (m != 0).to(torch.uint8)
    return m.contiguous()


# -------------------------------------------------------------------------------------------------- AdaIN family
def gate_modulate(g, f, out, mask=None, scale=1.0):
    g2, R, C, ldg = _rows(g)
    f2, _, _, ldf = _rows(f)
    o2, _, _, ldo = _rows_out(out)
    call("dasa_gate_modulate", _p(g2), ldg, _p(f2), ldf, _p(o2), ldo, _p(mask), float(scale), R, C, _stream())
    return out


def view_stats(d):
    """d [N, V, C] (rows may be strided) -> [N, 4C] mean/std/max/min over views."""
    N, V, C = d.shape
    assert d.stride(2) == 1
    stats = torch.empty(N, 4 * C, device=d.device, dtype=torch.float32)
    call("dasa_view_stats", _p(d), d.stride(1), d.stride(0), N, V, C, _p(stats), _stream())
    return stats


def channel_modulate(f, a, b, out=None):
    N, V, C = f.shape
    assert f.stride(2) == 1
    if out is None:
        out = torch.empty(N, V, C, device=f.device, dtype=torch.float32)
    call("dasa_channel_modulate", _p(f), f.stride(1), f.stride(0), _p(a.contiguous()), _p(None if b is None else b.contiguous()),
         _p(out), out.stride(1), out.stride(0), N, V, C, _stream())
    return out


def channel_modulate_bwd(dout, f, a, want_df=False, want_db=True):
    """Backward of channel_modulate in one pass over dout and f: returns (df | None, da [N, C], db [N, C] | None)."""
    N, V, C = f.shape
    assert f.stride(2) == 1 and dout.stride(2) == 1 and dout.shape == f.shape
    da = torch.empty(N, C, device=f.device, dtype=torch.float32)
    db = torch.empty(N, C, device=f.device, dtype=torch.float32) if want_db else None
    df = torch.empty(N, V, C, device=f.device, dtype=torch.float32) if want_df else None
    call("dasa_channel_modulate_bwd", _p(dout), dout.stride(1), dout.stride(0), _p(f), f.stride(1), f.stride(0),
         _p(a.contiguous()), _p(da), _p(db), _p(df), 0 if df is None else df.stride(1), 0 if df is None else df.stride(0),
         N, V, C, _stream())
    return df, da, db


def adain_rows(f, d, eps=1e-5, out=None):
    f2, R, C, ldf = _rows(f)
    d2, _, _, ldd = _rows(d)
    if out is None:
        out = torch.empty(f.shape, device=f.device, dtype=torch.float32)
    o2, _, _, ldo = _rows_out(out)
    call("dasa_adain_rows", _p(f2), ldf, _p(d2), ldd, _p(o2), ldo, R, C, float(eps), _stream())
    return out


# --------------------------------------------------------------------------------------------- attention kernels
def row_attention_fwd(ctx, t, mask=None, shift_k=0, headings=12, kappa_logits=None, wc=None, want_q=False):
    B, rows, D = ctx.shape
    assert ctx.stride(2) == 1
    if wc is None:
        wc = torch.empty(B, D, device=ctx.device, dtype=torch.float32)
    attn = torch.empty(B, rows, device=ctx.device, dtype=torch.float32)
    q = torch.empty(B, rows, device=ctx.device, dtype=torch.float32) if (want_q or shift_k > 0) else None
    kappa = torch.empty(B, shift_k, device=ctx.device, dtype=torch.float32) if shift_k > 0 else None
    if mask is not None:
        mask = mask.to(torch.uint8) if mask.dtype != torch.uint8 else mask
        assert mask.stride(1) == 1
    call("dasa_row_attention_fwd", _p(ctx), ctx.stride(1), ctx.stride(0), B, rows, D, _p(t), t.stride(0), _p(mask),
         mask.stride(0) if mask is not None else 0, shift_k, headings, _p(kappa_logits),
         kappa_logits.stride(0) if kappa_logits is not None else 0, _p(wc), wc.stride(0), _p(attn), _p(q), _p(kappa), _stream())
    return wc, attn, q, kappa


def gate_shift_attention_fwd(f, gate_pre, t, kappa_logits, shift_k=5, headings=12, chan_scale=None):
    """Fused DGAdaChannel gate -> ShiftSoftDotAttention (dasa_gate_shift_attention_fwd): f [B, rows, D] raw features, gate_pre
    [B, rows, C] = a_fc(d) pre-activations (C <= D gated channels), t [B, D] = linear_in(h), kappa_logits [B, k]. Returns
    (weighted context [B, D], pre-shift softmax [B, rows], shifted weights q, kappa) of the MODULATED features, which are never
    written to HBM. Forward only."""
    B, rows, D = f.shape
    C = gate_pre.shape[2]
    assert f.stride(2) == 1 and gate_pre.stride(2) == 1 and tuple(gate_pre.shape[:2]) == (B, rows)
    wc = torch.empty(B, D, device=f.device, dtype=torch.float32)
    attn = torch.empty(B, rows, device=f.device, dtype=torch.float32)
    q = torch.empty(B, rows, device=f.device, dtype=torch.float32)
    kappa = torch.empty(B, shift_k, device=f.device, dtype=torch.float32)
    call("dasa_gate_shift_attention_fwd", _p(f), f.stride(1), f.stride(0), B, rows, D, _p(gate_pre), gate_pre.stride(1),
         gate_pre.stride(0), C, _p(chan_scale), _p(t), t.stride(0), shift_k, headings, _p(kappa_logits), kappa_logits.stride(0),
         _p(wc), wc.stride(0), _p(attn), _p(q), _p(kappa), _stream())
    return wc, attn, q, kappa


def row_attention_bwd(ctx, t, attn, q, kappa, dwc, shift_k=0, headings=12, need_dctx=True, dctx=None, accumulate=False,
                      dt=None, dkl=None):
    """dt / dkl (optional): caller-owned output views (e.g. column ranges of one buffer that feeds a single stacked GEMM)."""
    B, rows, D = ctx.shape
    if need_dctx and dctx is None:
        dctx = torch.empty(B, rows, D, device=ctx.device, dtype=torch.float32)
    if dt is None:
        dt = torch.empty(B, D, device=ctx.device, dtype=torch.float32)
    if dkl is None:
        dkl = torch.empty(B, shift_k, device=ctx.device, dtype=torch.float32) if shift_k > 0 else None
    assert dt.stride(1) == 1 and (dkl is None or dkl.stride(1) == 1)
    assert dwc.stride(1) == 1
    call("dasa_row_attention_bwd", _p(ctx), ctx.stride(1), ctx.stride(0), B, rows, D, _p(t), t.stride(0), _p(attn), _p(q),
         _p(kappa), shift_k, headings, _p(dwc), dwc.stride(0), _p(dctx), dctx.stride(1) if dctx is not None else 0,
         dctx.stride(0) if dctx is not None else 0, int(accumulate), _p(dt), dt.stride(0), _p(dkl),
         dkl.stride(0) if dkl is not None else 0, _stream())
    return dctx, dt, dkl


def cand_logits_fwd(cand, t, leng):
    B, Nc, D = cand.shape
    assert cand.stride(2) == 1
    logit = torch.empty(B, Nc, device=cand.device, dtype=torch.float32)
    call("dasa_cand_logits_fwd", _p(cand), cand.stride(1), cand.stride(0), B, Nc, D, _p(t), t.stride(0), _p(leng), _p(logit),
         _stream())
    return logit


def cand_logits_bwd(cand, t, leng, dlogit, Dc, need_dcand=True):
    B, Nc, D = cand.shape
    dcand = torch.empty(B, Nc, Dc, device=cand.device, dtype=torch.float32) if need_dcand else None
    dt = torch.empty(B, D, device=cand.device, dtype=torch.float32)
    call("dasa_cand_logits_bwd", _p(cand), cand.stride(1), cand.stride(0), B, Nc, D, _p(t), t.stride(0), _p(leng),
         _p(dlogit.contiguous()), _p(dcand), dcand.stride(1) if need_dcand else 0, dcand.stride(0) if need_dcand else 0, Dc,
         _p(dt), dt.stride(0), _stream())
    return dcand, dt


# -------------------------------------------------------------------------------------------------------- LSTM
def lstm_pointwise_fwd(ga, gb, bias_a, bias_b, c_prev, h_prev, h_out, c_out, seq_out=None, acts_out=None, active=None, pos=0,
                       seq_mask=None, seq_scale=1.0):
    B, H = h_out.shape

    def ld(t):
        return 0 if t is None else t.stride(0)
    call("dasa_lstm_pointwise_fwd", _p(ga), ld(ga), _p(gb), ld(gb), _p(bias_a), _p(bias_b), _p(c_prev), ld(c_prev), _p(h_prev),
         ld(h_prev), _p(h_out), ld(h_out), _p(c_out), ld(c_out), _p(seq_out), ld(seq_out), _p(acts_out), ld(acts_out),
         _p(active), int(pos), B, H, _p(seq_mask), float(seq_scale), _stream())


def lstm_pointwise_bwd(dh, dh2, dc, acts, c_prev, c_new, dgates, dc_prev, dh_pass=None, active=None, pos=0, dh2_mask=None,
                       dh2_scale=1.0):
    B, H = c_new.shape

    def ld(t):
        return 0 if t is None else t.stride(0)
    call("dasa_lstm_pointwise_bwd", _p(dh), ld(dh), _p(dh2), ld(dh2), _p(dc), ld(dc), _p(acts), ld(acts), _p(c_prev), ld(c_prev),
         _p(c_new), ld(c_new), _p(dgates), ld(dgates), _p(dc_prev), ld(dc_prev), _p(dh_pass), ld(dh_pass), _p(active), int(pos),
         B, H, _p(dh2_mask), float(dh2_scale), _stream())


# ------------------------------------------------------------------------------------------------ encoder pieces
def embed_layernorm(ids, word, pos, type0, gamma, beta, eps, mask=None, scale=1.0):
    B, L = ids.shape
    Hd = word.shape[1]
    assert ids.dtype == torch.int64 and ids.stride(1) == 1
    out = torch.empty(B, L, Hd, device=word.device, dtype=torch.float32)
    call("dasa_embed_layernorm", _p(ids), ids.stride(0), B, L, Hd, _p(word), _p(pos), _p(type0), _p(gamma), _p(beta), float(eps),
         _p(mask), float(scale), _p(out), _stream())
    return out


def dropout_residual_layernorm(x, resid, gamma, beta, eps, mask=None, scale=1.0, post_mask=None, post_scale=1.0,
                               save=False, half_copy=False):
    """half_copy: also return an fp16 copy of the output ([R, Hd] contiguous, the A operand of the next fp16 GEMM)."""
    x2, R, Hd, ldx = _rows(x)
    out = torch.empty(x.shape, device=x.device, dtype=torch.float32)
    out16 = torch.empty(x.shape, device=x.device, dtype=torch.float16) if half_copy else None
    o2, _, _, ldo = _rows_out(out)
    r2, ldr = None, 0
    if resid is not None:
        r2, _, _, ldr = _rows(resid)
    stats = torch.empty(R, 2, device=x.device, dtype=torch.float32) if save else None
    z = torch.empty(R, Hd, device=x.device, dtype=torch.float32) if save else None
    call("dasa_dropout_residual_layernorm", _p(x2), ldx, _p(mask), float(scale), _p(r2), ldr, _p(gamma), _p(beta), float(eps),
         _p(post_mask), float(post_scale), _p(o2), ldo, _p(stats), _p(z), _p(out16), R, Hd, _stream())
    if half_copy:
        return out, out16
    return (out, stats, z) if save else out


def layernorm_bwd(dout, z, gamma, stats, dgamma, dbeta, mask=None, scale=1.0, post_mask=None, post_scale=1.0, need_dx=True):
    d2, R, Hd, ldd = _rows(dout)
    dresid = torch.empty(dout.shape, device=dout.device, dtype=torch.float32)
    dx = torch.empty(dout.shape, device=dout.device, dtype=torch.float32) if (need_dx and mask is not None) else None
    call("dasa_layernorm_bwd", _p(d2), ldd, _p(z), _p(gamma), _p(stats), _p(mask), float(scale), _p(post_mask), float(post_scale),
         _p(dx), Hd, _p(dresid), Hd, _p(dgamma), _p(dbeta), R, Hd, _stream())
    return (dx if dx is not None else dresid), dresid


def mha_fwd(q, k, v, heads, key_pad=None, drop_mask=None, drop_scale=1.0, save_probs=False, out_half=False):
    """q [B,Lq,Hd] / k,v [B,Lk,Hd] views with unit inner stride (slices of a fused QKV buffer are fine). out_half: the context
    is written as fp16 (forward-only; the A operand of the fp16 output projection)."""
    B, Lq, Hd = q.shape
    Lk = k.shape[1]
    dh = Hd // heads
    out = torch.empty(B, Lq, Hd, device=q.device, dtype=torch.float16 if out_half else torch.float32)
    probs = torch.empty(B, heads, Lq, Lk, device=q.device, dtype=torch.float32) if save_probs else None
    if key_pad is not None and key_pad.dtype != torch.uint8:
        key_pad = key_pad.to(torch.uint8)
    call("dasa_mha_fwd", _p(q), q.stride(1), q.stride(0), _p(k), k.stride(1), k.stride(0), _p(v), v.stride(1), v.stride(0),
         _p(key_pad), key_pad.stride(0) if key_pad is not None else 0, _p(drop_mask), float(drop_scale), _p(out), out.stride(1),
         out.stride(0), _p(probs), B, heads, Lq, Lk, dh, _precision, int(out_half), _stream())
    return (out, probs) if save_probs else out


def mha_fwd_varlen(q, k, v, heads, q_pack=None, k_pack=None, max_lq=None, max_lk=None, drop_mask=None, drop_scale=1.0,
                   out_half=False):
    """Attention over packed operands. q: [Nq, Hd] packed (q_pack = (off, len) int32 tensors) or dense [B, Lq, Hd] (q_pack None);
    k, v likewise. Returns out laid out like q."""
    Hd = q.shape[-1]
    dh = Hd // heads
    if q_pack is not None:
        B = q_pack[0].numel()
        out = torch.empty(q.shape[0], Hd, device=q.device, dtype=torch.float16 if out_half else torch.float32)
        ldq, ldo, sq = q.stride(0), Hd, 0
    else:
        B, max_lq = q.shape[0], q.shape[1]
        out = torch.empty(B, max_lq, Hd, device=q.device, dtype=torch.float16 if out_half else torch.float32)
        ldq, ldo, sq = q.stride(1), Hd, q.stride(0)
    if k_pack is not None:
        ldk, ldv, skv = k.stride(0), v.stride(0), 0
    else:
        max_lk = k.shape[1]
        ldk, ldv, skv = k.stride(1), v.stride(1), k.stride(0)
        assert v.stride(0) == k.stride(0)
    call("dasa_mha_fwd_varlen", _p(q), ldq, _p(q_pack[0]) if q_pack else None, _p(q_pack[1]) if q_pack else None, _p(k), ldk,
         _p(v), ldv, _p(k_pack[0]) if k_pack else None, _p(k_pack[1]) if k_pack else None, sq, skv, _p(drop_mask),
         float(drop_scale), _p(out), ldo, B, heads, int(max_lq), int(max_lk), dh, _precision, int(out_half), _stream())
    return out


class DropStream:
    """Dropout keep flags drawn in place by the consuming kernel (csrc/rng.cuh) instead of read from a mask tensor: stream
    byte e = 16-bit lane e & 3 of hash(seed, base + e // 4) >= p * 65536. `dropout_mask((n,), p, seed, base)` (n % 16 == 0)
    materialises the same bytes."""
    __slots__ = ("seed_dev", "seed", "base", "p")

    def __init__(self, seed_dev, seed, base, p):
        self.seed_dev, self.seed, self.base, self.p = seed_dev, int(seed), int(base), float(p)


def _drop_args(drop):
    """(mask ptr, seed_dev ptr, seed, base, p) of a keep-mask tensor, a DropStream or None."""
    if drop is None:
        return None, None, 0, 0, 0.0
    if isinstance(drop, DropStream):
        return None, _p(drop.seed_dev), drop.seed & 0xFFFFFFFFFFFFFFFF, drop.base, drop.p
    assert drop.dtype == torch.uint8 and drop.is_contiguous()
    return _p(drop), None, 0, 0, 0.0


def mha_h16_stream_bytes(B, heads, max_lq, max_lk):
    """Stream bytes dasa_mha_fwd_h16 consumes for its in-place probability dropout."""
    return B * heads * ((max_lq + 15) // 16) * (2 * ((max_lk + 15) // 16)) * 128


def mha_h16_stream_index(B, heads, max_lq, max_lk, device="cpu"):
    """int64 [B, heads, max_lq, max_lk]: the stream byte holding the keep flag of probability [b, h, r, j] (include/dasa_b200.h:
    dasa_mha_fwd_h16). Tests use it to materialise the mask the kernel draws."""
    nMT, nNT = (max_lq + 15) // 16, 2 * ((max_lk + 15) // 16)
    b = torch.arange(B, device=device).view(B, 1, 1, 1)
    h = torch.arange(heads, device=device).view(1, heads, 1, 1)
    r = torch.arange(max_lq, device=device).view(1, 1, max_lq, 1)
    j = torch.arange(max_lk, device=device).view(1, 1, 1, max_lk)
    lane = (r % 8) * 4 + (j % 8) // 2
    tile = ((b * heads + h) * nMT + r // 16) * nNT + j // 8
    return (tile * 32 + lane) * 4 + 2 * (j % 2) + (r % 16) // 8


def mha_fwd_h16(q, k, v, heads, q_pack=None, k_pack=None, max_lq=None, max_lk=None, key_pad=None, drop=None, drop_scale=1.0,
                out_half=True):
    """Forward-only attention on fp16 q / k / v (dasa_mha_fwd_h16). q: [Nq, Hd] packed (q_pack = (off, len)) or dense
    [B, Lq, Hd]; k, v likewise (slices of a fused fp16 QKV buffer are fine). drop: uint8 keep mask [B, heads, max_lq, max_lk],
    an ops.DropStream (flags drawn in the kernel) or None. Returns the context laid out like q (fp16, or fp32)."""
    assert q.dtype == torch.float16 and k.dtype == torch.float16 and v.dtype == torch.float16
    Hd = q.shape[-1]
    dh = Hd // heads
    odt = torch.float16 if out_half else torch.float32
    if q_pack is not None:
        B = q_pack[0].numel()
        out = torch.empty(q.shape[0], Hd, device=q.device, dtype=odt)
        ldq, ldo, sq, so = q.stride(0), Hd, 0, 0
    else:
        B, max_lq = q.shape[0], q.shape[1]
        out = torch.empty(B, max_lq, Hd, device=q.device, dtype=odt)
        ldq, ldo, sq, so = q.stride(1), Hd, q.stride(0), max_lq * Hd
    if k_pack is not None:
        ldk, ldv, skv = k.stride(0), v.stride(0), 0
    else:
        max_lk = k.shape[1]
        ldk, ldv, skv = k.stride(1), v.stride(1), k.stride(0)
        assert v.stride(0) == k.stride(0)
    if key_pad is not None and key_pad.dtype != torch.uint8:
        key_pad = key_pad.to(torch.uint8)
    mp, sp, seed, base, p = _drop_args(drop)
    call("dasa_mha_fwd_h16", _p(q), ldq, sq, _p(q_pack[0]) if q_pack else None, _p(q_pack[1]) if q_pack else None, _p(k), ldk,
         _p(v), ldv, skv, _p(k_pack[0]) if k_pack else None, _p(k_pack[1]) if k_pack else None, _p(key_pad),
         key_pad.stride(0) if key_pad is not None else 0, mp, sp, seed, base, p, float(drop_scale), _p(out), ldo, so,
         int(out_half), B, heads, int(max_lq), int(max_lk), dh, _stream())
    return out


def dropout_residual_layernorm_fwd(x, resid, gamma, beta, eps, drop=None, scale=1.0, half_copy=False):
    """Forward-only dropout -> + resid -> LayerNorm (dasa_dropout_residual_layernorm_fwd): x fp32 or fp16 (the c_half output of
    the fp16 GEMM), drop = uint8 keep mask shaped like x, an ops.DropStream or None. Returns out fp32 (and its fp16 copy)."""
    x2, R, Hd, ldx = _rows(x)
    out = torch.empty(x.shape, device=x.device, dtype=torch.float32)
    out16 = torch.empty(x.shape, device=x.device, dtype=torch.float16) if half_copy else None
    o2, _, _, ldo = _rows_out(out)
    r2, ldr = None, 0
    if resid is not None:
        r2, _, _, ldr = _rows(resid)
    mp, sp, seed, base, p = _drop_args(drop)
    call("dasa_dropout_residual_layernorm_fwd", _p(x2), int(x.dtype == torch.float16), ldx, mp, sp, seed, base, p, float(scale),
         _p(r2), ldr, _p(gamma), _p(beta), float(eps), _p(o2), ldo, _p(out16), R, Hd, _stream())
    return (out, out16) if half_copy else out


def gather_rows(src2d, idx_i32, out=None):
    R, C = idx_i32.numel(), src2d.shape[1]
    if out is None:
        out = torch.empty(R, C, device=src2d.device, dtype=torch.float32)
    call("dasa_gather_rows", _p(src2d), src2d.stride(0), _p(idx_i32), _p(out), out.stride(0), R, C, _stream())
    return out


def reverse_tokens_packed(x_packed, offsets_i32, lengths_i32, L):
    B, Hd = lengths_i32.numel(), x_packed.shape[1]
    out = torch.empty(B, L, Hd, device=x_packed.device, dtype=torch.float32)
    call("dasa_reverse_tokens_packed", _p(x_packed), _p(offsets_i32), _p(lengths_i32), _p(out), B, L, Hd, _stream())
    return out


def mha_bwd(q, k, v, probs, dout, heads, drop_mask=None, drop_scale=1.0):
    B, Lq, Hd = q.shape
    Lk = k.shape[1]
    dh = Hd // heads
    dq = torch.empty(B, Lq, Hd, device=q.device, dtype=torch.float32)
    dk = torch.empty(B, Lk, Hd, device=q.device, dtype=torch.float32)
    dv = torch.empty(B, Lk, Hd, device=q.device, dtype=torch.float32)
    dout = dout.contiguous()
    call("dasa_mha_bwd", _p(q), q.stride(1), q.stride(0), _p(k), k.stride(1), k.stride(0), _p(v), v.stride(1), v.stride(0),
         _p(probs), _p(drop_mask), float(drop_scale), _p(dout), dout.stride(1), dout.stride(0), _p(dq), dq.stride(1), dq.stride(0),
         _p(dk), dk.stride(1), dk.stride(0), _p(dv), dv.stride(1), dv.stride(0), B, heads, Lq, Lk, dh, _stream())
    return dq, dk, dv


def reverse_tokens(x, lengths_i32):
    B, L, Hd = x.shape
    x = x.contiguous()
    out = torch.empty_like(x)
    call("dasa_reverse_tokens", _p(x), _p(out), _p(lengths_i32), B, L, Hd, _stream())
    return out


# -------------------------------------------------------------------------------------------------- loss / optim
def masked_ce(logit, target, ignore_index, grad_scale, loss_acc, want_grad=True, want_action=True, want_stats=False):
    B, Nc = logit.shape
    assert logit.stride(1) == 1
    dlogit = torch.empty(B, Nc, device=logit.device, dtype=torch.float32) if want_grad else None
    action = torch.empty(B, device=logit.device, dtype=torch.int64) if want_action else None
    lp = torch.empty(B, device=logit.device, dtype=torch.float32) if want_stats else None
    ent = torch.empty(B, device=logit.device, dtype=torch.float32) if want_stats else None
    call("dasa_masked_ce", _p(logit), logit.stride(0), _p(target), int(ignore_index), B, Nc, float(grad_scale), _p(loss_acc),
         _p(dlogit), _p(action), _p(lp), _p(ent), _stream())
    return dlogit, action, lp, ent


def policy_sample_fwd(logit, u=None, action_in=None, want_probs=True):
    """Categorical(softmax(logit)): (action, log_prob(action), entropy, probs). action = action_in | sample(u) | argmax."""
    B, Nc = logit.shape
    assert logit.stride(1) == 1
    dev = logit.device
    action = torch.empty(B, device=dev, dtype=torch.int64)
    lp = torch.empty(B, device=dev, dtype=torch.float32)
    ent = torch.empty(B, device=dev, dtype=torch.float32)
    probs = torch.empty(B, Nc, device=dev, dtype=torch.float32) if want_probs else None
    call("dasa_policy_sample_fwd", _p(logit), logit.stride(0), B, Nc, _p(u), _p(action_in), _p(action), _p(lp), _p(ent),
         _p(probs), _stream())
    return action, lp, ent, probs


def policy_sample_bwd(probs, action, dlogp, dent, entropy):
    B, Nc = probs.shape
    dlogit = torch.empty(B, Nc, device=probs.device, dtype=torch.float32)
    call("dasa_policy_sample_bwd", _p(probs), _p(action), _p(dlogp), _p(dent), _p(entropy), B, Nc, _p(dlogit), Nc, _stream())
    return dlogit


def nav_reward(action, cand_leng, ignore_id, dist, last_dist, ended, reward, mask):
    call("dasa_nav_reward", _p(action), _p(cand_leng), int(ignore_id), _p(dist), _p(last_dist), _p(ended), _p(reward), _p(mask),
         action.numel(), _stream())


def a2c_loss(logp, ent, value, last_value, reward, mask, ended, gamma, ent_coef, normalize):
    """Returns (loss[1], total[1], dlogp, dent, dvalue) for [T,B] stacks (contiguous)."""
    T, B = logp.shape
    dev = logp.device
    loss = torch.empty(1, device=dev, dtype=torch.float32)
    total = torch.empty(1, device=dev, dtype=torch.float32)
    dlogp = torch.empty(T, B, device=dev, dtype=torch.float32)
    dvalue = torch.empty(T, B, device=dev, dtype=torch.float32)
    dent = torch.empty(T, B, device=dev, dtype=torch.float32) if ent is not None else None
    call("dasa_a2c_loss", _p(logp), _p(ent), _p(value), _p(last_value), _p(reward), _p(mask), _p(ended), float(gamma),
         float(ent_coef), {"none": 0, "total": 1, "batch": 2}[normalize], T, B, _p(loss), _p(total), _p(dlogp), _p(dent),
         _p(dvalue), _stream())
    return loss, total, dlogp, dent, dvalue


def rmsprop_step(param, grad, square_avg, lr, alpha=0.99, eps=1e-8, weight_decay=0.0, clip_coef=None, lr_scale=None):
    call("dasa_rmsprop_step", _p(param), _p(grad), _p(square_avg), param.numel(), float(lr), float(alpha), float(eps),
         float(weight_decay), _p(clip_coef), _p(lr_scale), _stream())


def lr_lambda(iter_dev, mult_dev, warm_steps, decay_start, decay_intervals, lr_decay, advance=1):
    """mult_dev[0] = lr_lambda(iter_dev[0]); iter_dev[0] += advance (agent_dg.py:219-227, on the device)."""
    assert iter_dev.dtype == torch.int32 and mult_dev.dtype == torch.float32
    call("dasa_lr_lambda", _p(iter_dev), int(warm_steps), int(decay_start), int(decay_intervals), float(lr_decay), _p(mult_dev),
         int(advance), _stream())


def sumsq(x, out):
    call("dasa_sumsq", _p(x), x.numel(), _p(out), _stream())


def clip_coef(sumsq_t, max_norm, coef):
    call("dasa_clip_coef", _p(sumsq_t), float(max_norm), _p(coef), _stream())


# ---------------------------------------------------------------------- persistent decoder rollout (csrc/decoder_persist.cu)
persistent_decoder = True      # route the decoder through the cooperative whole-rollout kernel when it applies (TF32 precision)


def decoder_rollout_supported(B, H, E, F, V, L, D, NK, shift_k):
    """True when dasa_decoder_rollout_fwd/_bwd take this geometry (B <= 32, shared-memory plan fits) AND the tensor-core precision
    is selected: the kernel multiplies in TF32, so the exact-fp32 mode keeps the per-op FFMA path."""
    if not persistent_decoder or _precision != PREC_TF32:
        return False
    return bool(lib.load().dasa_decoder_rollout_supported(B, H, E, F, V, L, D, NK, shift_k))


def _ld3(t):
    """(row, sample, action) strides of a [T, B, R, C] tensor with unit inner stride."""
    assert t.dim() == 4 and t.stride(3) == 1, (t.shape, t.stride())
    return t.stride(2), t.stride(1), t.stride(0)


def decoder_rollout_fwd(emb, feat, ctx, ctx_mask, h0, c0, m_hprev, m_h1, scale, w_feat, b_feat, w_lstm, b_ih, b_hh, w_att_in,
                        w_att_out, headings, shift_k):
    """emb [T,B,E], feat [T,B,V,F], ctx [T,B,L,D] -> dict of every [T, B, .] buffer dasa_decoder_fwd_t writes."""
    _chk(emb, feat, ctx, h0, c0, w_feat, b_feat, w_lstm, b_ih, b_hh, w_att_in, w_att_out)
    T, B, E = emb.shape
    V, F = feat.shape[2], feat.shape[3]
    L, D = ctx.shape[2], ctx.shape[3]
    H = h0.shape[1]
    NK = w_feat.shape[0]
    dev = emb.device
    KX, DC = E + F + H, D + H
    assert w_lstm.shape == (4 * H, KX) and w_att_in.shape == (D, H) and w_att_out.shape == (H, DC) and w_feat.shape[1] == H

    def new(*shape):
        return torch.empty(*shape, device=dev, dtype=torch.float32)
    o = {"hprev_drop": new(T, B, H), "tk": new(T, B, NK), "p": new(T, B, V), "q": new(T, B, V), "kappa": new(T, B, shift_k),
         "xh": new(T, B, KX), "acts": new(T, B, 4 * H), "c": new(T + 1, B, H), "h1": new(T, B, H), "cat": new(T, B, DC),
         "t2": new(T, B, D), "alpha": new(T, B, L), "htilde": new(T, B, H)}
    zpart = new(lib.load().dasa_decoder_rollout_scratch_floats(B))
    barrier = torch.empty(32, device=dev, dtype=torch.int32)
    emb, h0, c0 = emb.contiguous(), h0.contiguous(), c0.contiguous()
    if ctx_mask is not None:
        ctx_mask = ctx_mask if ctx_mask.dtype == torch.uint8 else ctx_mask.to(torch.uint8)
        assert ctx_mask.stride(1) == 1
    a = lib.DecoderFwd()
    a.T, a.B, a.H, a.E, a.F, a.V, a.L, a.D, a.headings, a.shift_k, a.NK = T, B, H, E, F, V, L, D, headings, shift_k, NK
    a.emb = _p(emb)
    a.feat = _p(feat)
    a.feat_ld_row, a.feat_ld_b, a.feat_ld_t = _ld3(feat)
    a.ctx = _p(ctx)
    a.ctx_ld_row, a.ctx_ld_b, a.ctx_ld_t = _ld3(ctx)
    a.ctx_mask, a.ctx_mask_ld = _p(ctx_mask), (ctx_mask.stride(0) if ctx_mask is not None else 0)
    a.h0, a.c0 = _p(h0), _p(c0)
    a.m_hprev, a.m_h1, a.drop_scale = _p(m_hprev), _p(m_h1), float(scale)
    hw = [half_weight(w) for w in (w_feat, w_lstm, w_att_in, w_att_out)]      # fp16 weight stream (cached per parameter epoch)
    a.w_feat, a.b_feat, a.w_lstm, a.b_ih, a.b_hh = _p(hw[0]), _p(b_feat), _p(hw[1]), _p(b_ih), _p(b_hh)
    a.w_att_in, a.w_att_out = _p(hw[2]), _p(hw[3])
    for k, v in o.items():
        setattr(a, k, _p(v))
    x16 = torch.empty(lib.load().dasa_decoder_rollout_x16_halves(T, B, H, E, F, D), device=dev, dtype=torch.float16)
    a.zpart, a.barrier, a.x16 = _p(zpart), _p(barrier), _p(x16)
    keep = (emb, feat, ctx, ctx_mask, h0, c0, zpart, barrier, hw, x16)      # referenced until the launch is enqueued
    call("dasa_decoder_rollout_fwd", ctypes.byref(a), _stream())
    del keep
    return o


def decoder_rollout_bwd(saved, feat, ctx, ctx_mask, m_hprev, m_h1, scale, w_feat_t, w_lstm_t, w_att_in_t, w_att_out_t, headings,
                        shift_k, d_htilde, d_h1, d_c_last):
    """Backward of decoder_rollout_fwd. `saved` = its output dict; the *_t weights are [in, out] rows (ops.transposed_weight)."""
    T, B, H = saved["htilde"].shape
    E = saved["xh"].shape[2] - feat.shape[3] - H
    V, F = feat.shape[2], feat.shape[3]
    L, D = ctx.shape[2], ctx.shape[3]
    NK = saved["tk"].shape[2]
    dev = feat.device

    def new(*shape):
        return torch.empty(*shape, device=dev, dtype=torch.float32)
    g = {"du": new(T, B, H), "dt2": new(T, B, D), "dgates": new(T, B, 4 * H), "dtk": new(T, B, NK), "demb": new(T, B, E),
         "dfeat": new(T, B, V, F), "dctx": new(T, B, L, D), "dh0": new(B, H), "dc0": new(B, H)}
    scratch = {"dcat": new(B, D + H), "dattn": new(B, F), "dhdir": new(B, H), "dc_carry": new(B, H)}
    zpart = new(lib.load().dasa_decoder_rollout_scratch_floats(B))
    barrier = torch.empty(32, device=dev, dtype=torch.int32)
    d_htilde = d_htilde.contiguous()
    d_h1 = None if d_h1 is None else d_h1.contiguous()
    d_c_last = None if d_c_last is None else d_c_last.contiguous()
    if ctx_mask is not None:
        ctx_mask = ctx_mask if ctx_mask.dtype == torch.uint8 else ctx_mask.to(torch.uint8)
    a = lib.DecoderBwd()
    a.T, a.B, a.H, a.E, a.F, a.V, a.L, a.D, a.headings, a.shift_k, a.NK = T, B, H, E, F, V, L, D, headings, shift_k, NK
    a.feat = _p(feat)
    a.feat_ld_row, a.feat_ld_b, a.feat_ld_t = _ld3(feat)
    a.ctx = _p(ctx)
    a.ctx_ld_row, a.ctx_ld_b, a.ctx_ld_t = _ld3(ctx)
    a.ctx_mask, a.ctx_mask_ld = _p(ctx_mask), (ctx_mask.stride(0) if ctx_mask is not None else 0)
    a.m_hprev, a.m_h1, a.drop_scale = _p(m_hprev), _p(m_h1), float(scale)
    hw = [half_weight(w) for w in (w_feat_t, w_lstm_t, w_att_in_t, w_att_out_t)]
    a.w_feat_t, a.ld_w_feat_t = _p(hw[0]), hw[0].stride(0)
    a.w_lstm_t, a.ld_w_lstm_t = _p(hw[1]), hw[1].stride(0)
    a.w_att_in_t, a.ld_w_att_in_t = _p(hw[2]), hw[2].stride(0)
    a.w_att_out_t, a.ld_w_att_out_t = _p(hw[3]), hw[3].stride(0)
    for k in ("tk", "p", "q", "kappa", "acts", "c", "cat", "t2", "alpha", "htilde"):
        setattr(a, k, _p(saved[k]))
    a.d_htilde, a.d_h1, a.d_c_last = _p(d_htilde), _p(d_h1), _p(d_c_last)
    for k in ("du", "dt2", "dgates", "dtk", "demb", "dctx", "dh0", "dc0"):
        setattr(a, k, _p(g[k]))
    a.dfeat = _p(g["dfeat"])
    a.dfeat_ld_row, a.dfeat_ld_b, a.dfeat_ld_t = _ld3(g["dfeat"])
    for k, v in scratch.items():
        setattr(a, k, _p(v))
    g16 = torch.empty(lib.load().dasa_decoder_rollout_g16_halves(T, B, H, D, NK), device=dev, dtype=torch.float16)
    a.zpart, a.barrier, a.g16 = _p(zpart), _p(barrier), _p(g16)
    keep = (d_htilde, d_h1, d_c_last, ctx_mask, scratch, zpart, barrier, hw, g16)
    call("dasa_decoder_rollout_bwd", ctypes.byref(a), _stream())
    del keep
    return g
