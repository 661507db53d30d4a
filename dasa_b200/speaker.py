"""Speaker (instruction generator) inference path of the augmented rollouts: model.SpeakerEncoder /
SpeakerDecoder (model.py:984-1078) with the reference's constructor arguments, forward signatures and state_dict keys, and the
greedy branch of Speaker.infer_batch (speaker.py:265-350). Everything numeric runs on the kernels of the navigation path:
the bi-LSTM sequence kernels, the single-pass soft-dot attention, the stacked-weight LSTM cell, the GEMMs, plus one word-select
kernel (argmax with <UNK> masked, <PAD> after <EOS>, ended flags) so that the 120-step decode never syncs with the host
(the reference does one .cpu() per word). Inference only (the agent calls infer_batch in eval mode under no_grad,
agent_dg.py:656-675); training the speaker is out of scope. No CPU fallback."""
import torch
import torch.nn as nn

from . import functions as Fn
from . import ops
from .modules import SoftDotAttention


def _bilstm(lstm, x):
    """nn.LSTM(batch_first, bidirectional) over full-length (un-packed) sequences: [B, L, In] -> [B, L, 2H]."""
    B, L, _ = x.shape
    lengths = torch.full((B,), L, dtype=torch.int32, device=x.device)
    out, _, _ = Fn.BiLSTMFn.apply(x, lengths, lstm.weight_ih_l0, lstm.weight_hh_l0, lstm.bias_ih_l0, lstm.bias_hh_l0,
                                  lstm.weight_ih_l0_reverse, lstm.weight_hh_l0_reverse, lstm.bias_ih_l0_reverse,
                                  lstm.bias_hh_l0_reverse)
    return out


class SpeakerEncoder(nn.Module):
    """model.py:984-1036 (bidirectional=True, the README configuration)."""

    def __init__(self, feature_size, hidden_size, dropout_ratio, bidirectional=True, angle_feat_size=128, featdropout=0.4):
        super().__init__()
        if not bidirectional:
            raise NotImplementedError("the speaker is built with --bidir True (param.py:101)")
        self.num_directions, self.hidden_size, self.num_layers, self.feature_size = 2, hidden_size, 1, feature_size
        self.lstm = nn.LSTM(feature_size, hidden_size // 2, 1, batch_first=True, bidirectional=True)
        self.drop = nn.Dropout(p=dropout_ratio)
        self.drop3 = nn.Dropout(p=featdropout)
        self.attention_layer = SoftDotAttention(hidden_size, feature_size)
        self.post_lstm = nn.LSTM(hidden_size, hidden_size // 2, 1, batch_first=True, bidirectional=True)

    def forward(self, action_embeds, feature, lengths=None, already_dropfeat=False):
        """action_embeds [B, L, F], feature [B, L, 36, F] -> ctx [B, L, hidden]. Eval mode only (dropouts are identities)."""
        if self.training:
            raise NotImplementedError("speaker training is out of scope: infer_batch runs the speaker in eval mode")
        ctx = _bilstm(self.lstm, action_embeds)
        B, L, _ = ctx.shape
        x, _ = self.attention_layer(ctx.reshape(B * L, self.hidden_size), feature.reshape(B * L, -1, self.feature_size))
        return _bilstm(self.post_lstm, x.view(B, L, -1))


class SpeakerDecoder(nn.Module):
    """model.py:1038-1078."""

    def __init__(self, vocab_size, embedding_size, padding_idx, hidden_size, dropout_ratio):
        super().__init__()
        self.hidden_size = hidden_size
        self.embedding = nn.Embedding(vocab_size, embedding_size, padding_idx)
        self.lstm = nn.LSTM(embedding_size, hidden_size, batch_first=True)
        self.drop = nn.Dropout(dropout_ratio)
        self.attention_layer = SoftDotAttention(hidden_size, hidden_size)
        self.projection = nn.Linear(hidden_size, vocab_size)
        self.baseline_projection = nn.Sequential(nn.Linear(hidden_size, 128), nn.ReLU(), nn.Dropout(dropout_ratio),
                                                 nn.Linear(128, 1))

    def forward(self, words, ctx, ctx_mask, h0, c0):
        """One decode step (words [B, 1], as infer_batch calls it): -> (logit [B, 1, V], h1 [1, B, H], c1 [1, B, H])."""
        if self.training:
            raise NotImplementedError("speaker training is out of scope: infer_batch runs the speaker in eval mode")
        if words.dim() != 2 or words.shape[1] != 1:
            raise NotImplementedError("teacher-forced multi-word decoding belongs to speaker training (out of scope)")
        l = self.lstm
        embeds = self.embedding.weight.index_select(0, words.reshape(-1))
        xh = torch.cat((embeds, h0[0]), 1)
        h1, c1 = Fn.LSTMCellFn.apply(xh, c0[0], l.weight_ih_l0, l.weight_hh_l0, l.bias_ih_l0, l.bias_hh_l0)
        x, _ = self.attention_layer(h1, ctx, ctx_mask)
        logit = Fn.linear(x, self.projection.weight, self.projection.bias)
        return logit.unsqueeze(1), h1.unsqueeze(0), c1.unsqueeze(0)


@torch.no_grad()
def infer_batch(encoder, decoder, can_feats, img_feats, lengths, bos, eos, pad, unk, max_decode=120, check_every=8,
                featdropmask=None, featdrop_scale=1.0):
    """Greedy branch of Speaker.infer_batch (speaker.py:265-350) given the shortest-path features from_shortest_path()
    produces: can_feats [B, L, F] (the taken candidates), img_feats [B, L, 36, F], lengths (list / tensor of path lengths).
    Returns the instructions as an int64 tensor [B, n] on the device, n = the step at which every sequence had emitted <EOS>
    (or max_decode), exactly the array the reference stacks. The ended flags stay on the device; the host looks at them every
    `check_every` steps only."""
    encoder.eval()
    decoder.eval()
    dev = can_feats.device
    B = can_feats.shape[0]
    if featdropmask is not None:
        # the agent's env-drop noise shared with the speaker (speaker.py:291-293; agent_dg.py:656-659): a uint8 keep vector [C]
        # with its scale; the RGB part of both feature tensors is multiplied, the angle part is left alone
        C = featdropmask.numel()
        keep = ops.as_keep_mask(featdropmask)

        def masked(x):
            rows = x.numel() // x.shape[-1]
            out = torch.empty_like(x)
            ops.dropout_apply(x[..., :C], keep.view(1, C).expand(rows, C).contiguous(), featdrop_scale, out=out[..., :C])
            ops.axpy2d(1.0, x[..., C:], out[..., C:], accumulate=False)
            return out
        can_feats, img_feats = masked(can_feats), masked(img_feats)
    ctx = encoder(can_feats, img_feats, lengths, already_dropfeat=featdropmask is not None)
    lengths = torch.as_tensor(lengths, device=dev)
    ctx_mask = (torch.arange(ctx.shape[1], device=dev)[None, :] >= lengths[:, None]).to(torch.uint8)     # utils.length2mask
    h_t = torch.zeros(1, B, decoder.hidden_size, device=dev)
    c_t = torch.zeros(1, B, decoder.hidden_size, device=dev)
    ended = torch.zeros(B, dtype=torch.uint8, device=dev)
    word = torch.full((B, 1), int(bos), dtype=torch.int64, device=dev)
    words = torch.full((B, max_decode), int(pad), dtype=torch.int64, device=dev)
    n = max_decode
    for i in range(max_decode):
        logits, h_t, c_t = decoder(word, ctx, ctx_mask, h_t, c_t)
        lg = logits[:, 0]
        nxt = torch.empty(B, dtype=torch.int64, device=dev)
        ops.call("dasa_speaker_select", ops._p(lg), lg.stride(0), B, lg.shape[1], int(unk), int(pad), int(eos), ops._p(ended),
                 ops._p(nxt), ops._p(words[:, i]), words.stride(0), ops._stream())
        word = nxt.view(B, 1)
        if (i + 1) % check_every == 0 and bool(ended.all()):
            break
    # the reference stops right after the step at which the last sequence ended: first column from which all are <PAD>
    done_at = (words == int(eos)).int().argmax(1)                       # position of each sequence's <EOS> (0 if none)
    has_eos = (words == int(eos)).any(1)
    if bool(has_eos.all()):
        n = int(done_at.max()) + 1
    return words[:, :n]
