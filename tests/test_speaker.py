"""Speaker inference path: oracle vs the reference-generated fixture (CPU tier), live reference
(build container) and the CUDA path vs the oracle (GPU tier)."""
import os

import pytest
import torch

from dasa_b200 import synth
from dasa_b200.config import SMALL
from oracle import load_reference
from oracle import restated as R
from oracle.make_golden_speaker import CASES, DIMS, TOK, speaker_case

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "speaker_small.pt")


def _oracle(seed, bias, max_decode):
    enc_sd, dec_sd, can, img, lengths = speaker_case(seed, bias)
    return R.speaker_infer_greedy(enc_sd, dec_sd, can, img, lengths, TOK["bos"], TOK["eos"], TOK["pad"], TOK["unk"], max_decode)


def test_oracle_speaker_matches_reference_golden():
    gold = torch.load(GOLD, weights_only=False)
    for case in gold["cases"]:
        words, ctx, logits = _oracle(case["seed"], case["eos_bias"], gold["max_decode"])
        assert torch.equal(words, case["words"])
        torch.testing.assert_close(ctx, case["ctx"], rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(logits[0], case["logits0"], rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(logits[-1], case["logits_last"], rtol=1e-4, atol=1e-5)


@pytest.mark.skipif(not load_reference.available(), reason="reference tree not mounted")
def test_oracle_speaker_matches_live_reference():
    import contextlib
    import io
    from oracle.make_golden_speaker import reference_infer
    with contextlib.redirect_stdout(io.StringIO()):
        ref = load_reference.load()
    enc_sd, dec_sd, can, img, lengths = speaker_case(5, 0.65, B=3, L=4)
    want_words, want_ctx, want_logits = reference_infer(ref, enc_sd, dec_sd, can, img, lengths, 16)
    words, ctx, logits = R.speaker_infer_greedy(enc_sd, dec_sd, can, img, lengths, TOK["bos"], TOK["eos"], TOK["pad"], TOK["unk"], 16)
    assert torch.equal(words, want_words)
    torch.testing.assert_close(ctx, want_ctx, rtol=1e-5, atol=1e-6)
    # state_dict contract of the drop-in modules
    from dasa_b200 import speaker as S
    enc = S.SpeakerEncoder(SMALL.feat, DIMS["rnn_dim"], 0.5, True)
    dec = S.SpeakerDecoder(DIMS["vocab"], DIMS["wemb"], TOK["pad"], DIMS["rnn_dim"], 0.5)
    assert [(k, tuple(v.shape)) for k, v in enc.state_dict().items()] == [(k, tuple(v.shape)) for k, v in enc_sd.items()]
    assert [(k, tuple(v.shape)) for k, v in dec.state_dict().items()] == [(k, tuple(v.shape)) for k, v in dec_sd.items()]


@pytest.mark.gpu
@pytest.mark.parametrize("seed,bias", list(CASES) + [(5, 0.65)])
def test_speaker_infer_matches_oracle(seed, bias):
    from dasa_b200 import speaker as S
    max_decode = 24
    enc_sd, dec_sd, can, img, lengths = speaker_case(seed, bias)
    want_words, want_ctx, want_logits = R.speaker_infer_greedy(enc_sd, dec_sd, can, img, lengths, TOK["bos"], TOK["eos"], TOK["pad"],
                                                               TOK["unk"], max_decode)
    enc = S.SpeakerEncoder(SMALL.feat, DIMS["rnn_dim"], 0.5, True).cuda().eval()
    dec = S.SpeakerDecoder(DIMS["vocab"], DIMS["wemb"], TOK["pad"], DIMS["rnn_dim"], 0.5).cuda().eval()
    enc.load_state_dict(enc_sd, strict=True)
    dec.load_state_dict(dec_sd, strict=True)
    with torch.no_grad():
        ctx = enc(can.cuda(), img.cuda(), lengths)
    err = float((ctx.cpu() - want_ctx).abs().max() / want_ctx.abs().max())
    assert err <= 1e-4, "speaker encoder ctx rel err %.2e" % err
    words = S.infer_batch(enc, dec, can.cuda(), img.cuda(), lengths, TOK["bos"], TOK["eos"], TOK["pad"], TOK["unk"], max_decode,
                          check_every=4)
    # greedy words are bit-exact wherever the oracle's top-2 logit margin exceeds the fp32 tolerance (north_star's rule)
    assert tuple(words.shape) == tuple(want_words.shape)
    margins = torch.stack([lg.topk(2, 1).values[:, 0] - lg.topk(2, 1).values[:, 1] for lg in want_logits], 1)
    safe = (margins > 1e-4).cumprod(1).bool()             # once a near-tie flips a word the continuation may differ
    assert torch.equal(words.cpu()[safe], want_words[safe])
    assert safe.float().mean() > 0.9


@pytest.mark.gpu
def test_speaker_full_geometry():
    """README geometry (features 2176, rnn_dim 512, wemb 256, 992-word vocabulary, 20 paths of up to 8 viewpoints): encoder ctx
    and the first decode steps against the oracle; the decoded array obeys the protocol invariants (nothing but <PAD> after
    <EOS>, no <UNK>, no <PAD> before <EOS>)."""
    from dasa_b200 import speaker as S
    from dasa_b200.config import FULL
    tok = dict(pad=0, unk=1, eos=2, bos=991)
    enc_sd, dec_sd = synth.speaker_state(FULL.feat, rnn_dim=512, wemb=256, vocab=992, seed=3)
    dec_sd["projection.bias"][tok["eos"]] += 0.6
    can, img, lengths = synth.speaker_inputs(20, 8, FULL, 3)
    want_words, want_ctx, want_logits = R.speaker_infer_greedy(enc_sd, dec_sd, can, img, lengths, tok["bos"], tok["eos"], tok["pad"],
                                                               tok["unk"], 6)
    enc = S.SpeakerEncoder(FULL.feat, 512, 0.5, True).cuda().eval()
    dec = S.SpeakerDecoder(992, 256, tok["pad"], 512, 0.5).cuda().eval()
    enc.load_state_dict(enc_sd, strict=True)
    dec.load_state_dict(dec_sd, strict=True)
    with torch.no_grad():
        ctx = enc(can.cuda(), img.cuda(), lengths)
    err = float((ctx.cpu() - want_ctx).abs().max() / want_ctx.abs().max())
    assert err <= 1e-4, "speaker encoder ctx rel err %.2e" % err
    words = S.infer_batch(enc, dec, can.cuda(), img.cuda(), lengths, tok["bos"], tok["eos"], tok["pad"], tok["unk"], 40).cpu()
    n = min(words.shape[1], want_words.shape[1])
    margins = torch.stack([lg.topk(2, 1).values[:, 0] - lg.topk(2, 1).values[:, 1] for lg in want_logits], 1)[:, :n]
    safe = (margins > 1e-4).cumprod(1).bool()
    assert torch.equal(words[:, :n][safe], want_words[:, :n][safe])
    assert not bool((words == tok["unk"]).any())
    for row in words.tolist():
        if tok["eos"] in row:
            i = row.index(tok["eos"])
            assert all(w == tok["pad"] for w in row[i + 1:]) and tok["pad"] not in row[:i]
        else:
            assert tok["pad"] not in row


@pytest.mark.gpu
def test_speaker_featdropmask():
    """infer_batch(featdropmask=noise) (speaker.py:291-296): the agent's env-drop mask multiplies the RGB part of both feature
    tensors before the encoder."""
    from dasa_b200 import speaker as S
    enc_sd, dec_sd, can, img, lengths = speaker_case(1, 0.65)
    C = SMALL.rgb_size
    gen = torch.Generator().manual_seed(4)
    keep = torch.rand(C, generator=gen) >= 0.4
    noise = keep.float() / 0.6
    can2, img2 = can.clone(), img.clone()
    can2[..., :C] *= noise
    img2[..., :C] *= noise
    want_words, want_ctx, want_logits = R.speaker_infer_greedy(enc_sd, dec_sd, can2, img2, lengths, TOK["bos"], TOK["eos"], TOK["pad"],
                                                               TOK["unk"], 16)
    enc = S.SpeakerEncoder(SMALL.feat, DIMS["rnn_dim"], 0.5, True).cuda().eval()
    dec = S.SpeakerDecoder(DIMS["vocab"], DIMS["wemb"], TOK["pad"], DIMS["rnn_dim"], 0.5).cuda().eval()
    enc.load_state_dict(enc_sd, strict=True)
    dec.load_state_dict(dec_sd, strict=True)
    words = S.infer_batch(enc, dec, can.cuda(), img.cuda(), lengths, TOK["bos"], TOK["eos"], TOK["pad"], TOK["unk"], 16,
                          featdropmask=keep.to(torch.uint8).cuda(), featdrop_scale=1 / 0.6).cpu()
    assert tuple(words.shape) == tuple(want_words.shape)
    margins = torch.stack([lg.topk(2, 1).values[:, 0] - lg.topk(2, 1).values[:, 1] for lg in want_logits], 1)
    safe = (margins > 1e-4).cumprod(1).bool()
    assert torch.equal(words[safe], want_words[safe]) and safe.float().mean() > 0.9
