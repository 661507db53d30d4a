"""TEST INFRASTRUCTURE ONLY (build container). tests/golden/speaker_small.pt from the UNMODIFIED reference model.SpeakerEncoder /
SpeakerDecoder (model.py:984-1078) driven by the greedy loop of Speaker.infer_batch (speaker.py:302-343).

    python -m oracle.make_golden_speaker
"""
import contextlib
import io
import os

import torch

from dasa_b200 import synth
from dasa_b200.config import SMALL
from oracle import load_reference

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "speaker_small.pt")
DIMS = dict(rnn_dim=128, wemb=64, vocab=200)
TOK = dict(pad=0, unk=1, eos=2, bos=199)


CASES = ((1, 0.65), (2, 0.65))      # (seed, <EOS> bias): one sequence never ends within the budget / all end by step 5 (early exit)


def speaker_case(seed=0, eos_bias=0.65, B=4, L=5):
    enc_sd, dec_sd = synth.speaker_state(SMALL.feat, seed=seed, **DIMS)
    dec_sd["projection.bias"][TOK["eos"]] += eos_bias   # sequences end at different steps within the decode budget
    can, img, lengths = synth.speaker_inputs(B, L, SMALL, seed)
    return enc_sd, dec_sd, can, img, lengths


def reference_infer(ref, enc_sd, dec_sd, can, img, lengths, max_decode):
    """The reference modules + the loop of speaker.py:296-343 (sampling=False)."""
    ref.args.featdropout, ref.args.angle_feat_size = SMALL.featdropout, SMALL.angle_size
    with contextlib.redirect_stdout(io.StringIO()):
        enc = ref.model.SpeakerEncoder(SMALL.feat, DIMS["rnn_dim"], 0.5, bidirectional=True)
    dec = ref.model.SpeakerDecoder(DIMS["vocab"], DIMS["wemb"], TOK["pad"], DIMS["rnn_dim"], 0.5)
    enc.load_state_dict(enc_sd, strict=True)
    dec.load_state_dict(dec_sd, strict=True)
    enc.eval(); dec.eval()
    with torch.no_grad():
        ctx = enc(can.clone(), img.clone(), lengths, already_dropfeat=False)
        B = ctx.shape[0]
        ctx_mask = torch.arange(ctx.shape[1])[None, :] >= lengths[:, None]               # utils.length2mask (utils.py:503-508)
        h_t = torch.zeros(1, B, DIMS["rnn_dim"]); c_t = torch.zeros(1, B, DIMS["rnn_dim"])
        ended = torch.zeros(B, dtype=torch.bool)
        word = torch.full((B, 1), TOK["bos"], dtype=torch.int64)
        words, logit_steps = [], []
        for i in range(max_decode):
            logits, h_t, c_t = dec(word, ctx, ctx_mask, h_t, c_t)
            logits = logits.squeeze()
            logits[:, TOK["unk"]] = -float("inf")
            logit_steps.append(logits.clone())
            values, word = logits.max(1)
            cpu_word = word.clone()
            cpu_word[ended] = TOK["pad"]
            words.append(cpu_word)
            word = word.view(-1, 1)
            ended = ended | (cpu_word == TOK["eos"])
            if bool(ended.all()):
                break
    return torch.stack(words, 1), ctx, logit_steps


def main():
    with contextlib.redirect_stdout(io.StringIO()):
        ref = load_reference.load()
    out = {"max_decode": 24, "cases": []}
    for seed, bias in CASES:
        enc_sd, dec_sd, can, img, lengths = speaker_case(seed, bias)
        words, ctx, logit_steps = reference_infer(ref, enc_sd, dec_sd, can, img, lengths, 24)
        out["cases"].append({"seed": seed, "eos_bias": bias, "words": words, "ctx": ctx, "logits0": logit_steps[0],
                             "logits_last": logit_steps[-1]})
        print("case", seed, bias, "words:\n", words)
    torch.save(out, OUT)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
