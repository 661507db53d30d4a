"""TEST INFRASTRUCTURE ONLY (build container; needs /root/reference). Fixtures that pin the BENCHMARKED configuration —
full geometry, B = 20 episodes (BASELINE.json configs[1]) — to the UNMODIFIED reference modules:

  tests/golden/bench_eval.pt    eval mode, T = 4 teacher-forced actions: loss, every logit, the greedy actions, strided
                                digests of h_t (agent_dg.py:725-851 around the reference's DGAdaChannel / DicEncoder /
                                BAttnDecoderLSTM).
  tests/golden/bench_train.pt   train mode, T = 2, every nn.Dropout mask drawn from a seeded CPU generator in call order:
                                loss, logits and a digest (norm + strided sample) of the gradient of every trainable tensor.
                                The masks themselves are NOT stored (60 M flags): the fixture keeps the generator seed and the
                                (tag, shape, p) of every dropout call in call order — the tags come from running the oracle with
                                the same masks, which also checks oracle == reference at this size — and the tests regenerate
                                the identical masks with torch.rand on the same seeded generator.

    python -m oracle.make_golden_bench
"""
import contextlib
import io
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from dasa_b200 import synth                                                     # noqa: E402
from dasa_b200.config import FULL                                               # noqa: E402
from oracle import load_reference                                               # noqa: E402
from oracle import restated as R                                                # noqa: E402
from oracle.make_golden import build_reference_modules, recorded_dropout, reference_rollout, sample   # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
EVAL_META = dict(cfg="FULL", seed=0, episodes=dict(B=20, T=4, seed=100))
TRAIN_META = dict(cfg="FULL", seed=0, episodes=dict(B=20, T=2, seed=101), mask_seed=4321)
GRAD_SAMPLE = 4099


def regenerate_masks(calls, mask_seed):
    """The keep masks of a recorded train-mode run: same generator, same call order, same shapes as
    oracle.make_golden.recorded_dropout drew them. calls = [(tag, shape, p), ...]. -> {tag: bool keep tensor}."""
    g = torch.Generator().manual_seed(mask_seed)
    out = {}
    for tag, shape, p in calls:
        assert tag not in out, "dropout tag drawn twice: %s" % tag
        out[tag] = torch.rand(tuple(shape), generator=g) >= p
    return out


class _TagRecorder:
    """oracle drops that replay masks in call order and note which tag consumed which mask."""
    training = True

    def __init__(self, masks):
        self.masks, self.i, self.calls = masks, 0, []

    def __call__(self, x, p, tag):
        m = self.masks[self.i]
        self.i += 1
        assert m.shape == x.shape, (tag, m.shape, x.shape)
        self.calls.append((tag, tuple(x.shape), float(p)))
        return x * m.to(x.dtype)


def main():
    with contextlib.redirect_stdout(io.StringIO()):
        ref = load_reference.load()
    torch.set_num_threads(8)
    cfg = FULL
    state = synth.policy_state(cfg, 0)
    mods = build_reference_modules(ref, cfg, state)

    # ------------------------------------------------------------------------------------------------ eval, T = 4
    for m in mods:
        m.eval()
    ep = synth.Episodes(cfg=cfg, **EVAL_META["episodes"])
    with torch.no_grad():
        loss, logits, hs = reference_rollout(ref, mods, cfg, ep, EVAL_META["episodes"]["T"])
    lg = torch.stack(logits)
    out = {"meta": EVAL_META, "loss": loss, "logits": lg, "actions": lg.argmax(-1),
           "h_t_sample": torch.stack([sample(h, 1031) for h in hs])}
    torch.save(out, os.path.join(GOLDEN, "bench_eval.pt"))
    print("bench_eval.pt written: loss %.6f" % float(loss))

    # ------------------------------------------------------------------------------- train, T = 2, recorded masks
    for m in mods:
        m.train()
        m.zero_grad()
    ep2 = synth.Episodes(cfg=cfg, **TRAIN_META["episodes"])
    rec = []
    with recorded_dropout(TRAIN_META["mask_seed"], rec):
        loss, logits, _ = reference_rollout(ref, mods, cfg, ep2, TRAIN_META["episodes"]["T"])
    loss.backward()
    enc, dec, cri, ada = mods
    grads = {}
    for name, m in (("encoder", enc), ("decoder", dec), ("adaIn", ada)):
        for k, prm in m.named_parameters():
            if prm.grad is not None:
                grads[name + "." + k] = {"norm": prm.grad.norm(), "sample": sample(prm.grad, GRAD_SAMPLE)}
    # tags in call order: run the oracle with the recorded (pre-scaled) masks; it must reproduce the reference's loss
    ost = {g: {k: v.clone() for k, v in d.items()} for g, d in state.items()}
    tagger = _TagRecorder(rec)
    with torch.no_grad():
        loss_o, logits_o, _ = R.teacher_rollout(ost, cfg, ep2, TRAIN_META["episodes"]["T"], drops=tagger)
    assert tagger.i == len(rec)
    assert abs(float(loss_o) - float(loss)) <= 1e-4 * abs(float(loss)), (float(loss_o), float(loss))
    calls = tagger.calls
    # the regenerated masks must be the recorded ones, bit for bit
    regen = regenerate_masks(calls, TRAIN_META["mask_seed"])
    for (tag, shape, p), m in zip(calls, rec):
        assert torch.equal(regen[tag], m != 0), tag
    tr = {"meta": TRAIN_META, "calls": calls, "loss": loss.detach(), "logits": torch.stack(logits).detach(), "grads": grads}
    torch.save(tr, os.path.join(GOLDEN, "bench_train.pt"))
    print("bench_train.pt written: loss %.6f, %d dropout calls, %d gradients" % (float(loss), len(calls), len(grads)))


if __name__ == "__main__":
    main()
