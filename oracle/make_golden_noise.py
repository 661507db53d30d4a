"""TEST INFRASTRUCTURE ONLY (build container). tests/golden/small_noise.pt: a teacher-forced rollout with the augmented-rollout
feature dropout (consistent_drop, after_adain, depth_drop; agent_dg.py:780-785) through the UNMODIFIED reference modules
(oracle/make_golden.reference_rollout), so the oracle's `noise` path stays pinned where the reference tree is not mounted.

    python -m oracle.make_golden_noise
"""
import contextlib
import io
import os

import torch

from dasa_b200 import synth
from dasa_b200.config import SMALL
from oracle import load_reference
from oracle.make_golden import build_reference_modules, reference_rollout

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "small_noise.pt")
META = dict(state_seed=2, episodes=dict(B=4, T=3, seed=33), noise_seed=2)


def noise_vector():
    gen = torch.Generator().manual_seed(META["noise_seed"])
    return (torch.rand(SMALL.rgb_size, generator=gen) >= SMALL.featdropout).float() / (1 - SMALL.featdropout)


def main():
    with contextlib.redirect_stdout(io.StringIO()):
        ref = load_reference.load()
    st = synth.policy_state(SMALL, META["state_seed"])
    mods = build_reference_modules(ref, SMALL, st)
    for m in mods:
        m.eval()
    ep = synth.Episodes(cfg=SMALL, **META["episodes"])
    with torch.no_grad():
        loss, logits, _ = reference_rollout(ref, mods, SMALL, ep, META["episodes"]["T"], noise=noise_vector())
    torch.save({"meta": META, "loss": loss, "logits": torch.stack(logits)}, OUT)
    print("wrote", OUT, os.path.getsize(OUT), "bytes, loss", float(loss))


if __name__ == "__main__":
    main()
