"""TEST INFRASTRUCTURE ONLY (build container). Generates tests/golden/env_rollout.pt from the UNMODIFIED reference env.py /
agent_dg.py helpers (oracle/ref_env_driver.py): a small graph, feature banks, episodes and the per-step tensors of a
teacher-forced and a closed-loop (injected actions) rollout. The GPU box has no reference tree; this file is what travels.

    python -m oracle.make_golden_env
"""
import os

import numpy as np
import torch

from tests.envcase import lists, random_actions, scenario

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "env_rollout.pt")
KEYS = ("input_a_t", "f_t", "d_t", "cand_feat", "cand_dfeat", "cand_leng", "target", "action", "dist", "reward", "mask", "ended",
        "viewIndex")


def main():
    from oracle.ref_env_driver import ReferenceEnv
    n, B, T, C = 16, 3, 6, 32
    for seed in range(20):                                     # first seed whose random walk never makes a zero-progress move
        g, rgb, dep, start, view, goal = scenario(n=n, B=B, T=T, C=C, seed=seed)
        actions = random_actions(g, start, T, seed)
        ref = ReferenceEnv("scanG", features=rgb, dfeatures=dep, rgb_size=C, **lists(g))
        sn, gn = [g.names[i] for i in start], [g.names[i] for i in goal]
        try:
            ref.new_episodes(sn, view, gn)
            teacher = ref.rollout(T, None)
            ref.new_episodes(sn, view, gn)
            closed = ref.rollout(T, actions)
        except NameError:
            continue
        break
    idx = {nm: i for i, nm in enumerate(g.names)}

    def pack(steps):
        out = []
        for t, s in enumerate(steps):
            # the [B, 36, C+A] panoramas are kept for the first two steps only (fixture size); candidates + scalars for all
            d = {k: torch.from_numpy(np.ascontiguousarray(np.asarray(s[k]))) for k in KEYS if t < 2 or k not in ("f_t", "d_t")}
            d["vp"] = torch.tensor([idx[v] for v in s["viewpoint"]], dtype=torch.int32)
            out.append(d)
        return out
    torch.save({"seed": seed, "graph": {k: v for k, v in lists(g).items()}, "rgb": torch.from_numpy(rgb), "dep": torch.from_numpy(dep),
                "start": torch.from_numpy(start), "view": torch.from_numpy(view), "goal": torch.from_numpy(goal),
                "actions": [torch.from_numpy(a) for a in actions], "teacher": pack(teacher), "closed": pack(closed),
                "C": C, "T": T}, OUT)
    print("wrote", OUT, os.path.getsize(OUT), "bytes (seed %d)" % seed)


if __name__ == "__main__":
    main()
