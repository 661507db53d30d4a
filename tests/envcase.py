"""Shared scenario builder for the environment tests: a graph, feature banks, episodes and (optionally) a random valid action
stream for the closed-loop (sampled feedback) mode."""
import numpy as np

from dasa_b200.navgraph import NavGraph


def scenario(n=24, B=5, T=7, C=32, seed=0, graph=None):
    g = graph or NavGraph.synthetic(n, seed)
    rng = np.random.RandomState(77 + seed)
    rgb = np.maximum(rng.randn(g.n, 36, C), 0).astype(np.float32) * 0.5
    dep = np.maximum(rng.randn(g.n, 36, C), 0).astype(np.float32) * 0.5
    start, view, goal = g.sample_episodes(B, seed, min_hops=2, max_hops=5)
    return g, rgb, dep, start, view, goal


def random_actions(g, start, T, seed=0, stop_prob=0.15, ignore_id=-100):
    """A valid closed-loop action stream: at each step a random candidate or END (index deg). Episodes keep acting after END,
    as sampled feedback does in the reference (agent_dg.py:876-882)."""
    rng = np.random.RandomState(500 + seed)
    vp = np.array(start).copy()
    acts = []
    for t in range(T):
        a = np.zeros(len(vp), np.int64)
        for b in range(len(vp)):
            d = int(g.deg[vp[b]])
            a[b] = d if rng.rand() < stop_prob else rng.randint(d)
            if a[b] != d:
                vp[b] = g.nbr[vp[b], a[b]]
        acts.append(a)
    return acts


def lists(g):
    return dict(names=g.names, nbrs=g.nbrs, weights=g.weights, headings=g.headings, elevations=g.elevations, points=g.points)
