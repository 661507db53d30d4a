"""CPU tier, build container only: the environment restatement (oracle/env_restated.py) and the host-side graph tables
(dasa_b200/navgraph.py) against the UNMODIFIED reference env.py / agent_dg.py / utils.py functions driven through
oracle/ref_env_driver.py (skipped where the reference tree is not mounted, e.g. the GPU box)."""
import json
import os

import numpy as np
import pytest

from dasa_b200.navgraph import NavGraph
from oracle import env_restated as E
from oracle import load_reference
from tests.envcase import lists, random_actions, scenario

pytestmark = pytest.mark.skipif(not load_reference.available(), reason="reference tree not mounted")
KEYS = ("input_a_t", "f_t", "d_t", "cand_feat", "cand_dfeat", "cand_leng", "target", "action", "dist", "reward", "mask", "ended",
        "viewIndex")


def _both(g, rgb, dep, start, view, goal, T, actions):
    from oracle.ref_env_driver import ReferenceEnv
    C = rgb.shape[-1]
    ref = ReferenceEnv("scanA", features=rgb, dfeatures=dep, rgb_size=C, **lists(g))
    ora = E.RefStyleEnv("scanA", features=rgb, dfeatures=dep, **lists(g))
    sn, gn = [g.names[i] for i in start], [g.names[i] for i in goal]
    ref.new_episodes(sn, view, gn)
    ora.new_episodes(sn, view, gn)
    return ref.rollout(T, actions), E.rollout(ora, T, C, 128, actions), ref, ora


@pytest.mark.parametrize("seed", [0, 1])
@pytest.mark.parametrize("closed_loop", [False, True])
def test_env_restatement_matches_reference(seed, closed_loop):
    g, rgb, dep, start, view, goal = scenario(seed=seed)
    T = 7
    actions = random_actions(g, start, T, seed) if closed_loop else None
    try:
        want, got, _, _ = _both(g, rgb, dep, start, view, goal, T, actions)
    except NameError:
        pytest.skip("random action stream hit the reference's zero-progress NameError")
    for t in range(T):
        for k in KEYS:
            assert np.array_equal(np.asarray(want[t][k]), np.asarray(got[t][k])), "step %d %s" % (t, k)
        assert want[t]["viewpoint"] == got[t]["viewpoint"]


def test_paths_match_networkx_and_tables():
    """all_pairs_paths (oracle) == networkx as env.py:195-198 calls it; NavGraph.dist / next_hop agree with both."""
    g, rgb, dep, start, view, goal = scenario(n=40, seed=3)
    _, _, ref, ora = _both(g, rgb, dep, start, view, goal, 1, None)
    rp, rd = ref.rb.paths["scanA"], ref.rb.distances["scanA"]
    op, od = ora.paths["scanA"], ora.distances["scanA"]
    idx = {n: i for i, n in enumerate(g.names)}
    for s in g.names:
        assert set(rp[s]) == set(op[s])
        for d in rp[s]:
            assert rp[s][d] == op[s][d]
            assert rd[s][d] == od[s][d]
            i, j = idx[s], idx[d]
            assert g.dist64[i, j] == rd[s][d]
            hop = g.next_hop[i, j]
            assert (hop == -1) if s == d else (g.names[g.nbr[i, hop]] == rp[s][d][1])


def test_real_connectivity_graph():
    """The same on a real Matterport connectivity graph shipped with the reference, loaded by the reference's own
    utils.load_nav_graphs (the builder keeps its edge order so Dijkstra ties resolve identically)."""
    import networkx as nx
    scan = "17DRP5sb8fy"
    path = "/root/reference/connectivity/%s_connectivity.json" % scan
    items = json.load(open(path))
    g = NavGraph.from_connectivity(items)
    ref = load_reference.load()
    cwd = os.getcwd()
    os.chdir("/root/reference")
    try:
        G = ref.utils.load_nav_graphs([scan])[scan]
    finally:
        os.chdir(cwd)
    paths = dict(nx.all_pairs_dijkstra_path(G))
    dists = dict(nx.all_pairs_dijkstra_path_length(G))
    idx = {n: i for i, n in enumerate(g.names)}
    assert set(idx) == set(G.nodes)
    for s in g.names:
        for d in g.names:
            i, j = idx[s], idx[d]
            assert g.dist64[i, j] == dists[s][d]
            hop = g.next_hop[i, j]
            assert (hop == -1) if s == d else (g.names[g.nbr[i, hop]] == paths[s][d][1])
    # and a teacher rollout through the reference env on that real graph
    rng = np.random.RandomState(0)
    rgb = rng.rand(g.n, 36, 16).astype(np.float32)
    dep = rng.rand(g.n, 36, 16).astype(np.float32)
    start, view, goal = g.sample_episodes(4, 0)
    want, got, _, _ = _both(g, rgb, dep, start, view, goal, 8, None)
    for t in range(8):
        for k in KEYS:
            assert np.array_equal(np.asarray(want[t][k]), np.asarray(got[t][k])), "step %d %s" % (t, k)
