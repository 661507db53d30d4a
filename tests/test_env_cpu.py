"""CPU tier: the environment oracle against the reference-generated fixture, and the host-side graph tables
(dasa_b200/navgraph.py) against the oracle's dict-of-dict paths. No compute call into the CUDA library."""
import os

import numpy as np
import torch

from dasa_b200.navgraph import NavGraph
from oracle import env_restated as E
from tests.envcase import lists, scenario

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "env_rollout.pt")


def test_oracle_env_matches_reference_golden():
    gold = torch.load(GOLD, weights_only=False)
    G = gold["graph"]
    names = G["names"]
    C, T = gold["C"], gold["T"]
    sn = [names[i] for i in gold["start"].tolist()]
    gn = [names[i] for i in gold["goal"].tolist()]
    for key, actions in (("teacher", None), ("closed", [a.numpy() for a in gold["actions"]])):
        env = E.RefStyleEnv("scanG", features=gold["rgb"].numpy(), dfeatures=gold["dep"].numpy(), **G)
        env.new_episodes(sn, gold["view"].numpy(), gn)
        steps = E.rollout(env, T, C, 128, actions)
        for t in range(T):
            for k, v in gold[key][t].items():
                if k == "vp":
                    assert [names[i] for i in v.tolist()] == steps[t]["viewpoint"]
                else:
                    assert np.array_equal(v.numpy(), np.asarray(steps[t][k])), "%s step %d %s" % (key, t, k)


def test_navgraph_tables_match_oracle_paths():
    g, *_ = scenario(n=48, seed=9)
    L = lists(g)
    paths, dists = E.all_pairs_paths(L["names"], L["nbrs"], L["weights"])
    for i, s in enumerate(g.names):
        assert len(paths[s]) == g.n                     # the synthetic builder returns a connected graph
        for j, d in enumerate(g.names):
            assert g.dist64[i, j] == dists[s][d]
            assert g.dist[i, j] == np.float32(dists[s][d])
            hop = g.next_hop[i, j]
            if i == j:
                assert hop == -1
            else:
                assert g.names[g.nbr[i, hop]] == paths[s][d][1]
    assert g.deg.max() <= 13 and g.deg.min() >= 1
    assert (g.nbr_point >= 0).all() and (g.nbr_point < 36).all()
    hops = g.hops()
    s, v, goal = g.sample_episodes(16, 0)
    assert ((hops[s, goal] >= 3) & (hops[s, goal] <= 7)).all() and ((v >= 12) & (v < 24)).all()


def test_angle_tables_bitexact():
    """cand_angle / view_angle / agent_angle hold exactly utils.angle_feature's float32 values (oracle restatement)."""
    g, *_ = scenario(n=12, seed=2)
    R30 = E.R30
    for hb in (0, 5, 10, 11):
        want = E.point_angle_feature(hb, 128)
        assert np.array_equal(np.tile(g.view_angle[hb], (1, 32)), want)
    for ix in (0, 13, 29, 35):
        assert np.array_equal(np.tile(g.agent_angle[ix], 32), E.angle_feature((ix % 12) * R30, (ix // 12 - 1) * R30, 128))
    i = int(np.argmax(g.deg))
    for k in range(int(g.deg[i])):
        for hb in (0, 7):
            want = E.angle_feature(g.headings[i][k] - hb * R30, g.elevations[i][k], 128)
            assert np.array_equal(np.tile(g.cand_angle[i, k, hb], 32), want)


def test_union_of_scans():
    """Several scans in one table set (a batch mixes scans, env.py:182-198 keeps one graph per scan): per-component tables equal
    the single-graph ones, cross-scan distances are inf, the teacher never leaves its scan."""
    g1, *_ = scenario(n=14, seed=1)
    g2, *_ = scenario(n=20, seed=2)
    u = NavGraph.union([g1, g2], ["scanA", "scanB"])
    assert u.n == 34 and u.names[0].startswith("scanA_") and u.names[14].startswith("scanB_")
    assert np.array_equal(u.dist64[:14, :14], g1.dist64) and np.array_equal(u.dist64[14:, 14:], g2.dist64)
    assert np.isinf(u.dist64[:14, 14:]).all() and np.isinf(u.dist64[14:, :14]).all()
    assert np.array_equal(u.next_hop[:14, :14], g1.next_hop) and np.array_equal(u.next_hop[14:, 14:], g2.next_hop)
    assert (u.next_hop[:14, 14:] == -1).all()
    assert np.array_equal(u.nbr[14:][u.nbr[14:] >= 0], g2.nbr[g2.nbr >= 0] + 14)
    h = u.hops()
    assert (h[:14, 14:] == -1).all() and np.array_equal(h[14:, 14:], g2.hops())
    s, v, goal = u.sample_episodes(12, 0, min_hops=2, max_hops=5)
    assert ((s < 14) == (goal < 14)).all()                      # start and goal always in the same scan


def test_oracle_submit_mask_blocks_the_way_back():
    """agent_dg.py:834-840 as restated in the oracle: after moving a -> b the candidate leading back to a is masked at b, the
    END slot never is, and the sets only grow."""
    g, rgb, dep, start, view, goal = scenario(n=24, B=4, seed=2)
    env = E.RefStyleEnv("s", features=rgb, dfeatures=dep, **lists(g))
    env.new_episodes([g.names[i] for i in start], view, [g.names[i] for i in goal])
    nc = g.dmax + 1
    visited = [set() for _ in start]
    obs = env.get_obs()
    m0 = E.submit_candidate_mask(obs, visited, nc)
    assert not m0.any()                                   # nothing visited yet except the current viewpoints
    first = [ob["viewpoint"] for ob in obs]
    env.make_equiv_action(np.zeros(len(start), np.int64), obs)          # everybody takes candidate 0
    obs = env.get_obs()
    m1 = E.submit_candidate_mask(obs, visited, nc)
    for i, ob in enumerate(obs):
        back = [k for k, c in enumerate(ob["candidate"]) if c["viewpointId"] == first[i]]
        assert back and all(m1[i, k] for k in back)
        assert not m1[i, len(ob["candidate"]):].any()     # END slot and padding
        assert visited[i] == {first[i], ob["viewpoint"]}
