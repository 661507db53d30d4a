"""GPU tier: the device-resident environment (dasa_b200/env.py, csrc/env.cu) against the CPU oracle (oracle/env_restated.py,
pinned against the reference's env.py / agent_dg.py) and against the reference-generated fixture tests/golden/env_rollout.pt.
Everything here is integer / table-lookup / copy work, so the bar is BIT-EXACT."""
import os
from dataclasses import replace
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from dasa_b200 import synth
from dasa_b200.config import FULL, SMALL
from dasa_b200.navgraph import NavGraph
from oracle import env_restated as E
from oracle import restated as R
from tests.envcase import lists, random_actions, scenario

pytestmark = pytest.mark.gpu
DEV = "cuda"
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "env_rollout.pt")
BUF_KEYS = ("input_a_t", "f_t", "d_t", "cand_feat", "cand_dfeat", "cand_leng", "target", "dist")


def _cfg(C):
    return replace(SMALL, rgb_size=C)


def _device_rollout(g, rgb, dep, start, view, goal, T, actions, cfg):
    from dasa_b200.env import DeviceEnv
    env = DeviceEnv(g, rgb, dep, cfg, DEV).reset(start, view, goal)
    buf = env.alloc(T)
    reward = torch.empty(T, env.B, device=DEV)
    mask = torch.empty(T, env.B, device=DEV)
    ended, vps, views = [], [], []
    for t in range(T):
        env.observe(buf, t)
        a = buf["target"][t] if actions is None else torch.as_tensor(actions[t]).to(DEV)
        env.step(a, reward[t], mask[t])
        ended.append(env.ended.clone())
        vps.append(env.vp.clone())
        views.append(env.view.clone())
    torch.cuda.synchronize()
    return env, buf, reward, mask, ended, vps, views


def _compare(steps, g, buf, reward, mask, ended, vps, views, T, name_index=None):
    for t in range(T):
        s = steps[t]
        nc_ref = np.asarray(s["cand_feat"]).shape[1]
        for k in BUF_KEYS:
            if k not in s:
                continue
            got = buf[k][t].cpu().numpy()
            want = np.asarray(s[k])
            if k in ("cand_feat", "cand_dfeat"):
                assert not got[:, nc_ref:].any(), "step %d %s: padding slots must be zero" % (t, k)
                got = got[:, :nc_ref]
            assert np.array_equal(got, want), "step %d %s" % (t, k)
        assert np.array_equal(reward[t].cpu().numpy(), np.asarray(s["reward"])), "step %d reward" % t
        assert np.array_equal(mask[t].cpu().numpy(), np.asarray(s["mask"])), "step %d mask" % t
        assert np.array_equal(ended[t].cpu().numpy().astype(bool), np.asarray(s["ended"]).astype(bool)), "step %d ended" % t
        assert np.array_equal(views[t].cpu().numpy(), np.asarray(s["viewIndex"])), "step %d viewIndex" % t
        if "vp" in s:
            assert np.array_equal(vps[t].cpu().numpy(), np.asarray(s["vp"]))
        else:
            assert [g.names[i] for i in vps[t].cpu().tolist()] == s["viewpoint"]


@pytest.mark.parametrize("seed,n,B,C", [(0, 24, 5, 32), (1, 40, 9, 64), (2, 30, 3, 2048)])
@pytest.mark.parametrize("closed_loop", [False, True])
def test_env_matches_oracle(seed, n, B, C, closed_loop):
    T = 8
    g, rgb, dep, start, view, goal = scenario(n=n, B=B, T=T, C=C, seed=seed)
    cfg = _cfg(C)
    actions = None
    if closed_loop:
        for s2 in range(seed, seed + 50):                      # a random walk without a zero-progress move (reference raises)
            actions = random_actions(g, start, T, s2)
            ora = E.RefStyleEnv("s", features=rgb, dfeatures=dep, **lists(g))
            ora.new_episodes([g.names[i] for i in start], view, [g.names[i] for i in goal])
            try:
                steps = E.rollout(ora, T, C, cfg.angle_size, actions)
                break
            except NameError:
                continue
    else:
        ora = E.RefStyleEnv("s", features=rgb, dfeatures=dep, **lists(g))
        ora.new_episodes([g.names[i] for i in start], view, [g.names[i] for i in goal])
        steps = E.rollout(ora, T, C, cfg.angle_size, None)
    env, buf, reward, mask, ended, vps, views = _device_rollout(g, rgb, dep, start, view, goal, T, actions, cfg)
    _compare(steps, g, buf, reward, mask, ended, vps, views, T)
    env.check()


def test_env_matches_reference_golden():
    """Vectors produced by the reference's own env.py / agent_dg.py helpers (oracle/make_golden_env.py)."""
    gold = torch.load(GOLD, weights_only=False)
    G = gold["graph"]
    g = NavGraph(G["nbrs"], G["weights"], G["headings"], G["elevations"], G["points"], G["names"])
    cfg = _cfg(gold["C"])
    T = gold["T"]
    for key, actions in (("teacher", None), ("closed", [a.numpy() for a in gold["actions"]])):
        steps = [{k: (v.numpy() if torch.is_tensor(v) else v) for k, v in s.items()} for s in gold[key]]
        env, buf, reward, mask, ended, vps, views = _device_rollout(g, gold["rgb"].numpy(), gold["dep"].numpy(), gold["start"].numpy(),
                                                                    gold["view"].numpy(), gold["goal"].numpy(), T, actions, cfg)
        _compare(steps, g, buf, reward, mask, ended, vps, views, T)


def test_env_error_flags():
    from dasa_b200.env import DeviceEnv
    g, rgb, dep, start, view, goal = scenario(seed=4)
    env = DeviceEnv(g, rgb, dep, _cfg(32), DEV).reset(start, view, goal)
    bad = torch.full((env.B,), 99, dtype=torch.int64, device=DEV)
    env.step(bad)
    with pytest.raises(RuntimeError):
        env.check()
    with pytest.raises(RuntimeError):
        DeviceEnv(g, rgb, dep, _cfg(32), "cpu")


class _OracleEpisodes:
    """Oracle-side episodes assembled from oracle/env_restated.rollout steps (what R.teacher_rollout / R.sample_rollout read)."""

    def __init__(self, steps, instr, cfg):
        self.seq, self.seq_mask, self.seq_lengths = instr
        self.B, self.T, self.cfg = len(steps[0]["target"]), len(steps), cfg
        self.steps = steps
        self.dist = torch.stack([torch.from_numpy(s["dist"]) for s in steps])

    def step(self, t):
        s = self.steps[t]
        return (torch.from_numpy(s["input_a_t"]), torch.from_numpy(s["f_t"]), torch.from_numpy(s["d_t"]),
                torch.from_numpy(s["cand_feat"]), torch.from_numpy(s["cand_dfeat"]), torch.from_numpy(s["cand_leng"]),
                torch.from_numpy(s["target"]))


@pytest.mark.parametrize("schedule", ["batched", "sequential"])
def test_teacher_rollout_on_device_env(schedule):
    """agent_dg teacher-forced rollout with the observations produced on the device == the oracle policy over the oracle
    environment (loss, logits, greedy actions, a weight gradient)."""
    from dasa_b200.env import DeviceEnv
    from dasa_b200.rollout import NavPolicy
    cfg, B, T = SMALL, 4, 5
    g, rgb, dep, start, view, goal = scenario(n=20, B=B, T=T, C=cfg.rgb_size, seed=6)
    instr = synth.instructions(B, cfg, 6)
    ora = E.RefStyleEnv("s", features=rgb, dfeatures=dep, **lists(g))
    ora.new_episodes([g.names[i] for i in start], view, [g.names[i] for i in goal])
    oep = _OracleEpisodes(E.rollout(ora, T, cfg.rgb_size, cfg.angle_size, None), instr, cfg)
    st = synth.policy_state(cfg, 0)
    ost = {grp: {k: v.clone().requires_grad_(True) for k, v in d.items()} for grp, d in st.items()}
    loss_ref, logits_ref, act_ref = R.teacher_rollout(ost, cfg, oep, T)
    loss_ref.backward()
    pol = NavPolicy(cfg, st, DEV).eval()
    env = DeviceEnv(g, rgb, dep, cfg, DEV).reset(start, view, goal)
    ep = env.teacher_episodes(T, instr)
    loss, logits, act = pol.teacher_rollout(ep, T, schedule=schedule)
    pol.backward(loss)
    torch.cuda.synchronize()
    assert abs(float(loss) - float(loss_ref)) <= 1e-4 * max(1e-6, abs(float(loss_ref)))
    for t in range(T):
        nc = logits_ref[t].shape[1]
        a, b = logits[t].cpu()[:, :nc], logits_ref[t]
        fin = torch.isfinite(b)
        assert torch.equal(torch.isfinite(a), fin)
        assert float((a[fin] - b[fin]).abs().max()) <= 1e-4 * float(b[fin].abs().max())
        assert torch.equal(act[t].cpu(), act_ref[t])
    gw, gr = pol.decoder.lstm.weight_hh.grad.cpu(), ost["decoder"]["lstm.weight_hh"].grad
    assert float((gw - gr).abs().max()) <= 1e-3 * float(gr.abs().max())
    env.check()


def test_sample_rollout_on_live_env():
    """Sampled-feedback A2C rollout in closed loop with the device environment (injected actions) == oracle policy + oracle env."""
    from dasa_b200.env import DeviceEnv
    from dasa_b200.rollout import NavPolicy
    cfg, B, T = SMALL, 4, 5
    g, rgb, dep, start, view, goal = scenario(n=20, B=B, T=T, C=cfg.rgb_size, seed=8)
    instr = synth.instructions(B, cfg, 8)
    for s2 in range(50):
        actions = random_actions(g, start, T, s2) + [np.array([int(cfg.ignore_id)] * B, np.int64)]
        ora = E.RefStyleEnv("s", features=rgb, dfeatures=dep, **lists(g))
        ora.new_episodes([g.names[i] for i in start], view, [g.names[i] for i in goal])
        try:
            steps = E.rollout(ora, T + 1, cfg.rgb_size, cfg.angle_size, actions)
            break
        except NameError:
            continue
    oep = _OracleEpisodes(steps, instr, cfg)
    st = synth.policy_state(cfg, 1)
    ost = {grp: {k: v.clone().requires_grad_(True) for k, v in d.items()} for grp, d in st.items()}
    acts_t = [torch.from_numpy(a) for a in actions[:T]]
    loss_ref, info_ref = R.sample_rollout(ost, cfg, oep, T, acts_t)
    loss_ref.backward()
    pol = NavPolicy(cfg, st, DEV).eval()
    env = DeviceEnv(g, rgb, dep, cfg, DEV).reset(start, view, goal)
    ep = env.live_episodes(T, instr)
    loss, info = pol.sample_rollout(ep, T, actions_in=[a.to(DEV) for a in acts_t])
    pol.backward(loss)
    torch.cuda.synchronize()
    assert torch.equal(info["reward"].cpu(), torch.stack(info_ref["rewards"]))
    assert torch.equal(info["mask"].cpu(), torch.stack(info_ref["masks"]))
    assert torch.equal(info["ended"].cpu().bool(), info_ref["ended"].bool())
    assert abs(float(loss) - float(loss_ref)) <= 2e-4 * max(1e-6, abs(float(loss_ref)))
    gw, gr = pol.decoder.lstm.weight_hh.grad.cpu(), ost["decoder"]["lstm.weight_hh"].grad
    assert float((gw - gr).abs().max()) <= 1e-3 * float(gr.abs().max())
    assert torch.equal(ep.traj[1:T + 1].cpu(), torch.tensor([[g.names.index(v) for v in s["viewpoint"]] for s in steps[:T]],
                                                              dtype=torch.int32))
    env.check()


def test_shortest_path_features_for_the_speaker():
    """DeviceEnv.shortest_path_features == Speaker.from_shortest_path restated over the oracle environment (bit-exact)."""
    from dasa_b200.env import DeviceEnv
    C = 32
    g, rgb, dep, start, view, goal = scenario(n=30, B=6, C=C, seed=11)
    cfg = _cfg(C)
    ora = E.RefStyleEnv("s", features=rgb, dfeatures=dep, **lists(g))
    ora.new_episodes([g.names[i] for i in start], view, [g.names[i] for i in goal])
    img, can, length = E.from_shortest_path(ora, C, cfg.angle_size)
    env = DeviceEnv(g, rgb, dep, cfg, DEV).reset(start, view, goal)
    (img_d, can_d), length_d = env.shortest_path_features(12)
    assert np.array_equal(length_d.cpu().numpy(), length)
    assert np.array_equal(can_d.cpu().numpy(), can)
    # panoramas: the reference keeps observing after an episode stopped (same viewpoint); compare all L steps
    assert np.array_equal(img_d.cpu().numpy(), img)


def test_env_properties_at_full_size():
    """Size-independent properties at BASELINE geometry (36 x 2176 views, 512 episodes, a 512-viewpoint graph), where the Python
    oracle would take minutes: the teacher reaches every goal in exactly hops(start, goal) moves, rewards are +1 per move and +2
    for the stop, masks switch off after the stop, trajectories follow graph edges, and the assembled panoramas / candidates
    equal a torch gather from the feature banks."""
    from dasa_b200.env import DeviceEnv
    cfg = FULL
    g = NavGraph.synthetic(512, seed=3)
    gen = torch.Generator().manual_seed(5)
    rgb = synth.resnet_like((g.n, cfg.views, cfg.rgb_size), gen)
    dep = synth.resnet_like((g.n, cfg.views, cfg.rgb_size), gen)
    B, T = 512, 10
    start, view, goal = g.sample_episodes(B, seed=1, min_hops=3, max_hops=7)
    hops = torch.from_numpy(g.hops()[start, goal])
    env = DeviceEnv(g, rgb, dep, cfg, DEV).reset(start, view, goal)
    buf = env.alloc(T)
    reward = torch.empty(T, B, device=DEV)
    mask = torch.empty(T, B, device=DEV)
    traj = torch.empty(T + 1, B, dtype=torch.int32, device=DEV)
    traj[0].copy_(env.vp)
    views_before = []
    for t in range(T):
        views_before.append(env.view.clone())
        env.observe(buf, t)
        env.step(buf["target"][t], reward[t], mask[t], traj_vp=traj[t + 1])
    env.check()
    traj_c, reward_c, mask_c = traj.cpu().long(), reward.cpu(), mask.cpu()
    tgt, leng = buf["target"].cpu(), buf["cand_leng"].cpu().long()
    assert torch.equal(traj_c[-1], torch.from_numpy(goal).long())                      # every goal reached
    for t in range(T):
        moving = t < hops
        stopping = t == hops
        after = t > hops
        assert torch.equal(reward_c[t][moving], torch.ones(int(moving.sum())))         # each teacher move shortens the path
        assert torch.equal(reward_c[t][stopping], torch.full((int(stopping.sum()),), 2.0))   # stop within 3 m of the goal
        assert float(reward_c[t][after].abs().max() if after.any() else 0.0) == 0.0
        assert torch.equal(mask_c[t], (~after).float())
        assert torch.equal(tgt[t][stopping], leng[t][stopping] - 1)                     # STOP = the END row
        assert bool((tgt[t][after] == cfg.ignore_id).all())
        # moves follow graph edges
        nbr = torch.from_numpy(g.nbr).long()
        ok = (nbr[traj_c[t]] == traj_c[t + 1][:, None]).any(1) | (traj_c[t] == traj_c[t + 1])
        assert bool(ok.all())
    assert torch.equal(buf["dist"][0].cpu(), torch.from_numpy(g.dist[start, goal]))
    # panoramas and candidates of step 2 against a torch gather from the banks
    t = 2
    vp = traj[t].long()
    C = cfg.rgb_size
    assert torch.equal(buf["f_t"][t][..., :C], env.rgb_bank[vp])
    assert torch.equal(buf["d_t"][t][..., :C], env.dep_bank[vp])
    assert torch.equal(buf["f_t"][t][..., C:], buf["d_t"][t][..., C:])
    va = torch.from_numpy(g.view_angle).to(DEV)[(views_before[t] % 12).long()]           # [B, 36, 4]
    assert torch.equal(buf["f_t"][t][..., C:], va.repeat(1, 1, cfg.angle_size // 4))
    deg = torch.from_numpy(g.deg).to(DEV)[vp].long()
    k = torch.arange(env.nc, device=DEV)[None, :]
    live = k < deg[:, None]
    pt = torch.from_numpy(g.nbr_point).to(DEV).long()[vp]                               # [B, dmax]
    want = env.rgb_bank[vp[:, None].expand(-1, env.dmax), pt]                            # [B, dmax, C]
    got = buf["cand_feat"][t][:, :env.dmax, :C]
    assert torch.equal(got[live[:, :env.dmax]], want[live[:, :env.dmax]])
    assert float(buf["cand_feat"][t][~live].abs().max()) == 0.0                         # END row and padding are zeros


def test_env_on_a_union_of_scans():
    """A batch that mixes episodes from two scans (disjoint union of two graphs in one table set): bit-exact vs the oracle."""
    g1, *_ = scenario(n=14, seed=1)
    g2, *_ = scenario(n=20, seed=2)
    u = NavGraph.union([g1, g2], ["scanA", "scanB"])
    T, C = 7, 32
    g, rgb, dep, start, view, goal = scenario(B=8, T=T, C=C, seed=2, graph=u)
    assert (start < 14).any() and (start >= 14).any()
    cfg = _cfg(C)
    ora = E.RefStyleEnv("s", features=rgb, dfeatures=dep, **lists(g))
    ora.new_episodes([g.names[i] for i in start], view, [g.names[i] for i in goal])
    steps = E.rollout(ora, T, C, cfg.angle_size, None)
    env, buf, reward, mask, ended, vps, views = _device_rollout(g, rgb, dep, start, view, goal, T, None, cfg)
    _compare(steps, g, buf, reward, mask, ended, vps, views, T)
    env.check()


@pytest.mark.parametrize("seed,n,B", [(0, 24, 5), (3, 40, 11)])
def test_submit_visited_mask_matches_oracle(seed, n, B):
    """--submit "avoiding cyclic path" (agent_dg.py:834-840): the device bitmap + mask kernel against the oracle's python sets on a
    random closed-loop walk; masked logits are -inf exactly where the oracle masks, untouched elsewhere."""
    from dasa_b200.env import DeviceEnv
    T, C = 9, 32
    g, rgb, dep, start, view, goal = scenario(n, B, T, C, seed)
    cfg = _cfg(C)
    acts = random_actions(g, start, T, seed, stop_prob=0.1)
    env = DeviceEnv(g, rgb, dep, cfg, DEV).reset(start, view, goal)
    ref = E.RefStyleEnv("scan0", **lists(g), features=rgb, dfeatures=dep, angle_size=cfg.angle_size)
    ref.new_episodes([g.names[i] for i in start], view, [g.names[i] for i in goal])
    visited = [set() for _ in range(B)]
    gen = torch.Generator().manual_seed(seed)
    for t in range(T):
        obs = ref.get_obs()
        want = E.submit_candidate_mask(obs, visited, env.nc)
        logit = torch.randn(B, env.nc, generator=gen).to(DEV)
        before = logit.clone()
        got = env.visited_mask(logit, want_mask=True)
        assert np.array_equal(got.cpu().numpy().astype(bool), want), "step %d mask" % t
        w = torch.from_numpy(want).to(DEV)
        assert torch.isneginf(logit[w]).all() and torch.equal(logit[~w], before[~w]), "step %d logits" % t
        a = np.asarray(acts[t])
        leng = [len(ob["candidate"]) + 1 for ob in obs]
        ref.make_equiv_action(E.env_action(a, leng), obs)
        env.step(torch.as_tensor(a).to(DEV))
    torch.cuda.synchronize()
    if t > 2:
        assert any(len(v) > 2 for v in visited)


def test_greedy_rollout_submit_never_revisits():
    """feedback='argmax' with args.submit (agent_dg.py:834-840, 871-875) in closed loop on the device environment: with the visited
    mask no episode ever steps onto a viewpoint it has been to (it may only stop), and the masked logits are -inf exactly for the
    candidates leading to its own past trajectory; without the mask the same policy does revisit (so the property is not vacuous)."""
    from dasa_b200.env import DeviceEnv
    from dasa_b200.rollout import NavPolicy
    cfg, B, T = SMALL, 8, 8
    g, rgb, dep, start, view, goal = scenario(n=12, B=B, T=T, C=cfg.rgb_size, seed=4)
    instr = synth.instructions(B, cfg, 4)
    pol = NavPolicy(cfg, synth.policy_state(cfg, 3), DEV).eval()
    revisits = {}
    for submit in (False, True):
        env = DeviceEnv(g, rgb, dep, cfg, DEV).reset(start, view, goal)
        ep = env.live_episodes(T, instr)
        actions, logits = pol.greedy_rollout(ep, T, submit=submit)
        torch.cuda.synchronize()
        env.check()
        traj = ep.traj[:T + 1].cpu().numpy()                  # [T + 1, B] viewpoints, traj[0] = start
        n_rev = 0
        for b in range(B):
            seen = {int(traj[0, b])}
            for t in range(T):
                a, v = int(actions[t][b]), int(traj[t, b])
                moved = a != int(g.deg[v])
                if submit:                                    # mask check against the trajectory so far
                    blocked = torch.isneginf(logits[t][b, :int(g.deg[v])]).cpu().numpy()
                    want = np.array([int(g.nbr[v, k]) in seen for k in range(int(g.deg[v]))])
                    assert np.array_equal(blocked, want), (b, t)
                if moved:
                    n_rev += int(traj[t + 1, b]) in seen
                    seen.add(int(traj[t + 1, b]))
        revisits[submit] = n_rev
    assert revisits[True] == 0
    assert revisits[False] > 0
