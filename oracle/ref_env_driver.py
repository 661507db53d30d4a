"""TEST INFRASTRUCTURE ONLY (build container: needs /root/reference). Runs the UNMODIFIED reference environment code —
env.R2RBatch._get_obs / make_candidate / _shortest_path_action (env.py:232-358) and Seq2SeqAgent.get_input_feat /
_teacher_action / make_equiv_action (agent_dg.py:286-391) — on a graph given as plain lists, so that oracle/env_restated.py can
be pinned against it and golden vectors can be generated (oracle/make_golden_env.py).

How the real env.py is imported offline (oracle/shims/env.py normally shadows it):
  * it is loaded from its file under the module name `ref_env_real`;
  * env.py:20-31 np.load()s args.depth_index_file / depth_value_file at import -> two empty .npy files in a scratch dir;
  * MatterSim comes from oracle/shims (utils.get_all_point_angle_feature only steps through the 36 views);
  * R2RBatch.__init__ needs the datasets, so the instance is created with object.__new__ and given exactly the attributes the
    methods above read: env (sims / features), batch, buffered_state_dict, angle_feature, paths, distances;
    paths / distances come from networkx exactly as env.py:195-198 computes them;
  * the simulator is oracle/sim_stub.GraphSim; torch.Tensor.cuda is patched to identity while the agent helpers run.
"""
import contextlib
import importlib.util
import io
import os
import sys
import tempfile
import types

import numpy as np
import torch

from . import load_reference
from .sim_stub import GraphSim

_real_env = None


def load_real_env():
    global _real_env
    if _real_env is not None:
        return _real_env
    with contextlib.redirect_stdout(io.StringIO()):
        ref = load_reference.load()
    scratch = tempfile.mkdtemp(prefix="dasa_env_")
    np.save(os.path.join(scratch, "idx.npy"), np.zeros((0, 2)))
    np.save(os.path.join(scratch, "val.npy"), np.zeros((0, 4)))
    ref.args.depth_index_file = os.path.join(scratch, "idx.npy")
    ref.args.depth_value_file = os.path.join(scratch, "val.npy")
    spec = importlib.util.spec_from_file_location("ref_env_real", os.path.join(load_reference.REFERENCE_SRC, "env.py"))
    mod = importlib.util.module_from_spec(spec)
    old_cwd = os.getcwd()
    os.chdir(scratch)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            spec.loader.exec_module(mod)
    finally:
        os.chdir(old_cwd)
    _real_env = (ref, mod)
    return _real_env


@contextlib.contextmanager
def cuda_is_identity():
    old = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.Tensor.cuda = old


class ReferenceEnv:
    """The reference's R2RBatch + Seq2SeqAgent helpers over one graph (lists as taken by dasa_b200.navgraph.NavGraph)."""

    def __init__(self, scan, names, nbrs, weights, headings, elevations, points, features, dfeatures, rgb_size, angle_size=128):
        import networkx as nx
        self.ref, self.env_mod = load_real_env()
        ref, env_mod = self.ref, self.env_mod
        assert ref.args.angle_feat_size == angle_size
        ref.args.dfeatures = "imagenet"
        self.scan, self.names, self.nbrs = scan, names, nbrs
        G = nx.Graph()                                            # edges in adjacency order, like utils.load_nav_graphs
        for i in range(len(names)):
            for k, j in enumerate(nbrs[i]):
                G.add_edge(names[i], names[j], weight=weights[i][k])
        rb = object.__new__(env_mod.R2RBatch)
        rb.paths = {scan: dict(nx.all_pairs_dijkstra_path(G))}                   # env.py:195
        rb.distances = {scan: dict(nx.all_pairs_dijkstra_path_length(G))}        # env.py:198
        with contextlib.redirect_stdout(io.StringIO()):
            rb.angle_feature = ref.utils.get_all_point_angle_feature()           # env.py:171
        rb.buffered_state_dict = {}
        for i, nm in enumerate(names):     # what the un-buffered branch would have stored (env.py:291-298)
            rb.buffered_state_dict["%s_%s" % (scan, nm)] = [
                {"normalized_heading": headings[i][k], "elevation": elevations[i][k], "scanId": scan, "viewpointId": names[j],
                 "pointId": int(points[i][k]), "idx": k + 1} for k, j in enumerate(nbrs[i])]
        rb.sim = None
        feats = {scan + "_" + names[i]: features[i] for i in range(len(names))}
        env_mod.depth_features.depth_map = {scan + "_" + names[i]: dfeatures[i] for i in range(len(names))}
        eb = object.__new__(env_mod.EnvBatch)
        eb.features, eb.sims = feats, []
        rb.env = eb
        self.rb = rb
        ag = object.__new__(ref.agent_dg.Seq2SeqAgent)
        ag.env, ag.feature_size = rb, rgb_size
        self.agent = ag
        self.rgb_size, self.angle_size = rgb_size, angle_size

    def new_episodes(self, start_names, start_views, goal_names):
        self.rb.env.sims = [GraphSim(self.scan, self.names, self.nbrs) for _ in start_names]
        for sim, s, v in zip(self.rb.env.sims, start_names, start_views):
            sim.newEpisode(self.scan, s, (int(v) % 12) * (np.pi * 2.0 / 12), (int(v) // 12 - 1) * (np.pi / 6.0))
        self.rb.batch = [{"instr_id": "i%d" % i, "instructions": "", "path": [s, g], "path_id": i}
                         for i, (s, g) in enumerate(zip(start_names, goal_names))]

    def rollout(self, T, actions=None, ignoreid=-100):
        """Environment side of vl_rollout (agent_dg.py:692-935) with the reference's own helpers. Same return layout as
        oracle.env_restated.rollout."""
        ag, rb = self.agent, self.rb
        B = len(rb.batch)
        with cuda_is_identity():
            obs = np.array(rb._get_obs())
            last_dist = np.zeros(B, np.float32)
            for i, ob in enumerate(obs):
                last_dist[i] = ob["distance"]
            ended = np.array([False] * B)
            steps = []
            for t in range(T):
                input_a_t, f_t, d_t, cand, cand_d, leng = ag.get_input_feat(obs)
                target = ag._teacher_action(obs, ended)
                a_t = target if actions is None else torch.as_tensor(np.asarray(actions[t]), dtype=torch.int64)
                cpu_a_t = a_t.cpu().numpy().copy()
                for i, next_id in enumerate(cpu_a_t):                              # agent_dg.py:891-893
                    if next_id == (leng[i] - 1) or next_id == ignoreid:
                        cpu_a_t[i] = -1
                dist_before = np.array([ob["distance"] for ob in obs], np.float32)
                ag.make_equiv_action(cpu_a_t, obs, None, None)
                obs = np.array(rb._get_obs())
                # agent_dg.py:897-932
                dist = np.zeros(B, np.float32)
                reward = np.zeros(B, np.float32)
                mask = np.ones(B, np.float32)
                for i, ob in enumerate(obs):
                    dist[i] = ob["distance"]
                    if ended[i]:
                        reward[i] = 0.
                        mask[i] = 0.
                    else:
                        action_idx = cpu_a_t[i]
                        if action_idx == -1:
                            reward[i] = 2. if dist[i] < 3 else -2.
                        else:
                            reward[i] = - (dist[i] - last_dist[i])
                            if reward[i] > 0:
                                reward[i] = 1
                            elif reward[i] < 0:
                                reward[i] = -1
                            else:
                                raise NameError("The action doesn't change the move")
                last_dist[:] = dist
                ended[:] = np.logical_or(ended, (cpu_a_t == -1))
                steps.append({"input_a_t": input_a_t.numpy(), "f_t": f_t.numpy(), "d_t": d_t.numpy(), "cand_feat": cand.numpy(),
                              "cand_dfeat": cand_d.numpy(), "cand_leng": np.array(leng, np.int32), "target": target.numpy(),
                              "action": a_t.numpy().astype(np.int64), "dist": dist_before, "reward": reward, "mask": mask,
                              "ended": ended.copy(), "viewpoint": [ob["viewpoint"] for ob in obs],
                              "viewIndex": np.array([ob["viewIndex"] for ob in obs])})
        return steps
