"""TEST INFRASTRUCTURE ONLY — never imported by the product path (dasa_b200/).

CPU restatement (plain Python + numpy, dict-of-dict state exactly like the reference keeps it) of the environment side of
the agent_dg rollout, used as the oracle for the device-resident environment (dasa_b200/env.py, csrc/env.cu):

  R2RBatch._load_nav_graphs      env.py:182-198   (networkx all_pairs_dijkstra_path / _path_length; restated in
                                                   `all_pairs_paths` following networkx 3.x _dijkstra_multisource, the
                                                   un-vendored third-party routine those two calls run)
  R2RBatch._shortest_path_action env.py:232-238
  R2RBatch.make_candidate        env.py:240-311   (the buffered branch :299-311; the un-buffered branch only fills
                                                   buffered_state_dict from the simulator, which is not available offline)
  R2RBatch._get_obs              env.py:317-358
  utils.angle_feature            utils.py:361-368
  utils.get_point_angle_feature  utils.py:386-405 (simulator heading of view ix = (ix % 12) * 30deg, elevation
                                                   (ix // 12 - 1) * 30deg in discretised mode)
  Seq2SeqAgent._feature_variable/_dfeature_variable/_candidate_variable/get_input_feat   agent_dg.py:286-323
  Seq2SeqAgent._teacher_action   agent_dg.py:325-344
  Seq2SeqAgent.make_equiv_action agent_dg.py:358-391 (net effect on a discretised simulator: viewIndex = the candidate's
                                                   pointId, location = the candidate's viewpoint)
  reward / mask / ended          agent_dg.py:890-935

PINNED: tests/test_oracle_env_vs_reference.py runs the reference's own env.py / agent_dg.py functions (imported unmodified,
driven by a fake discretised simulator, `.cuda()` patched to identity) on the same graphs and compares every tensor
bit-exactly; tests/golden/env_rollout.pt carries reference-generated vectors to the GPU box.
"""
import heapq
import math
from itertools import count

import numpy as np

R30 = math.radians(30)


def angle_feature(heading, elevation, size):
    """utils.py:361-368."""
    return np.array([math.sin(heading), math.cos(heading), math.sin(elevation), math.cos(elevation)] * (size // 4),
                    dtype=np.float32)


def point_angle_feature(base_view, size):
    """utils.py:386-405 with the discretised simulator's headings."""
    feature = np.empty((36, size), np.float32)
    base_heading = (base_view % 12) * R30
    for ix in range(36):
        heading = (ix % 12) * R30 - base_heading
        feature[ix, :] = angle_feature(heading, (ix // 12 - 1) * R30, size)
    return feature


def all_pairs_paths(names, nbrs, weights):
    """dict(nx.all_pairs_dijkstra_path(G)), dict(nx.all_pairs_dijkstra_path_length(G)) (env.py:195-198) for the graph whose
    adjacency (in edge insertion order) is nbrs / weights."""
    paths, dists = {}, {}
    for s in range(len(names)):
        dist, seen, pth = {}, {s: 0}, {s: [s]}
        c = count()
        fringe = [(0, next(c), s)]
        while fringe:
            d, _, v = heapq.heappop(fringe)
            if v in dist:
                continue
            dist[v] = d
            for k, u in enumerate(nbrs[v]):
                vu = dist[v] + weights[v][k]
                if u in dist:
                    continue
                if u not in seen or vu < seen[u]:
                    seen[u] = vu
                    heapq.heappush(fringe, (vu, next(c), u))
                    pth[u] = pth[v] + [u]
        paths[names[s]] = {names[v]: [names[x] for x in p] for v, p in pth.items()}
        dists[names[s]] = {names[v]: d for v, d in dist.items()}
    return paths, dists


class RefStyleEnv:
    """One scan of R2RBatch with its python-dict state; the simulator is reduced to (viewpointId, viewIndex) per agent."""

    def __init__(self, scan, names, nbrs, weights, headings, elevations, points, features, dfeatures, angle_size=128):
        self.scan, self.names, self.angle_size = scan, names, angle_size
        self.features = {scan + "_" + names[i]: features[i] for i in range(len(names))}       # EnvBatch.features
        self.depth_map = {scan + "_" + names[i]: dfeatures[i] for i in range(len(names))}     # Depth_Features.depth_map
        self.buffered_state_dict = {}
        for i, nm in enumerate(names):                                                        # env.py:291-298
            self.buffered_state_dict["%s_%s" % (scan, nm)] = [
                {"normalized_heading": headings[i][k], "elevation": elevations[i][k], "scanId": scan,
                 "viewpointId": names[j], "pointId": int(points[i][k]), "idx": k + 1} for k, j in enumerate(nbrs[i])]
        self.paths, self.distances = {}, {}
        self.paths[scan], self.distances[scan] = all_pairs_paths(names, nbrs, weights)
        self.angle_feature = [point_angle_feature(b, angle_size) for b in range(36)]           # env.py:171
        self.sims, self.batch = [], []

    def new_episodes(self, start_names, start_views, goal_names):
        self.sims = [{"viewpointId": s, "viewIndex": int(v)} for s, v in zip(start_names, start_views)]
        self.batch = [{"path": [s, g]} for s, g in zip(start_names, goal_names)]               # only path[0], path[-1] are read

    def _shortest_path_action(self, vp, goal):                                                 # env.py:232-238
        if vp == goal:
            return goal
        return self.paths[self.scan][vp][goal][1]

    def make_candidate(self, feature, dfeature, scanId, viewpointId, viewId):                  # env.py:299-311
        base_heading = (viewId % 12) * R30
        out = []
        for c in self.buffered_state_dict["%s_%s" % (scanId, viewpointId)]:
            c_new = c.copy()
            ix = c_new["pointId"]
            loc_heading = c_new["normalized_heading"] - base_heading
            c_new["heading"] = loc_heading
            angle_feat = angle_feature(c_new["heading"], c_new["elevation"], self.angle_size)
            c_new["feature"] = np.concatenate((feature[ix], angle_feat), -1)
            c_new["dfeature"] = np.concatenate((dfeature[ix], angle_feat), -1)
            c_new.pop("normalized_heading")
            out.append(c_new)
        return out

    def get_obs(self):                                                                          # env.py:317-358
        obs = []
        for i, sim in enumerate(self.sims):
            long_id = self.scan + "_" + sim["viewpointId"]
            feature, dfeature = self.features[long_id], self.depth_map[long_id]
            base_view_id = sim["viewIndex"]
            goal = self.batch[i]["path"][-1]
            obs.append({
                "scan": self.scan, "viewpoint": sim["viewpointId"], "viewIndex": base_view_id,
                "heading": (base_view_id % 12) * R30, "elevation": (base_view_id // 12 - 1) * R30,
                "candidate": self.make_candidate(feature, dfeature, self.scan, sim["viewpointId"], base_view_id),
                "feature": np.concatenate((feature, self.angle_feature[base_view_id]), -1),
                "dfeature": np.concatenate((dfeature, self.angle_feature[base_view_id]), -1),
                "teacher": self._shortest_path_action(sim["viewpointId"], goal),
                "distance": self.distances[self.scan][sim["viewpointId"]][goal],
            })
        return obs

    def make_equiv_action(self, cpu_a_t, obs):                                                  # agent_dg.py:358-391
        for i, action in enumerate(cpu_a_t):
            if action != -1:
                c = obs[i]["candidate"][action]
                self.sims[i]["viewIndex"] = c["pointId"]          # tune up / down, then turn right until trg_point
                self.sims[i]["viewpointId"] = c["viewpointId"]    # take_action(select_candidate['idx'])


def get_input_feat(obs, rgb_size, angle_size, views=36):
    """agent_dg.py:286-323 -> numpy arrays (the reference then calls torch.from_numpy(...).cuda())."""
    F = rgb_size + angle_size
    input_a_t = np.zeros((len(obs), angle_size), np.float32)
    for i, ob in enumerate(obs):
        input_a_t[i] = angle_feature(ob["heading"], ob["elevation"], angle_size)
    f_t = np.empty((len(obs), views, F), np.float32)
    d_t = np.empty((len(obs), views, F), np.float32)
    for i, ob in enumerate(obs):
        f_t[i], d_t[i] = ob["feature"], ob["dfeature"]
    leng = [len(ob["candidate"]) + 1 for ob in obs]
    cand = np.zeros((len(obs), max(leng), F), np.float32)
    cand_d = np.zeros((len(obs), max(leng), F), np.float32)
    for i, ob in enumerate(obs):
        for j, c in enumerate(ob["candidate"]):
            cand[i, j, :], cand_d[i, j, :] = c["feature"], c["dfeature"]
    return input_a_t, f_t, d_t, cand, cand_d, leng


def teacher_action(obs, ended, ignoreid=-100):
    """agent_dg.py:325-344."""
    a = np.zeros(len(obs), dtype=np.int64)
    for i, ob in enumerate(obs):
        if ended[i]:
            a[i] = ignoreid
        else:
            for k, candidate in enumerate(ob["candidate"]):
                if candidate["viewpointId"] == ob["teacher"]:
                    a[i] = k
                    break
            else:
                assert ob["teacher"] == ob["viewpoint"]
                a[i] = len(ob["candidate"])
    return a


def submit_candidate_mask(obs, visited, max_cand):
    """agent_dg.py:834-840 (args.submit, "avoiding cyclic path"): the current viewpoint joins visited[ob_id]; a candidate whose
    viewpointId was visited is masked. `visited`: list of python sets, mutated like the reference's. Returns bool [B, max_cand]."""
    mask = np.zeros((len(obs), max_cand), dtype=bool)
    for ob_id, ob in enumerate(obs):
        visited[ob_id].add(ob["viewpoint"])
        for c_id, c in enumerate(ob["candidate"]):
            if c["viewpointId"] in visited[ob_id]:
                mask[ob_id][c_id] = True
    return mask


def env_action(a_t, cand_leng, ignoreid=-100):
    """agent_dg.py:890-893: <end> and ignore become -1."""
    cpu_a_t = np.array(a_t, dtype=np.int64).copy()
    for i, next_id in enumerate(cpu_a_t):
        if next_id == (cand_leng[i] - 1) or next_id == ignoreid:
            cpu_a_t[i] = -1
    return cpu_a_t


def reward_update(cpu_a_t, obs_after, ended, last_dist):
    """agent_dg.py:900-932; mutates ended and last_dist like the reference loop. Returns (reward, mask) float32 [B]."""
    B = len(obs_after)
    dist = np.zeros(B, np.float32)
    reward = np.zeros(B, np.float32)
    mask = np.ones(B, np.float32)
    for i, ob in enumerate(obs_after):
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
    return reward, mask


def rollout(env, T, rgb_size, angle_size, actions=None, ignoreid=-100):
    """The environment side of vl_rollout for T steps. actions=None: feedback='teacher' (a_t = target); otherwise
    actions[t] (int64 [B]) plays the sampled / greedy action. Returns per-step dicts of numpy arrays."""
    obs = env.get_obs()
    B = len(obs)
    last_dist = np.zeros(B, np.float32)
    for i, ob in enumerate(obs):
        last_dist[i] = ob["distance"]
    ended = np.array([False] * B)
    steps = []
    for t in range(T):
        input_a_t, f_t, d_t, cand, cand_d, leng = get_input_feat(obs, rgb_size, angle_size)
        target = teacher_action(obs, ended, ignoreid)
        a_t = target if actions is None else np.asarray(actions[t])
        cpu_a_t = env_action(a_t, leng, ignoreid)
        dist_before = np.array([ob["distance"] for ob in obs], np.float32)
        env.make_equiv_action(cpu_a_t, obs)
        obs = env.get_obs()
        reward, mask = reward_update(cpu_a_t, obs, ended, last_dist)
        steps.append({"input_a_t": input_a_t, "f_t": f_t, "d_t": d_t, "cand_feat": cand, "cand_dfeat": cand_d,
                      "cand_leng": np.array(leng, np.int32), "target": target, "action": np.array(a_t, np.int64),
                      "dist": dist_before, "reward": reward, "mask": mask, "ended": ended.copy(),
                      "viewpoint": [ob["viewpoint"] for ob in obs], "viewIndex": np.array([ob["viewIndex"] for ob in obs])})
    return steps


def from_shortest_path(env, rgb_size, angle_size, ignoreid=-100):
    """Speaker.from_shortest_path (speaker.py:163-198) + Speaker._teacher_action / _candidate_variable (:137-161): follow the
    teacher until every episode stopped. Returns (img_feats [B, L, 36, F], can_feats [B, L, F], length [B])."""
    obs = env.get_obs()
    B, F = len(obs), rgb_size + angle_size
    ended = np.array([False] * B)
    length = np.zeros(B, np.int64)
    img_feats, can_feats = [], []
    while not ended.all():
        f_t = np.empty((B, 36, F), np.float32)
        for i, ob in enumerate(obs):
            f_t[i] = ob["feature"]
        img_feats.append(f_t)
        ta = teacher_action(obs, ended, ignoreid)
        for i, act in enumerate(ta):
            if act < 0 or act == len(obs[i]["candidate"]):
                ta[i] = -1
        cf = np.zeros((B, F), np.float32)
        for i, (ob, act) in enumerate(zip(obs, ta)):
            if act != -1:
                cf[i, :] = ob["candidate"][act]["feature"]
        can_feats.append(cf)
        env.make_equiv_action(ta, obs)
        length += (1 - ended)
        ended[:] = np.logical_or(ended, (ta == -1))
        obs = env.get_obs()
    return np.stack(img_feats, 1), np.stack(can_feats, 1), length
