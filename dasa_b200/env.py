"""Device-resident navigation environment: R2RBatch._get_obs / make_candidate (env.py:240-358), the
agent's get_input_feat / _candidate_variable / _teacher_action (agent_dg.py:300-344), make_equiv_action (:358-391) and the
reward bookkeeping (:890-935) as two kernels (csrc/env.cu) over graph tables + RGB / depth feature banks that stay in HBM.

The reference keeps the feature store in host RAM, rebuilds every observation with numpy loops and copies ~16 MB to the GPU
per step, then syncs on `a_t.cpu()` to drive the simulator. Here the whole store is HBM-resident (10 567 viewpoints x 36 views
x 2048 floats x 2 banks = 6.2 GB, a fraction of the 180 GB), an observation is one gather kernel and an environment step
one tiny kernel: nothing crosses PCIe inside a rollout and the rollout stays CUDA-graph-capturable.

There is no CPU fallback: tables are uploaded once, all per-step work runs through the C ABI (include/dasa_b200.h).
"""
import numpy as np
import torch

from . import ops
from .config import FULL, PolicyConfig
from .navgraph import NavGraph

FIELDS = ("input_a_t", "f_t", "d_t", "cand_feat", "cand_dfeat", "cand_leng", "target")


class DeviceEnv:
    def __init__(self, graph: NavGraph, rgb_bank, dep_bank, cfg: PolicyConfig = FULL, device="cuda"):
        """rgb_bank / dep_bank: [n_vp, 36, C] float32 (host or device) — ResNet-152 RGB features and the depth-image features
        of every viewpoint (env.py:20-31, 78-113)."""
        if not str(device).startswith("cuda"):
            raise RuntimeError("DeviceEnv runs on a CUDA device only (no CPU path)")
        assert tuple(rgb_bank.shape) == (graph.n, cfg.views, cfg.rgb_size) and rgb_bank.shape == dep_bank.shape
        self.graph, self.cfg, self.device = graph, cfg, device
        self.n_vp, self.dmax, self.nc = graph.n, graph.dmax, graph.dmax + 1
        dev = device
        self.rgb_bank = torch.as_tensor(rgb_bank, dtype=torch.float32).to(dev).contiguous()
        self.dep_bank = torch.as_tensor(dep_bank, dtype=torch.float32).to(dev).contiguous()
        for k in ("nbr", "nbr_point", "deg", "cand_angle", "view_angle", "agent_angle", "dist", "next_hop"):
            setattr(self, "t_" + k, torch.from_numpy(np.ascontiguousarray(getattr(graph, k))).to(dev))
        self.err = torch.zeros(1, dtype=torch.int32, device=dev)
        self.B = 0

    # ------------------------------------------------------------------------------------------------- episodes
    def reset(self, start_vp, start_view, goal):
        """R2RBatch.reset + the rollout's initial bookkeeping (agent_dg.py:692-710): place B agents. Inputs: int arrays
        (host, pinned or not) — the ONLY per-rollout host->device traffic of the environment (12 B per episode)."""
        B = len(start_vp)
        dev = self.device
        if self.B != B:
            self.B = B
            self.vp = torch.empty(B, dtype=torch.int32, device=dev)
            self.view = torch.empty(B, dtype=torch.int32, device=dev)
            self.goal = torch.empty(B, dtype=torch.int32, device=dev)
            self.ended = torch.empty(B, dtype=torch.uint8, device=dev)
            self.last_dist = torch.empty(B, dtype=torch.float32, device=dev)
        self.vp.copy_(torch.as_tensor(start_vp, dtype=torch.int32), non_blocking=True)
        self.view.copy_(torch.as_tensor(start_view, dtype=torch.int32), non_blocking=True)
        self.goal.copy_(torch.as_tensor(goal, dtype=torch.int32), non_blocking=True)
        self.ended.zero_()
        self.err.zero_()
        self.visited = None                                     # --submit: allocated (zeroed) by the first visited_mask()
        # last_dist = the start's distance to the goal (agent_dg.py:693-695): one gather on the device
        torch.index_select(self.t_dist.view(-1), 0, self.vp.long() * self.n_vp + self.goal.long(), out=self.last_dist)
        return self

    def alloc(self, T):
        """Observation buffers for T steps ([T, B, ...]; every step keeps its own slice because backward re-reads it)."""
        cfg, B, dev = self.cfg, self.B, self.device
        f32 = dict(dtype=torch.float32, device=dev)
        return {"input_a_t": torch.empty(T, B, cfg.angle_size, **f32), "f_t": torch.empty(T, B, cfg.views, cfg.feat, **f32),
                "d_t": torch.empty(T, B, cfg.views, cfg.feat, **f32), "cand_feat": torch.empty(T, B, self.nc, cfg.feat, **f32),
                "cand_dfeat": torch.empty(T, B, self.nc, cfg.feat, **f32),
                "cand_leng": torch.empty(T, B, dtype=torch.int32, device=dev),
                "target": torch.empty(T, B, dtype=torch.int64, device=dev), "dist": torch.empty(T, B, **f32)}

    def observe(self, buf, t):
        """Write the observation of the current state into slot t of `buf` (from alloc())."""
        cfg = self.cfg
        p = ops._p
        ops.call("dasa_env_observe", p(self.rgb_bank), p(self.dep_bank), p(self.t_nbr), p(self.t_nbr_point), p(self.t_deg),
                 p(self.t_cand_angle), p(self.t_view_angle), p(self.t_agent_angle), p(self.t_dist), p(self.t_next_hop),
                 self.n_vp, self.dmax, p(self.vp), p(self.view), p(self.goal), p(self.ended), self.B, cfg.views, cfg.rgb_size,
                 cfg.angle_size, self.nc, cfg.headings, cfg.ignore_id, p(buf["f_t"][t]), p(buf["d_t"][t]), cfg.views * cfg.feat,
                 p(buf["cand_feat"][t]), p(buf["cand_dfeat"][t]), self.nc * cfg.feat, p(buf["input_a_t"][t]),
                 p(buf["cand_leng"][t]), p(buf["target"][t]), p(buf["dist"][t]), ops._stream())

    def step(self, action, reward=None, mask=None, traj_vp=None, traj_view=None):
        """Apply one action per episode (int64 candidate indices; END = deg, or ignore_id) and do the reward / mask / ended
        bookkeeping of agent_dg.py:890-935 on the device."""
        p = ops._p
        ops.call("dasa_env_step", p(action), self.cfg.ignore_id, p(self.t_nbr), p(self.t_nbr_point), p(self.t_deg), self.dmax,
                 p(self.t_dist), self.n_vp, p(self.vp), p(self.view), p(self.goal), p(self.ended), p(self.last_dist), p(reward),
                 p(mask), p(traj_vp), p(traj_view), p(self.err), self.B, ops._stream())

    def visited_mask(self, logit=None, want_mask=False):
        """--submit (agent_dg.py:834-840): the current viewpoints join the per-episode visited sets; candidates leading back to
        a visited viewpoint get logit = -inf (in place). Returns the uint8 [B, nc] mask when `want_mask`."""
        words = (self.n_vp + 31) // 32
        if self.visited is None:
            self.visited = torch.zeros(self.B, words, dtype=torch.int32, device=self.device)
        blocked = torch.empty(self.B, self.nc, dtype=torch.uint8, device=self.device) if want_mask else None
        p = ops._p
        assert logit is None or (logit.dtype == torch.float32 and logit.stride(1) == 1 and logit.shape[1] >= self.nc)
        ops.call("dasa_env_visited_mask", p(self.vp), p(self.t_nbr), p(self.t_deg), self.dmax, self.n_vp, p(self.visited), words,
                 p(blocked), self.nc, p(logit), 0 if logit is None else logit.stride(0), self.B, ops._stream())
        return blocked

    def check(self):
        """Host-side check of the device error word (one sync; call it after a rollout, not inside)."""
        e = int(self.err.item())
        if e & 1:
            raise RuntimeError("an action outside the candidate list reached the environment")
        if e & 2:
            raise NameError("The action doesn't change the move")       # agent_dg.py:925

    # ----------------------------------------------------------------------------------------- rollout front ends
    def teacher_episodes(self, T, instr):
        """Teacher-forced trajectories (feedback='teacher': the agent takes the shortest-path action, agent_dg.py:868-869)
        unrolled on the device: T x (observe, step). The result quacks like rollout.DeviceEpisodes (resident [T, B, ...]
        tensors), so NavPolicy.teacher_rollout can batch AdaIN + encoder over all T actions."""
        buf = self.alloc(T)
        traj = torch.empty(T + 1, self.B, dtype=torch.int32, device=self.device)
        traj[0].copy_(self.vp)
        for t in range(T):
            self.observe(buf, t)
            self.step(buf["target"][t], traj_vp=traj[t + 1])
        return EnvEpisodes(self, buf, T, instr, traj=traj)

    def shortest_path_features(self, max_steps=None):
        """Speaker.from_shortest_path (speaker.py:163-198) on the device: follow the teacher from the current state until every
        episode has stopped. Returns ((img_feats [B, L, 36, F], can_feats [B, L, F]), length [B]) — the inputs of the speaker's
        encoder: the panorama of every visited viewpoint and the feature row of the candidate taken there (zeros for STOP).
        One host sync at the end (the padded length L = longest path)."""
        T = int(max_steps or self.cfg.max_action)
        buf = self.alloc(T)
        for t in range(T):
            self.observe(buf, t)
            self.step(buf["target"][t])
        tgt, leng = buf["target"], buf["cand_leng"].long()                       # [T, B]
        stop = (tgt == leng - 1) | (tgt == self.cfg.ignore_id)                     # speaker.py:184-186
        idx = tgt.clamp(min=0)[:, :, None, None].expand(T, self.B, 1, self.cfg.feat)
        can = torch.gather(buf["cand_feat"], 2, idx).squeeze(2)
        can = torch.where(stop[:, :, None], torch.zeros_like(can), can)
        ended_before = (stop.long().cumsum(0) - stop.long()) > 0                    # ended BEFORE step t
        length = (~ended_before).long().sum(0)                                      # length += (1 - ended), speaker.py:189
        L = int(length.max())
        return (buf["f_t"][:L].transpose(0, 1).contiguous(), can[:L].transpose(0, 1).contiguous()), length

    def live_episodes(self, T, instr):
        """Closed-loop episodes for sampled / greedy feedback: observation t is produced when the policy asks for it and
        the policy's own action drives the transition (EnvEpisodes.advance)."""
        return EnvEpisodes(self, self.alloc(T + 1), T + 1, instr, live=True,
                           traj=torch.empty(T + 2, self.B, dtype=torch.int32, device=self.device))


class EnvEpisodes:
    """Duck-types rollout.DeviceEpisodes on top of a DeviceEnv. instr = (seq [B,80] int64, mask [B,Lmax] bool, lengths [B]
    [, lengths as a host list]) as produced by the tokenizer side (host tensors are uploaded once per rollout)."""

    FIELDS = FIELDS
    resident = True

    def __init__(self, env, buf, T, instr, live=False, traj=None):
        self.env, self.buf, self.T, self.B, self.cfg, self.live = env, buf, T, env.B, env.cfg, live
        dev = env.device
        seq, mask, lengths = instr[:3]
        self.seq, self.seq_mask = seq.to(dev, non_blocking=True), mask.to(dev, non_blocking=True)
        self.seq_lengths = lengths.to(dev, non_blocking=True).to(torch.int32)
        # host copy of the lengths (lets the encoder drop padding rows without a sync); pass it as instr[3] when `lengths`
        # already lives on the device (e.g. inside a CUDA-graph capture, where .tolist() would be an illegal sync)
        self.seq_lengths_host = list(instr[3]) if len(instr) > 3 else [int(x) for x in lengths.tolist()]
        for k in FIELDS:
            setattr(self, k, buf[k])
        self.dist = buf["dist"] if not live else None
        self.traj = traj
        self.ended = env.ended
        self._observed = -1
        if live and traj is not None:
            traj[0].copy_(env.vp)

    def step(self, t):
        if self.live and t > self._observed:
            assert t == self._observed + 1, "live observations are produced in order"
            self.env.observe(self.buf, t)
            self._observed = t
        return tuple(self.buf[k][t] for k in FIELDS)

    def target_at(self, t):
        return self.buf["target"][t]

    def advance(self, t, action, reward=None, mask=None):
        """The policy's action of step t drives the environment (sampled / greedy feedback)."""
        self.env.step(action, reward, mask, traj_vp=None if self.traj is None else self.traj[t + 1])
