"""The agent_dg rollout loop around the drop-in modules (agent_dg.py:633-1033, the teacher-forced / greedy branches),
with a device-resident synthetic stand-in for the environment (there is no simulator offline) and the optimizer step
of Seq2SeqAgent.optim_step (agent_dg.py:1389-1405) on fused kernels.

Differences from the reference loop that do not change the numbers:
  * the 4 clone()s + 2 strided copy-backs per step around adaIn (agent_dg.py:764-768) disappear: the gate GEMM reads the
    stride-2176 slices in place and its epilogue writes the AdaIN'd copy directly;
  * the decoder's per-element drop_env masks (model.py:506-508, 556-557) are folded into that epilogue, so the decoder is
    called with already_dropfeat=True on tensors that are already dropped (same values);
  * no per-step host synchronisation: losses/actions stay on the device (the reference syncs at agent_dg.py:890).
"""
import torch

from . import functions as Fn
from . import modules as M
from . import ops
from .config import FULL, PolicyConfig


class DeviceEpisodes:
    """synth.Episodes moved to the GPU (optionally re-uploaded from pinned host memory each step for the e2e metric)."""

    FIELDS = ("input_a_t", "f_t", "d_t", "cand_feat", "cand_dfeat", "cand_leng", "target")

    def __init__(self, ep, device="cuda", resident=True):
        self.B, self.T, self.cfg = ep.B, ep.T, ep.cfg
        self.host = ep
        self.device = device
        self.seq = ep.seq.to(device)
        self.seq_mask = ep.seq_mask.to(device)
        self.seq_lengths = ep.seq_lengths.to(device).to(torch.int32)
        self.seq_lengths_host = [int(x) for x in ep.seq_lengths.tolist()]   # lets the encoder drop padding rows without a sync
        self.dist = ep.dist.to(device) if hasattr(ep, "dist") else None      # [T+1, B] goal distances (sampled feedback)
        self.resident = resident
        if resident:
            for k in self.FIELDS:
                setattr(self, k, getattr(ep, k).to(device))
        else:
            # one device buffer per (field, step): the tensors are saved for backward, so they cannot be recycled inside a
            # rollout (and must not share an autograd version counter)
            self.stage = {k: [torch.empty_like(getattr(ep, k)[0], device=device) for _ in range(ep.T)] for k in self.FIELDS}

    def step(self, t):
        if self.resident:
            return tuple(getattr(self, k)[t] for k in self.FIELDS)
        out = []
        for k in self.FIELDS:                       # host -> device copy of this step's inputs (pinned source)
            dst = self.stage[k][t]
            dst.copy_(getattr(self.host, k)[t], non_blocking=True)
            out.append(dst)
        return tuple(out)

    def target_at(self, t):
        """Teacher action of step t (device tensor; valid after step(t) in the non-resident mode)."""
        return getattr(self, "target")[t] if self.resident else self.stage["target"][t]


class NavPolicy:
    """encoder + decoder + critic + adaIn, built like Seq2SeqAgent.__init__ (agent_dg.py:149-201)."""

    def __init__(self, cfg: PolicyConfig = FULL, state=None, device="cuda"):
        self.cfg, self.device = cfg, device
        self.encoder, self.decoder, self.critic, self.adaIn = M.build_policy(cfg, state, device)
        self.models = (self.encoder, self.decoder, self.critic, self.adaIn)
        self._opt = None
        self._flat = None
        self.schedule = "batched"      # teacher-forced rollouts: see teacher_rollout()
        self.batch_language = True     # evaluate the 9 language layers for all T actions of a rollout in one batched pass

    def train(self):
        for m in self.models:
            m.train()
        return self

    def eval(self):
        for m in self.models:
            m.eval()
        return self

    def zero_grad(self):
        if getattr(self, "_flat", None) is not None:
            for g in self._flat:
                g["flat_g"].zero_()
            return
        for m in self.models:
            for p in m.parameters():
                if p.grad is not None:
                    p.grad.zero_()

    # ------------------------------------------------------------------------------------------------ one nav step
    def make_noise(self):
        """noise = decoder.drop_env(ones(feature_size)) of the augmented (speaker) rollouts (agent_dg.py:656, 677): a uint8
        keep vector [C]; its scale is 1 / (1 - featdropout). None in eval mode (drop_env is the identity there)."""
        if not self.decoder.training:
            return None
        m, _ = M.dropout_source().mask("env.noise", (self.cfg.rgb_size,), self.cfg.featdropout, True, self.device)
        return m

    def _consistent(self, noise, f_t, n_cand_rows):
        """consistent_drop with --env_drop_stage after_adain --depth_drop (agent_dg.py:780-785): the [C] mask, shared by the
        batch, the views and all steps, replaces the decoder's per-element drop_env masks on the AdaIN'd views / candidates
        (folded into the gate GEMM's epilogue as row-broadcast masks) and also multiplies the RAW views the encoder reads."""
        cfg = self.cfg
        C, A = cfg.rgb_size, cfg.angle_size
        scale = 1.0 / (1.0 - cfg.featdropout)
        noise = ops.as_keep_mask(noise)
        rows_f = f_t.shape[0] * f_t.shape[1]
        m_f = noise.view(1, C).expand(rows_f, C).contiguous()
        m_c = noise.view(1, C).expand(n_cand_rows, C).contiguous()
        f_enc = torch.empty(f_t.shape, device=f_t.device, dtype=torch.float32)
        ops.dropout_apply(f_t[..., :C], m_f, scale, out=f_enc[..., :C])
        ops.axpy2d(1.0, f_t[..., C:], f_enc[..., C:], accumulate=False)
        return m_f, m_c, scale, f_enc

    def step(self, ep, t, carry, lang_out=None, want_ctx=False, noise=None):
        """Loop body of vl_rollout up to the masked logits (agent_dg.py:727-841). carry = None at t == 0. `noise`: the [C]
        keep vector of an augmented rollout (see make_noise / _consistent)."""
        cfg, tr = self.cfg, self.decoder.training
        a_t, f_t, d_t, cand, cand_d, leng, _ = ep.step(t)
        src = M.dropout_source()
        B, V, _ = f_t.shape
        C = cfg.rgb_size
        if noise is not None:
            m_f, m_c, s_f, f_t_enc = self._consistent(noise, f_t, B * cand.shape[1])
            s_c = s_f
        else:
            m_f, s_f = src.mask("dec.feat", (B, V, C), cfg.featdropout, tr, f_t.device)
            m_c, s_c = src.mask("dec.cand", (B, cand.shape[1], C), cfg.featdropout, tr, f_t.device)
            f_t_enc = f_t
        df_t = self.adaIn.gate_features(f_t, d_t, m_f, s_f)                 # K1 views
        cand_g = self.adaIn.gate_features(cand, cand_d, m_c, s_c)           # K1 candidates
        f_t = f_t_enc                                                       # what the encoder sees
        ctx, en_h, en_c, _, _ = self.encoder(ep.seq, ep.seq_mask, ep.seq_lengths, f_t_all=f_t, lang_out=lang_out,
                                             lengths_host=ep.seq_lengths_host)                                  # RAW f_t
        prev_h1, c_0 = (en_h, en_c) if carry is None else carry
        h_t, c_t, logit, h1, _ = self.decoder(a_t, df_t, cand_g, prev_h1, prev_h1, c_0, ctx, ep.seq_mask,
                                              already_dropfeat=True, cand_leng=leng)
        if want_ctx:
            return logit, h_t, (h1, c_t), ctx
        return logit, h_t, (h1, c_t)

    # ------------------------------------------------------------------------------------------- teacher-forced rollout
    def teacher_rollout(self, ep, T=None, ml_weight=0.4, tag_steps=True, schedule=None, noise=None):
        """feedback='teacher', train_rl=False (agent_dg.py:1368-1370): returns (loss tensor [1], logits list, actions list).
        loss = sum_t CE_sum(logit_t, target_t) * ml_weight / B  (agent_dg.py:850, 1024).

        schedule 'sequential': every action runs AdaIN -> encoder -> decoder in turn, as the reference loop does (the only
          order possible when the next observation depends on the sampled / greedy action).
        schedule 'batched' (default for teacher forcing): the agent follows the teacher, so the trajectory and all T
          observations are known up front; AdaIN gates, the cross-modal layers and the bi-LSTM of all T actions run as ONE
          batch of T*B sequences (identical per-(action, episode) arithmetic and dropout masks), only the recurrent decoder
          stays sequential."""
        T = ep.T if T is None else T
        schedule = schedule or self.schedule
        src = M.dropout_source()
        base_prefix = src.prefix
        cfg, tr = self.cfg, self.decoder.training
        carry, total, logits, actions = None, None, [], []
        if schedule == "batched" and ep.resident:
            B, C = ep.B, cfg.rgb_size
            f_all = ep.f_t[:T].reshape(T * B, cfg.views, cfg.feat)
            d_all = ep.d_t[:T].reshape(T * B, cfg.views, cfg.feat)
            nc = ep.cand_feat.shape[2]
            cand_all = ep.cand_feat[:T].reshape(T * B, nc, cfg.feat)
            candd_all = ep.cand_dfeat[:T].reshape(T * B, nc, cfg.feat)
            dev = f_all.device
            f_enc = f_all
            if noise is not None:                   # augmented rollout: one [C] mask for every view / candidate row and step
                m_f, m_c, s_f, f_enc = self._consistent(noise, f_all, T * B * nc)
                s_c = s_f
            else:
                m_f, s_f = src.mask_steps("dec.feat", (B, cfg.views, C), cfg.featdropout, tr, dev, T)
                m_c, s_c = src.mask_steps("dec.cand", (B, nc, C), cfg.featdropout, tr, dev, T)
            df_all = self.adaIn.gate_features(f_all, d_all, m_f, s_f)                 # K1, all actions
            candg_all = self.adaIn.gate_features(cand_all, candd_all, m_c, s_c)
            ctx_all, en_h, en_c = self.encoder.encode_rollout(ep.seq, ep.seq_mask, ep.seq_lengths, f_enc, T,
                                                              lengths_host=ep.seq_lengths_host)
            L = ctx_all.shape[1]
            ctx_steps = ctx_all.view(T, B, L, -1).unbind(0)          # unbind: one stacked gradient instead of T zero-filled ones
            df_steps = df_all.view(T, B, cfg.views, cfg.feat).unbind(0)
            cand_steps = candg_all.view(T, B, nc, cfg.feat).unbind(0)
            # Off the recurrent path, hence batched over the T actions: the action embeddings (before the loop) and the
            # candidate logits + cross entropy (after it). The loop keeps only what h_tilde_t -> h_tilde_{t+1} needs.
            mask_u8 = ep.seq_mask.to(torch.uint8)            # converted once (the attention kernel reads uint8 pad flags)
            emb_all = self.decoder.embed_actions(ep.input_a_t[:T].reshape(T * B, -1), T)
            # the recurrent part of all T actions: ONE cooperative launch (csrc/decoder_persist.cu) where the kernel applies ...
            h_tilde_all = self.decoder.rollout_steps(emb_all, df_all, ctx_all, mask_u8, en_h, en_c, T)
            if h_tilde_all is None:                          # ... else one decoder call per action (exact-fp32 mode, B > 32)
                emb_steps = emb_all.view(T, B, -1).unbind(0)
                h_tildes = []
                for t in range(T):
                    if tag_steps:
                        src.prefix = base_prefix + "t%d." % t
                    prev_h1, c_0 = (en_h, en_c) if carry is None else carry
                    h_t, c_t, _, h1, _ = self.decoder(None, df_steps[t], None, prev_h1, prev_h1, c_0, ctx_steps[t], mask_u8,
                                                      already_dropfeat=True, emb=emb_steps[t], want_logit=False)
                    carry = (h1, c_t)
                    h_tildes.append(h1)
                src.prefix = base_prefix
                h_tilde_all = torch.cat(h_tildes, 0)
            logit_all = self.decoder.candidate_logits_steps(h_tilde_all, candg_all, ep.cand_leng[:T].reshape(T * B), T)
            total, a_all = Fn.MaskedCEFn.apply(logit_all, ep.target[:T].reshape(T * B), cfg.ignore_id)
            logits = list(logit_all.view(T, B, nc).unbind(0))
            actions = list(a_all.view(T, B).unbind(0))
            return total * (ml_weight / ep.B), logits, actions
        # the instruction-only language stack of all T actions in one batched pass (per-action dropout masks preserved)
        lang_all = self.encoder.language_for_rollout(ep.seq, ep.seq_mask, T, ep.seq_lengths_host) if self.batch_language else None
        for t in range(T):
            if tag_steps:
                src.prefix = base_prefix + "t%d." % t
            logit, h_t, carry = self.step(ep, t, carry, None if lang_all is None else lang_all[t], noise=noise)
            loss_t, a_t = Fn.MaskedCEFn.apply(logit, ep.target_at(t), self.cfg.ignore_id)
            total = loss_t if total is None else total + loss_t
            logits.append(logit)
            actions.append(a_t)
        src.prefix = base_prefix
        return total * (ml_weight / ep.B), logits, actions

    # ------------------------------------------------------------------------------- sampled feedback + A2C (a10, a11)
    def sample_rollout(self, ep, T=None, actions_in=None, gamma=0.9, ent_coef=0.01, normalize="total", tag_steps=True):
        """feedback='sample', train_rl=True, train_ml=None (agent_dg.py:725-999) over a pre-generated observation stream
        holding T+1 observations and goal distances `ep.dist` [T+1,B] (there is no simulator offline). Per action:
        AdaIN -> encoder -> decoder -> Categorical sample (device RNG, or `actions_in[t]` injected) with log-prob and entropy
        -> reward / mask / ended kernel; then the extra decoder + critic call on the next observation, the critic over all T
        hidden states as ONE batch, and the fused A2C epilogue. Nothing is read back to the host inside the rollout.
        Returns (loss [1], dict)."""
        T = (ep.T - 1) if T is None else T
        live = getattr(ep, "live", False)       # env.EnvEpisodes: the sampled action drives a device-resident environment
        assert ep.resident and (live or ep.dist is not None) and ep.T >= T + 1, \
            "sample_rollout needs T+1 resident observations + dist (or live environment episodes)"
        cfg, B, dev = self.cfg, ep.B, ep.f_t.device
        src = M.dropout_source()
        base_prefix = src.prefix
        ended = ep.ended if live else torch.zeros(B, dtype=torch.uint8, device=dev)
        reward = torch.empty(T, B, device=dev)
        mask = torch.empty(T, B, device=dev)
        carry, ctx, hidden, logps, ents, actions, logits = None, None, [], [], [], [], []
        lang_all = self.encoder.language_for_rollout(ep.seq, ep.seq_mask, T, ep.seq_lengths_host) if self.batch_language else None
        for t in range(T):
            if tag_steps:
                src.prefix = base_prefix + "t%d." % t
            logit, h_t, carry, ctx = self.step(ep, t, carry, None if lang_all is None else lang_all[t], want_ctx=True)
            hidden.append(h_t)
            logits.append(logit)
            u = torch.rand(B, device=dev) if actions_in is None else None
            a_t, lp, en = Fn.PolicySampleFn.apply(logit, u, None if actions_in is None else actions_in[t])
            if live:    # make_equiv_action + next state + reward / mask / ended on the device (agent_dg.py:890-935)
                ep.advance(t, a_t, reward[t], mask[t])
            else:
                ops.nav_reward(a_t, ep.cand_leng[t], cfg.ignore_id, ep.dist[t + 1], ep.dist[t], ended, reward[t], mask[t])
            logps.append(lp)
            ents.append(en)
            actions.append(a_t)
        # last action in A2C (agent_dg.py:945-957): the decoder sees the RAW next observation and applies its own drop_env
        a_n, f_n, _, cand_n, _, _, _ = ep.step(T)
        src.prefix = base_prefix + "last."
        h1, c_t = carry
        last_h, _, _, _, _ = self.decoder(a_n, f_n.clone(), cand_n.clone(), hidden[-1], h1, c_t, ctx, ep.seq_mask,
                                          already_dropfeat=False)
        last_value = self.critic(last_h).detach().reshape(B)
        src.prefix = base_prefix
        values = self.critic.forward_steps(torch.cat(hidden, 0), T).view(T, B)       # a8: one batched call, per-step masks
        loss, total = Fn.A2CLossFn.apply(torch.stack(logps), torch.stack(ents), values, last_value, reward, mask, ended,
                                         gamma, ent_coef, normalize)
        return loss, {"logits": logits, "actions": actions, "logps": logps, "ents": ents, "values": values,
                      "last_value": last_value, "reward": reward, "mask": mask, "ended": ended, "total": total}

    def backward(self, loss):
        """loss.backward() + the deferred, batched weight-gradient GEMMs (functions.defer_weight_grads)."""
        loss.backward()
        Fn.flush_weight_grads()

    @torch.no_grad()
    def greedy_rollout(self, ep, T=None, submit=False):
        """feedback='argmax' decode (agent_dg.py:871-875) over pre-generated observations: per-step greedy actions.
        submit=True (args.submit, agent_dg.py:834-840; live environment episodes only): candidates that lead back to an already
        visited viewpoint are masked before the argmax."""
        T = ep.T if T is None else T
        carry, actions, logits = None, [], []
        if submit and not getattr(ep, "live", False):
            raise ValueError("submit=True needs live environment episodes (the visited sets follow the agent's own actions)")
        for t in range(T):
            logit, h_t, carry = self.step(ep, t, carry)
            if submit:
                ep.env.visited_mask(logit)
            _, a_t, _, _ = ops.masked_ce(logit, None, self.cfg.ignore_id, 0.0, None, want_grad=False)
            if getattr(ep, "live", False):
                ep.advance(t, a_t)
            actions.append(a_t)
            logits.append(logit)
        return actions, logits

    # ------------------------------------------------------------------------------------------------- optimizer (a12)
    def flatten_parameters(self):
        """Re-home the trainable parameters of each optimizer group (encoder / decoder / critic / adaIn, agent_dg.py:214-241)
        in one flat fp32 buffer, with a matching flat gradient buffer that every `p.grad` views: zero_grad is one memset per
        group, the data-parallel reduction one all-reduce per group, clip + RMSprop one launch per group. The frozen BERT stack
        (detached in the train config, vilmodel.py:1377-1410) is excluded and marked requires_grad=False; in the finetune
        config (update_add_layer) the cross-modal layers and the vision encoder stay trainable."""
        groups = []
        for name, m, clip in (("encoder", self.encoder, 40.0), ("decoder", self.decoder, 40.0), ("critic", self.critic, None),
                              ("adaIn", self.adaIn, None)):
            ps = []
            for k, p in m.named_parameters():
                finetuned = self.cfg.update_add_layer and (k.startswith("bert.addlayer.") or k.startswith("bert.vision_encoder."))
                if name == "encoder" and k.startswith("bert.") and not finetuned:
                    p.requires_grad_(False)
                    continue
                ps.append(p)
            # every parameter starts on a 256-byte boundary (a 5-element bias must not knock the weights behind it off the
            # 16-byte alignment TMA needs: misaligned weights silently fall back to the FFMA GEMM). The padding stays zero in
            # the parameter, gradient and RMSprop buffers, for which the update is exactly zero.
            ALIGN = 64
            offs, n = [], 0
            for p in ps:
                offs.append(n)
                n += (p.numel() + ALIGN - 1) // ALIGN * ALIGN
            flat_p = torch.zeros(n, device=self.device, dtype=torch.float32)
            flat_g = torch.zeros(n, device=self.device, dtype=torch.float32)
            for p, off in zip(ps, offs):
                k = p.numel()
                flat_p[off:off + k].copy_(p.data.reshape(-1))
                p.data = flat_p[off:off + k].view_as(p)
                p.grad = flat_g[off:off + k].view_as(p)
            groups.append({"name": name, "params": ps, "clip": clip, "flat_p": flat_p, "flat_g": flat_g,
                           "flat_sq": torch.zeros_like(flat_p)})
        self._flat = groups
        self._opt = {"sumsq": torch.zeros(1, device=self.device), "coef": torch.ones(1, device=self.device)}
        return groups

    # ------------------------------------------------------------------------------------- checkpoints (agent_dg.py:1466-1510)
    def _named_groups(self):
        names = ("encoder", "decoder", "critic", "adaIn")
        return [(n, m) for n, m in zip(names, self.models) if m is not None]

    def _square_avg(self, name, p, index):
        """RMSprop second-moment tensor of parameter `p` (flat layout: a view of the group's flat buffer; else the list)."""
        if getattr(self, "_flat", None) is not None:
            for g in self._flat:
                if g["name"] == name:
                    for q in g["params"]:
                        if q is p:
                            off = (p.data_ptr() - g["flat_p"].data_ptr()) // 4
                            return g["flat_sq"][off:off + p.numel()].view_as(p)
            return None
        if self._opt is not None and "groups" in self._opt:
            for g in self._opt["groups"]:
                if g["name"] == name:
                    for q, sq in zip(g["params"], g["sq"]):
                        if q is p:
                            return sq
        return None

    def save(self, epoch, path, lr=1e-4):
        """Seq2SeqAgent.save (agent_dg.py:1466-1487): {name: {'epoch', 'state_dict', 'optimizer'}} for encoder / decoder /
        critic / adaIn, with the optimizer entry in torch.optim.RMSprop's state_dict layout (param index = position in
        module.parameters(), 'square_avg' + 'step' per parameter that has been updated), so the reference's load() with
        --loadOptim accepts it."""
        import os
        the_dir, _ = os.path.split(path)
        if the_dir:
            os.makedirs(the_dir, exist_ok=True)
        states = {}
        for name, model in self._named_groups():
            params = list(model.parameters())
            state = {}
            for i, p in enumerate(params):
                sq = self._square_avg(name, p, i) if self.iteration > 0 else None
                if sq is not None and p.requires_grad:
                    state[i] = {"step": torch.tensor(float(self.iteration)), "square_avg": sq.detach().clone()}
            group = {"lr": self.group_lr(name, lr), "momentum": 0, "alpha": 0.99, "eps": 1e-08,
                     "centered": False, "weight_decay": 0, "capturable": False, "foreach": None, "maximize": False,
                     "differentiable": False, "params": list(range(len(params)))}
            states[name] = {"epoch": epoch + 1,
                            "state_dict": {k: v.detach().clone() for k, v in model.state_dict().items()},
                            "optimizer": {"state": state, "param_groups": [group]}}
        torch.save(states, path)

    def load(self, path, load_optim=False):
        """Seq2SeqAgent.load (agent_dg.py:1489-1510): parameters (and, with load_optim = --loadOptim, the RMSprop second
        moments and the step count) from a checkpoint written by the reference or by save(). Returns the epoch."""
        states = torch.load(path, map_location="cpu", weights_only=False)
        for name, model in self._named_groups():
            if name not in states:
                continue
            state = model.state_dict()
            if set(state.keys()) != set(states[name]["state_dict"].keys()):
                print("NOTICE: DIFFERENT KEYS IN THE LISTEREN")
            with torch.no_grad():       # copy in place: parameters may be views of the flat buffers
                for k, v in states[name]["state_dict"].items():
                    if k in state:
                        state[k].copy_(v)
            if load_optim:
                opt = states[name]["optimizer"]
                if getattr(self, "_flat", None) is None and (self._opt is None or "groups" not in self._opt):
                    self._build_optimizer(1e-4)
                for i, p in enumerate(model.parameters()):
                    ent = opt["state"].get(i)
                    sq = self._square_avg(name, p, i)
                    if ent is not None and sq is not None:
                        with torch.no_grad():
                            sq.copy_(ent["square_avg"])
                        self.iteration = max(self.iteration, int(float(ent["step"])))
                        if self._opt is not None:
                            self._opt.pop("iter_dev", None)       # the device-side schedule counter restarts from self.iteration
        Fn.invalidate_weight_caches()
        return states["encoder"]["epoch"] - 1

    def grad_buffers(self):
        return [g["flat_g"] for g in self._flat]

    def param_buffers(self):
        return [g["flat_p"] for g in self._flat]

    def _build_optimizer(self, lr):
        groups = []
        for name, m, clip in (("encoder", self.encoder, 40.0), ("decoder", self.decoder, 40.0), ("critic", self.critic, None),
                              ("adaIn", self.adaIn, None)):
            ps = [p for p in m.parameters() if p.requires_grad]
            groups.append({"name": name, "params": ps, "clip": clip, "sq": [torch.zeros_like(p) for p in ps]})
        self._opt = {"groups": groups, "lr": lr, "sumsq": torch.zeros(1, device=self.device),
                     "coef": torch.ones(1, device=self.device)}

    @staticmethod
    def lr_lambda(iter_count, warm_steps=1000, decay_start=4000, decay_intervals=2000, lr_decay=0.2):
        """The multiplier of agent_dg.py:219-227 (README flags --warm_steps 1000 --decay_start 4000 --decay_intervals 2000
        --lr_decay 0.2): linear warm-up, flat, then a 0.2x step every 2000 iterations."""
        if warm_steps > 0 and iter_count < warm_steps:
            return (1.0 + iter_count) / warm_steps
        if iter_count < decay_start:
            return 1.0
        return lr_decay ** ((iter_count - decay_start) // decay_intervals)

    def group_lr(self, name, lr, use_lr_scheduler=None):
        """LambdaLR on the decoder, critic and adaIn optimizers; the encoder optimizer has no scheduler (agent_dg.py:230-241).
        Optimizer step number i (0-based) runs with lr * lr_lambda(i), as torch's LambdaLR does. Without --use_lr_scheduler the
        reference STILL constructs the adaIn LambdaLR (agent_dg.py:238) and never steps it, which pins the adaIn rate at
        lr * lr_lambda(0) = lr / warm_steps: mirrored here."""
        use = self.use_lr_scheduler if use_lr_scheduler is None else use_lr_scheduler
        if name == "encoder":
            return lr
        if not use:
            return lr * self.lr_lambda(0, **self.lr_schedule) if name == "adaIn" else lr
        return lr * self.lr_lambda(self.iteration, **self.lr_schedule)

    iteration = 0
    lr_schedule = {}            # overrides of lr_lambda's README defaults (--warm_steps 1000 --decay_start 4000 ...)
    use_lr_scheduler = True     # README.md:82-96 trains with --use_lr_scheduler

    def _schedule_args(self):
        d = dict(warm_steps=1000, decay_start=4000, decay_intervals=2000, lr_decay=0.2)
        d.update(self.lr_schedule)
        return d

    def optim_step(self, lr=1e-4, use_lr_scheduler=None):
        """clip_grad_norm_(encoder, 40), clip_grad_norm_(decoder, 40), RMSprop on all four groups, then the three LambdaLR
        schedulers (agent_dg.py:1389-1405). Parameters that never received a gradient are skipped, like torch.optim does for
        grad=None (in the flat layout they carry an all-zero gradient, for which the RMSprop update is exactly zero).
        Flat layout: the base rate is a host scalar, the LambdaLR multiplier is computed ON THE DEVICE from a device-resident
        iteration counter (dasa_lr_lambda), so a CUDA graph that contains this call keeps following the schedule when replayed."""
        use = self.use_lr_scheduler if use_lr_scheduler is None else use_lr_scheduler
        if getattr(self, "_flat", None) is not None:
            o = self._opt
            if use:
                if "iter_dev" not in o:
                    o["iter_dev"] = torch.tensor([self.iteration], dtype=torch.int32, device=self.device)
                    o["mult_dev"] = torch.ones(1, dtype=torch.float32, device=self.device)
                sa = self._schedule_args()
                ops.lr_lambda(o["iter_dev"], o["mult_dev"], sa["warm_steps"], sa["decay_start"], sa["decay_intervals"], sa["lr_decay"])
            for g in self._flat:
                coef = None
                if g["clip"] is not None:
                    o["sumsq"].zero_()
                    ops.sumsq(g["flat_g"], o["sumsq"])
                    ops.clip_coef(o["sumsq"], g["clip"], o["coef"])
                    coef = o["coef"]
                scheduled = use and g["name"] != "encoder"
                ops.rmsprop_step(g["flat_p"], g["flat_g"], g["flat_sq"], lr if scheduled else self.group_lr(g["name"], lr, use), 0.99,
                                 1e-8, 0.0, coef, o["mult_dev"] if scheduled else None)
            self.iteration += 1
            Fn.invalidate_weight_caches()       # parameters changed behind autograd's back (raw-pointer update)
            return
        if self._opt is None:
            self._build_optimizer(lr)
        o = self._opt
        for g in o["groups"]:
            coef = None
            live = [(p, sq) for p, sq in zip(g["params"], g["sq"]) if p.grad is not None]
            if g["clip"] is not None:
                o["sumsq"].zero_()
                for p, _ in live:
                    ops.sumsq(p.grad, o["sumsq"])
                ops.clip_coef(o["sumsq"], g["clip"], o["coef"])
                coef = o["coef"]
            glr = self.group_lr(g["name"], lr, use)
            for p, sq in live:
                ops.rmsprop_step(p.data, p.grad, sq, glr, 0.99, 1e-8, 0.0, coef)
        self.iteration += 1
        Fn.invalidate_weight_caches()
