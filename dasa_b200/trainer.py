"""Seq2SeqAgent.train's inner iteration (agent_dg.py:1347-1405: zero_grad -> accumulate_gradient(feedback) -> optim_step) as
a product API, single GPU or data parallel (one process per GPU, episodes sharded by rank).

    tr = RolloutTrainer(policy, T=35, feedback="teacher", lr=1e-4)       # world / rank from torch.distributed when initialised
    loss = tr.step(episodes)                                             # one optimizer step; `loss` stays on the device

What it adds over calling NavPolicy.teacher_rollout / sample_rollout / optim_step by hand:

  * normalisation for a sharded batch. The reference is a single process: the ML loss is divided by the batch size
    (agent_dg.py:1024) and the A2C loss by the batch-GLOBAL number of live (episode, action) pairs `total`
    (agent_dg.py:988-994). Here every rank scales its ML loss by ml_weight / (B_local * world), and the A2C loss is computed
    un-normalised, its `total` is summed over the ranks (one 4-byte all-reduce before the backward pass) and the loss divided by
    that global count: the SUM of the per-rank gradients is then exactly the gradient of one process running all episodes.
  * overlapped gradient reduction. The weight-gradient GEMMs are deferred to the end of the backward pass and flushed group by
    group, largest gradient buffer first (decoder, 120 MB of the 190 MB); each group's flat buffer is all-reduced asynchronously
    (NCCL's own stream) as soon as its GEMMs are enqueued, under the remaining groups' GEMMs. The optimizer waits for all of them.
  * the whole step (rollouts, backward, all-reduces, clip + RMSprop) captured as ONE CUDA graph, at any world size; the dropout
    seed lives in device memory and is advanced by a kernel inside the graph, so every replay draws fresh masks.
"""
import torch

from . import functions as Fn
from . import modules as M


class RolloutTrainer:
    def __init__(self, policy, T, feedback="teacher", lr=1e-4, ml_weight=0.4, world=None, dropout_source=None,
                 use_lr_scheduler=None, overlap=True, gamma=0.9, ent_coef=0.01, normalize="total", passes=None):
        """feedback 'teacher': one teacher-forced rollout per step (BASELINE configs[1]); 'sample': accumulate_gradient('sample')
        = teacher-forced rollout + sampled A2C rollout (agent_dg.py:1352-1356). passes: ml_weight of every accumulate_gradient
        pass of one optimizer step (finetune: GT env + augmented env, train.py:226-243); default one pass with `ml_weight`.
        use_lr_scheduler None = the policy's setting (README: on). The LambdaLR multiplier lives on the device, so a captured
        graph follows the schedule."""
        import torch.distributed as dist
        self.pol, self.T, self.feedback, self.lr = policy, T, feedback, lr
        self.ml_weights = list(passes) if passes is not None else [ml_weight]
        if world is None:
            world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.world = world
        self.src = dropout_source if dropout_source is not None else M.dropout_source()
        self.use_lr_scheduler, self.overlap = use_lr_scheduler, overlap
        self.gamma, self.ent_coef, self.normalize = gamma, ent_coef, normalize
        self.graph, self.graph_loss, self.graph_launches = None, None, 0
        self._works = []
        self._early = set()             # names of the optimizer groups already flushed + reduced during the backward pass
        if policy._flat is None:
            policy.flatten_parameters()

    # ------------------------------------------------------------------------------------------------- data parallel
    def broadcast_parameters(self, src=0):
        """Replicate rank `src`'s parameters (and RMSprop state) on every rank."""
        if self.world <= 1:
            return
        import torch.distributed as dist
        for g in self.pol._flat:
            dist.broadcast(g["flat_p"], src)
            dist.broadcast(g["flat_sq"], src)

    def _global_total(self, total):
        if self.world <= 1:
            return total
        import torch.distributed as dist
        t = total.detach().clone()
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t

    def _reduce_group(self, g):
        if self.world <= 1:
            return
        import torch.distributed as dist
        if self.overlap:
            self._works.append(dist.all_reduce(g["flat_g"], op=dist.ReduceOp.SUM, async_op=True))
        else:
            dist.all_reduce(g["flat_g"], op=dist.ReduceOp.SUM)

    def _early_reduce(self):
        """Hook at the start of the encoder's bi-LSTM backward (batched teacher-forced rollout, last backward pass of the step):
        every decoder / critic node has run, so those groups' weight-gradient GEMMs are flushed and their all-reduce (120 of the
        232 MB) starts now, under the bi-LSTM's latency-bound recurrence instead of after it."""
        for g in sorted(self.pol._flat, key=lambda g: -g["flat_g"].numel()):
            if g["name"] in ("decoder", "critic"):
                Fn.flush_weight_grads(owner=g["flat_g"])
                self._reduce_group(g)
                self._early.add(g["name"])

    def _flush_and_reduce(self):
        """Deferred weight-gradient GEMMs group by group (largest flat buffer first), each group's all-reduce started as soon as
        its gradients are complete."""
        groups = sorted(self.pol._flat, key=lambda g: -g["flat_g"].numel())
        for g in groups:
            if g["name"] in self._early:
                if Fn.pending_weight_grads(g["flat_g"]):
                    raise RuntimeError("a %s gradient arrived after the group's early all-reduce had started" % g["name"])
                continue
            Fn.flush_weight_grads(owner=g["flat_g"])
            self._reduce_group(g)
        self._early = set()
        Fn.flush_weight_grads()                               # anything outside the flat buffers (none in practice)

    def _wait_reductions(self):
        for w in self._works:
            w.wait()
        self._works = []

    # ---------------------------------------------------------------------------------------------------- one iteration
    def accumulate(self, ep, actions_in=None):
        """zero_grad + every accumulate_gradient pass, backward included (weight-gradient GEMMs still queued when deferred).
        Returns the summed loss (device tensor [1], this rank's share). actions_in: injected actions of the sampled rollout
        ([T] list of [B] tensors; tests) instead of the device RNG's."""
        pol, T, world = self.pol, self.T, self.world
        pol.zero_grad()
        self.src.advance()
        losses = []
        with M.use_dropout_source(self.src):
            for wi, w in enumerate(self.ml_weights):
                loss, _, _ = pol.teacher_rollout(ep, T, w / world, tag_steps=False)
                if self.feedback == "sample":
                    # The IL rollout is back-propagated as soon as it ends: same accumulated gradients as summing the two
                    # losses first (agent_dg.py:1352-1356), half the activation footprint.
                    loss.backward()
                    losses.append(loss.detach())
                    if world > 1 and self.normalize == "total":
                        rl, out = pol.sample_rollout(ep, T, tag_steps=False, gamma=self.gamma, ent_coef=self.ent_coef,
                                                     normalize="none", actions_in=actions_in)
                        rl = rl / self._global_total(out["total"]).clamp_(min=1.0)
                    else:
                        rl, out = pol.sample_rollout(ep, T, tag_steps=False, gamma=self.gamma, ent_coef=self.ent_coef,
                                                     normalize=self.normalize, actions_in=actions_in)
                        if world > 1:
                            rl = rl / world if self.normalize == "batch" else rl
                    rl.backward()
                    losses.append(rl.detach())
                else:
                    last = wi == len(self.ml_weights) - 1
                    if world > 1 and self.overlap and last and getattr(pol, "schedule", None) == "batched":
                        Fn.pre_encoder_backward = self._early_reduce
                    try:
                        loss.backward()
                    finally:
                        Fn.pre_encoder_backward = None
                    losses.append(loss.detach())
        total = losses[0].reshape(1)
        for x in losses[1:]:
            total = total + x.reshape(1)
        return total

    def finish(self):
        """flush the deferred weight gradients, all-reduce, clip + RMSprop (+ LambdaLR multiplier)."""
        self._flush_and_reduce()
        self._wait_reductions()
        self.pol.optim_step(self.lr, use_lr_scheduler=self.use_lr_scheduler)

    def reduce_gradients(self):
        """Deferred weight-gradient GEMMs + the all-reduces, completed (what finish() does before the optimizer)."""
        self._flush_and_reduce()
        self._wait_reductions()

    def step_eager(self, ep):
        loss = self.accumulate(ep)
        self.finish()
        return loss

    def capture(self, ep):
        """Capture step_eager(ep) as one CUDA graph (ep's tensors are the graph's inputs: refill them in place between replays).
        Call after at least one eager step (allocator, shared-memory attributes, NCCL communicator warm)."""
        from . import lib
        Fn.invalidate_weight_caches()                        # cached transposes must be rebuilt INSIDE the graph every replay
        torch.cuda.synchronize()
        torch.cuda.empty_cache()                             # the eager warm-up's cached blocks cannot serve the graph's private pool
        g = torch.cuda.CUDAGraph()
        l0 = lib.launches
        with torch.cuda.graph(g):
            loss = self.step_eager(ep() if callable(ep) else ep)
        self.graph, self.graph_loss, self.graph_launches = g, loss, lib.launches - l0
        return g

    def release_graph(self):
        self.graph, self.graph_loss = None, None
        Fn.invalidate_weight_caches()

    def step(self, ep=None):
        """One optimizer step: replays the captured graph when there is one (ep ignored: the graph reads the captured tensors),
        else runs eagerly."""
        if self.graph is not None:
            self.graph.replay()
            self.pol.iteration += 1
            return self.graph_loss
        return self.step_eager(ep)
