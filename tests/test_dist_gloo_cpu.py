"""CPU tier: the data-parallel host logic (dasa_b200/dist.py) with world_size 2 over gloo — episode sharding, the
per-rank loss factor and the SUM all-reduce reproduce the single-process global-batch gradients (oracle arithmetic)."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

from dasa_b200 import dist as ddist
from dasa_b200 import synth
from dasa_b200.config import SMALL


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _slice_episodes(ep, lo, hi):
    import copy
    sub = copy.copy(ep)
    sub.B = hi - lo
    for k in ("input_a_t", "f_t", "d_t", "cand_feat", "cand_dfeat", "cand_leng", "target"):
        setattr(sub, k, getattr(ep, k)[:, lo:hi])
    sub.seq, sub.seq_mask, sub.seq_lengths = ep.seq[lo:hi], ep.seq_mask[lo:hi], ep.seq_lengths[lo:hi]
    return sub


def _worker(rank, world, port, out):
    from oracle import restated as R
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE=str(world), RANK=str(rank), LOCAL_RANK=str(rank))
    torch.set_num_threads(2)
    w, r, _ = ddist.init("gloo")
    assert (w, r) == (world, rank)
    cfg = SMALL
    st = synth.policy_state(cfg, 3)
    train = {k: v.requires_grad_(True) for k, v in st["decoder"].items()}
    ep = synth.Episodes(4, 2, cfg, seed=9)
    lo, hi = ddist.shard(ep.B, rank, world)
    sub = _slice_episodes(ep, lo, hi)
    # each rank: ml_weight / global batch (teacher_rollout divides by the LOCAL batch, so pass ml * B_local / B_global)
    loss, _, _ = R.teacher_rollout(st, cfg, sub, 2, ml_weight=0.4 * sub.B / ep.B)
    loss.backward()
    names = [k for k, v in train.items() if v.grad is not None]
    flat = torch.cat([train[k].grad.reshape(-1) for k in names])
    n = ddist.allreduce_sum_([flat], world)
    assert n == 1
    assert abs(ddist.max_over_ranks(rank, "cpu", world) - (world - 1)) < 1e-6
    if rank == 0:
        torch.save({"flat": flat, "names": names}, out)
    torch.distributed.destroy_process_group()


def test_shard_partition():
    for n, w in ((20, 8), (7, 2), (512, 8), (3, 4)):
        spans = [ddist.shard(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1
    assert ddist.loss_scale(0.4, 40) == pytest.approx(0.01)


@pytest.mark.timeout(300)
def test_two_rank_gradients_equal_global_batch(tmp_path):
    from oracle import restated as R
    out = str(tmp_path / "g.pt")
    mp.start_processes(_worker, args=(2, _free_port(), out), nprocs=2, join=True, start_method="spawn")
    got = torch.load(out)
    cfg = SMALL
    st = synth.policy_state(cfg, 3)
    train = {k: v.requires_grad_(True) for k, v in st["decoder"].items()}
    ep = synth.Episodes(4, 2, cfg, seed=9)
    loss, _, _ = R.teacher_rollout(st, cfg, ep, 2, ml_weight=0.4)
    loss.backward()
    want = torch.cat([train[k].grad.reshape(-1) for k in got["names"]])
    torch.testing.assert_close(got["flat"], want, rtol=1e-4, atol=1e-6)


# ---------------------------------------------------------------------------------------------------------------------------
# The PRODUCT data-parallel loop (dasa_b200/trainer.py RolloutTrainer: per-rank ML scaling, batch-global A2C `total`,
# per-group all-reduce, optimizer hand-off) over gloo with a stand-in policy: the rollouts are small differentiable torch
# functions of the flat parameter buffers, so the test isolates exactly the host logic that has no GPU dependency. The same loop
# with the real NavPolicy on two GPUs is tests/test_gpu_dist.py.
class _StubPolicy:
    """Quacks like rollout.NavPolicy for RolloutTrainer: two optimizer groups with flat parameter / gradient buffers."""

    def __init__(self):
        g = torch.Generator().manual_seed(5)
        self._flat = []
        for name, n in (("decoder", 7), ("critic", 3)):
            p = torch.randn(n, generator=g, dtype=torch.float64).requires_grad_(True)
            p.grad = torch.zeros(n, dtype=torch.float64)
            self._flat.append({"name": name, "params": [p], "clip": None, "flat_p": p.data, "flat_g": p.grad,
                               "flat_sq": torch.zeros(n, dtype=torch.float64)})
        self.iteration, self.steps_taken = 0, 0

    def zero_grad(self):
        for grp in self._flat:
            grp["flat_g"].zero_()

    def _features(self, ep):
        return ep["x"] @ self._flat[0]["params"][0], ep["x"][:, :3] @ self._flat[1]["params"][0]

    def teacher_rollout(self, ep, T, ml_weight, tag_steps=False):
        # sum over the episodes * ml_weight / B_local, the shape of agent_dg.py:850, 1024
        a, c = self._features(ep)
        return ((a ** 2).sum() + c.sum()) * ml_weight / ep["x"].shape[0], None, None

    def sample_rollout(self, ep, T, tag_steps=False, gamma=0.9, ent_coef=0.01, normalize="total", actions_in=None):
        a, c = self._features(ep)
        live = ep["live"]                                      # 0/1 per episode: the A2C mask; `total` = number of live pairs
        rl = (live * (a * c)).sum()
        total = live.sum().reshape(1)
        if normalize == "total":
            rl = rl / total.clamp(min=1.0)
        elif normalize == "batch":
            rl = rl / ep["x"].shape[0]
        return rl, {"total": total}

    def optim_step(self, lr, use_lr_scheduler=None):
        self.steps_taken += 1


def _stub_episodes(lo, hi):
    g = torch.Generator().manual_seed(11)
    x = torch.randn(6, 7, generator=g, dtype=torch.float64)
    live = torch.tensor([1., 0., 0., 1., 1., 1.], dtype=torch.float64)       # 1 vs 3 live pairs on the two shards
    return {"x": x[lo:hi], "live": live[lo:hi]}


def _trainer_worker(rank, world, port, out, feedback):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE=str(world), RANK=str(rank), LOCAL_RANK=str(rank))
    torch.set_num_threads(1)
    from dasa_b200 import modules as M
    from dasa_b200.trainer import RolloutTrainer
    ddist.init("gloo")
    pol = _StubPolicy()
    lo, hi = ddist.shard(6, rank, world)
    tr = RolloutTrainer(pol, T=2, feedback=feedback, world=world, dropout_source=M.DropoutSource(seed=1, device="cpu"))
    tr.broadcast_parameters()
    tr.step(_stub_episodes(lo, hi))
    assert pol.steps_taken == 1
    if rank == 0:
        torch.save([g["flat_g"].clone() for g in pol._flat], out)
    torch.distributed.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("feedback", ["teacher", "sample"])
def test_rollout_trainer_two_ranks_equal_one_process(tmp_path, feedback):
    """RolloutTrainer on 2 gloo ranks (episodes 0-2 / 3-5 with 1 vs 3 live pairs) leaves in every flat gradient buffer
    exactly the gradient of ONE process running all 6 episodes: ML loss / global batch (agent_dg.py:1024), A2C loss / the
    batch-global `total` (agent_dg.py:988-994) - the normaliser the round-1 loop got wrong."""
    from dasa_b200 import modules as M
    from dasa_b200.trainer import RolloutTrainer
    out = str(tmp_path / "g.pt")
    mp.start_processes(_trainer_worker, args=(2, _free_port(), out, feedback), nprocs=2, join=True, start_method="spawn")
    got = torch.load(out)
    pol = _StubPolicy()
    tr = RolloutTrainer(pol, T=2, feedback=feedback, world=1, dropout_source=M.DropoutSource(seed=1, device="cpu"))
    tr.step(_stub_episodes(0, 6))
    for g, w in zip(got, pol._flat):
        torch.testing.assert_close(g, w["flat_g"], rtol=1e-12, atol=1e-14)
        assert float(w["flat_g"].abs().max()) > 0
