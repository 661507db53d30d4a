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
