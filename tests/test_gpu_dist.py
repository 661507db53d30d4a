"""GPU tier: data-parallel training through the product API (dasa_b200/trainer.py RolloutTrainer) with world_size 2 — the
all-reduced gradients of two ranks holding B episodes each equal the gradients of ONE process running all 2B episodes, for the
teacher-forced rollout (ML loss / global batch, agent_dg.py:1024) and for accumulate_gradient('sample') (A2C loss / batch-global
`total`, agent_dg.py:988-994). Backend: NCCL with one GPU per rank when the box has two GPUs, else gloo with both ranks on
cuda:0 (NCCL refuses two ranks on one device; gloo all-reduces CUDA tensors through the host). fp32 precision: the only
difference between the two runs is the summation order over episodes, held to 1e-4."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

from dasa_b200 import synth
from dasa_b200.config import SMALL

pytestmark = pytest.mark.gpu

B_RANK, T, SEED = 3, 3, 21


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _slice(ep, lo, hi):
    import copy
    sub = copy.copy(ep)
    sub.B = hi - lo
    for k in ("input_a_t", "f_t", "d_t", "cand_feat", "cand_dfeat", "cand_leng", "target", "dist"):
        setattr(sub, k, getattr(ep, k)[:, lo:hi].contiguous())
    sub.seq, sub.seq_mask, sub.seq_lengths = ep.seq[lo:hi], ep.seq_mask[lo:hi], ep.seq_lengths[lo:hi]
    return sub


def _actions(ep, T, seed):
    gen = torch.Generator().manual_seed(seed)
    acts = []
    for t in range(T):
        leng = ep.cand_leng[t].long()
        a = (torch.rand(ep.B, generator=gen) * (leng - 1).float()).long().clamp(max=leng - 2).clamp(min=0)
        stop = torch.rand(ep.B, generator=gen) < 0.3
        stop[0] = False
        acts.append(torch.where(stop, leng - 1, a))
    return acts


def _gradients(world, rank, feedback, dev):
    """accumulate + reduce on this rank's shard; returns {group: flat gradient} and the loss."""
    from dasa_b200.rollout import DeviceEpisodes, NavPolicy
    from dasa_b200.trainer import RolloutTrainer
    cfg = SMALL
    ep = synth.Episodes(B_RANK * 2, T + 1, cfg, seed=SEED)
    acts = _actions(ep, T, 5)
    lo, hi = (0, ep.B) if world == 1 else (rank * B_RANK, (rank + 1) * B_RANK)
    sub = DeviceEpisodes(_slice(ep, lo, hi), dev)
    pol = NavPolicy(cfg, synth.policy_state(cfg, 4), dev).eval()
    tr = RolloutTrainer(pol, T, feedback=feedback, world=world)
    loss = tr.accumulate(sub, actions_in=[a[lo:hi].to(dev) for a in acts])
    tr.reduce_gradients()
    torch.cuda.synchronize()
    return {g["name"]: g["flat_g"].detach().cpu().clone() for g in pol._flat}, loss.detach().cpu()


def _worker(rank, world, port, backend, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE=str(world), RANK=str(rank), LOCAL_RANK=str(rank))
    dev = "cuda:%d" % (rank if backend == "nccl" else 0)
    torch.cuda.set_device(dev)
    dist.init_process_group(backend, rank=rank, world_size=world)
    res = {}
    for feedback in ("teacher", "sample"):
        grads, loss = _gradients(world, rank, feedback, dev)
        losses = [None] * world
        dist.all_gather_object(losses, loss)
        res[feedback] = (grads, sum(losses))
    if rank == 0:
        torch.save(res, out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_rank_gradients_equal_single_process_global_batch(tmp_path):
    backend = "nccl" if torch.cuda.device_count() >= 2 else "gloo"
    out = str(tmp_path / "dp.pt")
    mp.start_processes(_worker, args=(2, _free_port(), backend, out), nprocs=2, join=True, start_method="spawn")
    got = torch.load(out)
    for feedback in ("teacher", "sample"):
        want, loss = _gradients(1, 0, feedback, "cuda:0")
        grads, loss2 = got[feedback]
        assert abs(float(loss2) - float(loss)) <= 1e-4 * abs(float(loss)), (feedback, float(loss2), float(loss))
        checked = 0
        for name, w in want.items():
            if float(w.abs().max()) == 0.0:
                assert float(grads[name].abs().max()) == 0.0, (feedback, name)
                continue
            e = float((grads[name].double() - w.double()).norm() / w.double().norm())
            assert e <= 1e-4, "%s feedback, %s gradients: 2-rank all-reduced vs single-process global batch differ by %.3e (%s)" % (
                feedback, name, e, backend)
            checked += 1
        assert checked >= (3 if feedback == "teacher" else 4), (feedback, checked)
