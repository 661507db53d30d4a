"""CPU tier, build container only: the oracle restatement against the LIVE reference modules imported from
/root/reference through the shims (skipped where the reference tree is not mounted, e.g. the GPU box)."""
import contextlib
import io

import pytest
import torch

from dasa_b200 import synth
from dasa_b200.config import SMALL
from oracle import load_reference
from oracle import restated as R

pytestmark = pytest.mark.skipif(not load_reference.available(), reason="reference tree not mounted")


@pytest.fixture(scope="module")
def ref_mods():
    from oracle.make_golden import build_reference_modules
    with contextlib.redirect_stdout(io.StringIO()):
        ref = load_reference.load()
    st = synth.policy_state(SMALL, 2)
    mods = build_reference_modules(ref, SMALL, st)
    for m in mods:
        m.eval()
    return ref, st, mods


def test_state_dict_contract(ref_mods):
    """synth state dicts carry exactly the reference's keys and shapes."""
    ref, st, (enc, dec, cri, ada) = ref_mods
    for m, k in ((enc, "encoder"), (dec, "decoder"), (cri, "critic"), (ada, "adaIn")):
        want = {n: tuple(v.shape) for n, v in m.state_dict().items()}
        have = {n: tuple(v.shape) for n, v in st[k].items()}
        assert want == have
        assert list(want) == list(have)          # same order too


@pytest.mark.parametrize("seed", [0, 1])
def test_step_matches_reference(ref_mods, seed):
    ref, st, (enc, dec, cri, ada) = ref_mods
    ep = synth.Episodes(4, 2, SMALL, seed=20 + seed, stress=bool(seed))
    C = SMALL.rgb_size
    a_t, f_t, d_t, cand, cand_d, leng, tgt = ep.step(0)
    with torch.no_grad():
        df = f_t.clone(); df[..., :C] = ada(f_t[..., :C].clone(), d_t[..., :C].clone())
        cf = cand.clone(); cf[..., :C] = ada(cand[..., :C].clone(), cand_d[..., :C].clone())
        ctx, eh, ec, _, _ = enc(ep.seq, mask=ep.seq_mask, lengths=ep.seq_lengths, f_t_all=f_t.clone())
        h1, c1, lg, ht, _ = dec(a_t, df, cf, eh, eh, ec, ctx, ep.seq_mask)
        logit, h_t, (h1o, c_t), aux = R.policy_step(st, SMALL, ep.seq, ep.seq_mask, ep.seq_lengths, ep.step(0), None)
    fin = torch.isfinite(logit)
    torch.testing.assert_close(logit[fin], lg[fin], rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(h_t, h1, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(h1o, ht, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(c_t, c1, rtol=1e-4, atol=1e-5)


def test_decoder_eval_does_not_mutate_inputs(ref_mods):
    """eval(): reference leaves feature/cand_feat bit-unchanged even with already_dropfeat=False (Appendix B)."""
    ref, st, (enc, dec, cri, ada) = ref_mods
    ep = synth.Episodes(3, 1, SMALL, seed=9)
    a_t, f_t, d_t, cand, cand_d, leng, tgt = ep.step(0)
    f0, c0 = f_t.clone(), cand.clone()
    h = torch.zeros(3, SMALL.hidden)
    ctx = torch.randn(3, ep.seq_mask.shape[1], SMALL.ctx_dim)
    with torch.no_grad():
        dec(a_t, f_t, cand, h, h, h, ctx, ep.seq_mask)
    assert torch.equal(f_t, f0) and torch.equal(cand, c0)


def test_checkpoint_roundtrip_with_reference_agent(tmp_path):
    """Seq2SeqAgent.save / load (agent_dg.py:1466-1510), run UNMODIFIED on the reference modules, against NavPolicy.save / load:
    a reference-written snapshot loads into the drop-in modules (parameters and RMSprop state) and a snapshot written here
    loads through the reference's own load() with --loadOptim."""
    from dasa_b200.rollout import NavPolicy
    from oracle.make_golden import build_reference_modules
    with contextlib.redirect_stdout(io.StringIO()):
        ref = load_reference.load()
    st = synth.policy_state(SMALL, 4)
    enc, dec, cri, ada = build_reference_modules(ref, SMALL, st)
    agent = object.__new__(ref.agent_dg.Seq2SeqAgent)
    agent.encoder, agent.decoder, agent.critic, agent.adaIn = enc, dec, cri, ada
    mk = lambda m: torch.optim.RMSprop(m.parameters(), lr=1e-4)
    agent.encoder_optimizer, agent.decoder_optimizer, agent.critic_optimizer, agent.adaIn_optimizer = mk(enc), mk(dec), mk(cri), mk(ada)
    # give the optimizers some state: one step on synthetic gradients of the trainable tensors
    gen = torch.Generator().manual_seed(0)
    for m in (dec, cri, ada):
        for p in m.parameters():
            p.grad = torch.randn(p.shape, generator=gen) * 0.01
    for k, p in enc.named_parameters():
        if not k.startswith("bert."):
            p.grad = torch.randn(p.shape, generator=gen) * 0.01
    for o in (agent.encoder_optimizer, agent.decoder_optimizer, agent.critic_optimizer, agent.adaIn_optimizer):
        o.step()
    ref.args.adaIn_type, ref.args.loadOptim = "channel", True
    path = str(tmp_path / "snap" / "ref_agent")
    agent.save(7, path)                                                        # the reference's own save()

    pol = NavPolicy(SMALL, synth.policy_state(SMALL, 9), "cpu")                # different weights, CPU-resident modules
    pol.flatten_parameters()
    assert pol.load(path, load_optim=True) == 7
    for name, m_ref, m_new in (("encoder", enc, pol.encoder), ("decoder", dec, pol.decoder), ("critic", cri, pol.critic),
                               ("adaIn", ada, pol.adaIn)):
        a, b = m_ref.state_dict(), m_new.state_dict()
        assert list(a) == list(b)
        for k in a:
            assert torch.equal(a[k], b[k]), (name, k)
    sq_ref = agent.decoder_optimizer.state_dict()["state"]
    for i, p in enumerate(pol.decoder.parameters()):
        if i in sq_ref:
            assert torch.equal(pol._square_avg("decoder", p, i), sq_ref[i]["square_avg"])
    assert pol.iteration == 1

    # and back: a snapshot written here through the reference's load()
    with torch.no_grad():
        for p in pol.decoder.parameters():
            p.add_(0.5)
    path2 = str(tmp_path / "snap" / "new_agent")
    pol.save(11, path2)
    with contextlib.redirect_stdout(io.StringIO()):
        assert agent.load(path2) == 11                                         # the reference's own load(), --loadOptim
    for (k, a), b in zip(dec.state_dict().items(), pol.decoder.state_dict().values()):
        assert torch.equal(a, b), k
    sq_new = agent.decoder_optimizer.state_dict()["state"]
    for i, p in enumerate(pol.decoder.parameters()):
        if i in sq_new and pol._square_avg("decoder", p, i) is not None and p.requires_grad and i in sq_ref:
            assert torch.equal(sq_new[i]["square_avg"], pol._square_avg("decoder", p, i))


def test_consistent_drop_rollout_matches_reference(ref_mods):
    """Aug-rollout feature dropout (consistent_drop, after_adain, depth_drop): one [C] mask shared by batch, views and steps on
    the AdaIN'd candidates, the raw and the AdaIN'd views; decoder called with already_dropfeat=True (agent_dg.py:780-785)."""
    from oracle.make_golden import reference_rollout
    ref, st, mods = ref_mods
    ep = synth.Episodes(4, 3, SMALL, seed=33)
    gen = torch.Generator().manual_seed(2)
    noise = (torch.rand(SMALL.rgb_size, generator=gen) >= SMALL.featdropout).float() / (1 - SMALL.featdropout)
    with torch.no_grad():
        loss_ref, logits_ref, _ = reference_rollout(ref, mods, SMALL, ep, 3, noise=noise)
        loss, logits, _ = R.teacher_rollout(st, SMALL, ep, 3, noise=noise)
        loss_plain, _, _ = R.teacher_rollout(st, SMALL, ep, 3)
    assert abs(float(loss) - float(loss_ref)) <= 1e-5 * abs(float(loss_ref))
    assert abs(float(loss) - float(loss_plain)) > 1e-4 * abs(float(loss_ref))           # the mask does change the rollout
    for a, b in zip(logits, logits_ref):
        fin = torch.isfinite(b)
        assert torch.equal(fin, torch.isfinite(a))
        torch.testing.assert_close(a[fin], b[fin], rtol=1e-4, atol=1e-5)
