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
