"""CPU tier: host-side logic that needs no kernel — the LR schedule, the stacked-weight cache, the flat parameter layout."""
import torch

from dasa_b200 import ops, synth
from dasa_b200.config import SMALL
from dasa_b200.rollout import NavPolicy


def test_lr_lambda_matches_torch_lambdalr():
    """agent_dg.py:219-227 with the README flags (--warm_steps 1000 --decay_start 4000 --decay_intervals 2000 --lr_decay 0.2)."""
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.RMSprop([p], lr=1e-4)

    def ref(iter_count, warm_steps=1000, decay_start=4000, decay_intervals=2000, lr_decay=0.2):   # the closure of agent_dg.py:219
        if warm_steps > 0 and iter_count < warm_steps:
            return (1.0 + iter_count) / warm_steps
        elif iter_count < decay_start:
            return 1.0
        return lr_decay ** ((iter_count - decay_start) // decay_intervals)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, ref)
    for it in (0, 1, 500, 999, 1000, 3999, 4000, 5999, 6000, 8000, 10000):
        assert NavPolicy.lr_lambda(it) == ref(it)
    assert abs(sched.get_last_lr()[0] - 1e-4 * NavPolicy.lr_lambda(0)) < 1e-12


def test_stacked_weights_cache_follows_parameter_updates():
    w1, w2 = torch.nn.Parameter(torch.randn(6, 4)), torch.nn.Parameter(torch.randn(3, 4))
    b2 = torch.nn.Parameter(torch.randn(3))
    W, b = ops.stacked_weights((w1, w2), 0, 8, (None, b2))
    assert W.shape == (16, 4) and torch.equal(W[:6], w1.detach()) and torch.equal(W[6:9], w2.detach()) and not W[9:].any()
    assert torch.equal(b[6:9], b2.detach()) and not b[:6].any() and not b[9:].any()
    assert ops.stacked_weights((w1, w2), 0, 8, (None, b2))[0] is W            # cached
    with torch.no_grad():
        w2.add_(1.0)                                                           # in-place update bumps the version
    W2, _ = ops.stacked_weights((w1, w2), 0, 8, (None, b2))
    assert W2 is not W and torch.equal(W2[6:9], w2.detach())
    ops.weights_epoch += 1                                                     # raw-pointer optimizer updates bump the epoch
    assert ops.stacked_weights((w1, w2), 0, 8, (None, b2))[0] is not W2
    Wc, none = ops.stacked_weights((torch.nn.Parameter(torch.randn(5, 3)), torch.nn.Parameter(torch.randn(5, 2))), 1)
    assert Wc.shape == (5, 5) and none is None


def test_flat_parameter_layout_is_aligned_and_aliasing():
    pol = NavPolicy(SMALL, synth.policy_state(SMALL, 0), "cpu")
    groups = pol.flatten_parameters()
    assert [g["name"] for g in groups] == ["encoder", "decoder", "critic", "adaIn"]
    for g in groups:
        base = g["flat_p"].data_ptr()
        for p in g["params"]:
            off = (p.data_ptr() - base) // 4
            assert off % 64 == 0                                               # 256-byte boundaries (TMA needs 16)
            assert p.grad is not None and (p.grad.data_ptr() - g["flat_g"].data_ptr()) // 4 == off
        # writes through the flat buffer are visible in the module parameters (one RMSprop launch per group)
        g["flat_p"].zero_()
        assert all(float(p.abs().max()) == 0.0 for p in g["params"])
    enc_trainable = {k for k, p in pol.encoder.named_parameters() if p.requires_grad}
    assert enc_trainable and not any(k.startswith("bert.") for k in enc_trainable)   # frozen BERT stack in the train config


def test_module_flags_default_from_param_args_when_loaded():
    """the reference's modules read flags from the global `param.args`; ours take them as kwargs that default to
    param.args.<flag> when r2r_src's `param` module is loaded (and to the README values otherwise)."""
    import sys
    import types
    from dasa_b200 import config, modules as M
    assert config.reference_args() is None or "param" in sys.modules
    saved = sys.modules.get("param")
    fake = types.ModuleType("param")
    fake.args = types.SimpleNamespace(angle_feat_size=128, featdropout=0.25, use_shift=True, shift_kernel_size=3, critic_dim=64,
                                      dropout=0.3, ab_type="a", a_type="sigmoid")
    sys.modules["param"] = fake
    try:
        dec = M.BAttnDecoderLSTM(64, 32, 0.3, feature_size=64 + 128)
        assert dec.featdropout == 0.25 and dec.feat_att_layer.kernel_size == 3 and dec.drop_env.p == 0.25
        cri = M.Critic()
        assert cri.dim == 64 and cri.p == 0.3
        assert M.BAttnDecoderLSTM(64, 32, 0.3, feature_size=64 + 128, shift_kernel_size=5).feat_att_layer.kernel_size == 5   # explicit wins
        fake.args.a_type = "gumbel_sigmoid"
        import pytest
        with pytest.raises(NotImplementedError):
            M.DGAdaChannel(64)
    finally:
        if saved is None:
            del sys.modules["param"]
        else:
            sys.modules["param"] = saved
    dec = M.BAttnDecoderLSTM(64, 32, 0.3, feature_size=64 + 128)
    assert dec.featdropout == 0.4 and dec.feat_att_layer.kernel_size == 5
