"""Static shape/flag description of the agent_dg navigation policy (the fields of the reference's global
`args` that the hot-path modules read: SURVEY.md §5.6; param.py:18-215; README.md:82-96)."""
from dataclasses import dataclass, replace


@dataclass(frozen=True)
class PolicyConfig:
    # feature geometry (agent_dg.py:28-29, param.py:36,107)
    rgb_size: int = 2048          # args.feature_size  (ResNet-152 pool5)
    angle_size: int = 128         # args.angle_feat_size
    views: int = 36               # 3 elevations x 12 headings (env.py:81-82)
    headings: int = 12
    # decoder (agent_dg.py:193; model.py:425-443)
    hidden: int = 1024            # args.d_hidden_size
    action_emb: int = 64          # args.aemb
    shift_kernel: int = 5         # args.shift_kernel_size
    dropout: float = 0.5          # args.dropout
    featdropout: float = 0.4      # args.featdropout
    # encoder (agent_dg.py:161; r2rmodel.py:2204-2250; bert-base-uncased config)
    enc_hidden: int = 1024        # args.d_enc_hidden_size (bi-LSTM hidden per direction)
    enc_dropout: float = 0.4      # args.d_dropout_ratio
    bert_hidden: int = 768
    bert_heads: int = 12
    bert_inter: int = 3072
    bert_dropout: float = 0.1
    bert_eps: float = 1e-12
    vocab: int = 30522
    max_pos: int = 512
    type_vocab: int = 2
    la_layers: int = 9            # args.d_la_layers
    vl_layers: int = 3            # args.d_vl_layers
    max_input: int = 80           # args.maxInput
    update_add_layer: bool = False  # finetune config sets True
    # critic / RL (model.py:970-982; param.py:150)
    critic_dim: int = 1024
    gamma: float = 0.9
    ignore_id: int = -100
    max_action: int = 35

    @property
    def feat(self):               # FEATURE_ALL_SIZE (agent_dg.py:29)
        return self.rgb_size + self.angle_size

    @property
    def ctx_dim(self):
        return 2 * self.enc_hidden

    @property
    def elevations(self):
        return self.views // self.headings

    def small(self):
        """Shrunk geometry for fast CPU tests / golden fixtures (all structure kept)."""
        # bert_hidden stays 768: the reference hard-codes the bi-LSTM input width (r2rmodel.py:2231,2233)
        return replace(self, rgb_size=256, angle_size=128, hidden=128, action_emb=64, enc_hidden=128,
                       bert_hidden=768, bert_heads=12, bert_inter=256, vocab=1200, max_pos=128,
                       la_layers=2, vl_layers=2, critic_dim=128, max_input=24)


def reference_args():
    """The reference's global flag object (`from param import args`, param.py:18-215) when the host program has ALREADY imported
    `param` (i.e. we are running inside r2r_src), else None. Never imported from here: param.py parses sys.argv at import."""
    import sys
    mod = sys.modules.get("param")
    return getattr(mod, "args", None) if mod is not None else None


def flag(name, explicit, default):
    """Value of a flag the reference's modules read from the global `args` at construction time: the explicit
    constructor kwarg when given, else param.args.<name> when the reference's param module is loaded, else the README default."""
    if explicit is not None:
        return explicit
    a = reference_args()
    return getattr(a, name, default) if a is not None else default


FULL = PolicyConfig()
SMALL = FULL.small()
