"""TEST INFRASTRUCTURE ONLY (oracle shim). The reference depends on the un-vendored, un-pinned
`pytorch_transformers` package purely for two base classes (vilmodel.py:20, r2rmodel.py:9,
r2rpretrain_class.py:2-3); all transformer arithmetic is defined locally in vilmodel.py. This shim
supplies those names with bert-base-uncased hyper-parameters (the values
BertConfig.from_pretrained('bert-base-uncased') would have fetched from the network, r2rmodel.py:2229)."""
import torch
from torch import nn


class BertConfig:
    def __init__(self, **kw):
        self.vocab_size = 30522
        self.hidden_size = 768
        self.num_hidden_layers = 12
        self.num_attention_heads = 12
        self.intermediate_size = 3072
        self.hidden_act = "gelu"
        self.hidden_dropout_prob = 0.1
        self.attention_probs_dropout_prob = 0.1
        self.max_position_embeddings = 512
        self.type_vocab_size = 2
        self.initializer_range = 0.02
        self.layer_norm_eps = 1e-12
        self.output_attentions = False
        self.output_hidden_states = False
        self.torchscript = False
        self.pruned_heads = {}
        for k, v in kw.items():
            setattr(self, k, v)

    # tests may shrink the transformer by setting BertConfig.OVERRIDES before constructing a DicEncoder
    OVERRIDES = {}

    @classmethod
    def from_pretrained(cls, name, **kw):
        cfg = cls(**kw)
        if "large" in str(name):
            cfg.hidden_size, cfg.num_hidden_layers = 1024, 24
            cfg.num_attention_heads, cfg.intermediate_size = 16, 4096
        for k, v in cls.OVERRIDES.items():
            setattr(cfg, k, v)
        return cfg


class BertPreTrainedModel(nn.Module):
    config_class = BertConfig
    base_model_prefix = "bert"

    def __init__(self, config, *a, **k):
        super().__init__()
        self.config = config

    def _init_weights(self, module):
        if isinstance(module, (nn.Linear, nn.Embedding)):
            module.weight.data.normal_(mean=0.0, std=self.config.initializer_range)
        elif isinstance(module, nn.LayerNorm):
            module.bias.data.zero_()
            module.weight.data.fill_(1.0)
        if isinstance(module, nn.Linear) and module.bias is not None:
            module.bias.data.zero_()

    def init_weights(self):
        self.apply(self._init_weights)

    @classmethod
    def from_pretrained(cls, *a, **k):
        raise RuntimeError("no pretrained checkpoints offline")


class BertTokenizer:  # pragma: no cover
    @classmethod
    def from_pretrained(cls, *a, **k):
        raise RuntimeError("no tokenizer vocabulary offline")


class BertForMaskedLM(nn.Module):  # pragma: no cover
    pass
