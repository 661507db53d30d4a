"""TEST INFRASTRUCTURE ONLY (oracle shim): names imported by r2rmodel.py:7, never executed."""
from torch import nn


class BertModel(nn.Module):
    pass


class OpenAIGPTModel(nn.Module):
    pass
