"""TEST INFRASTRUCTURE ONLY (oracle shim): name needed by r2rpretrain_class.py:3, never executed."""
from torch import nn


class BertOnlyMLMHead(nn.Module):
    def __init__(self, config):
        super().__init__()
