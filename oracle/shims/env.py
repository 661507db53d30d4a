"""TEST INFRASTRUCTURE ONLY (oracle shim). Shadows /root/reference/r2r_src/env.py, which loads the
Matterport depth-feature index at import time (env.py:31) and therefore cannot be imported offline.
agent_dg.py only needs the name R2RBatch to exist (agent_dg.py:15)."""


class R2RBatch:  # pragma: no cover - never instantiated by the oracle
    def __init__(self, *a, **k):
        raise RuntimeError("the simulator-backed environment is not available offline")
