"""TEST INFRASTRUCTURE ONLY (oracle shim): name imported by modeling_utils.py:17."""


def cached_path(p, *a, **k):
    return p
