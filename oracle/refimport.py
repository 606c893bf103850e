"""Import the UNMODIFIED reference from /root/reference (build container only).

Used by oracle/make_golden.py and tests/test_oracle_vs_reference.py (skipped
when /root/reference is absent, e.g. on the GPU box).  Nothing is copied: the
reference's own modules are put on sys.path together with two stub packages
for dependencies that are missing from this image.
"""
import os
import sys

REF = os.environ.get("GDSS_REFERENCE_ROOT", "/root/reference")
_STUBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_stubs")


def available() -> bool:
    return os.path.isfile(os.path.join(REF, "solver.py"))


def activate() -> str:
    if not available():
        raise RuntimeError(f"reference tree not found at {REF}")
    for p in (_STUBS, REF):
        if p not in sys.path:
            sys.path.insert(0, p)
    return REF


def load_ckpt(rel: str):
    """rel e.g. 'ZINC250k/gdss_zinc250k_v2' -> the raw checkpoint dict."""
    import torch
    activate()
    return torch.load(os.path.join(REF, "checkpoints", rel + ".pth"), map_location="cpu", weights_only=False)
