"""Import the reference's own model code UNMODIFIED over the shim packages.

TEST INFRASTRUCTURE ONLY; works only where /root/reference exists (this
container, never the GPU box).  Used by oracle/make_golden.py to generate the
fixtures under tests/golden/ and by the CPU tests that pin oracle/pf_oracle.py
against the reference when it is present.
"""
import os
import sys

REFERENCE_ROOT = os.environ.get("PF_REFERENCE_ROOT", "/root/reference")
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "pharmacoforge", "models"))


def load():
    """Returns the reference's `pharmacoforge` package (models imported)."""
    if not available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    for p in (_SHIMS, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    import pharmacoforge  # noqa: F401
    import pharmacoforge.models.pharmacodiff  # noqa: F401
    import pharmacoforge.models.dynamics_gvp  # noqa: F401
    import pharmacoforge.models.gvp  # noqa: F401
    import pharmacoforge.utils  # noqa: F401
    import pharmacoforge.dataset.protein_pharm_dataset  # noqa: F401
    return pharmacoforge
