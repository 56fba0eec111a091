"""rdkit placeholder: imported by analysis/pharm_builder.py:4 and dataset modules, unused on the hot path."""
from . import Chem  # noqa: F401
