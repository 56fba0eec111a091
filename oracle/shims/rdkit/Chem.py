"""rdkit.Chem placeholder (import-only)."""
