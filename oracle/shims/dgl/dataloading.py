"""dgl.dataloading.GraphDataLoader placeholder (import-only; protein_pharm_dataset.py:9)."""


class GraphDataLoader:
    def __init__(self, *a, **k):
        raise NotImplementedError("dataset plumbing is out of scope for the oracle")
