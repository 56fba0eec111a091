"""dgl.data.DGLDataset placeholder (protein_pharm_dataset.py:18 subclasses it; never instantiated here)."""


class DGLDataset:
    def __init__(self, name=None, **kwargs):
        self._name = name
