"""torch_scatter is imported but never called by the reference (pharmacodiff.py:12,21)."""


def segment_coo(*a, **k):
    raise NotImplementedError


def segment_csr(*a, **k):
    raise NotImplementedError
