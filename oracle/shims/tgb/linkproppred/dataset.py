"""Import shim: utils/DataLoader.py:5 imports tgb at module level; nothing on the PE path uses it."""


class LinkPropPredDataset:  # pragma: no cover
    def __init__(self, *a, **k):
        raise RuntimeError("tgb is not available offline")
