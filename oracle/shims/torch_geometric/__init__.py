"""Import shim for the un-vendored `torch_geometric` dependency of the reference's PE initialisation
(utils/PositionalEncoding.py:3-7). Only `torch_geometric.utils` names that module imports exist."""
