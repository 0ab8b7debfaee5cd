"""lstep_b200 — B200-native (sm_100a) hot path of L-STEP's positional-encoding step.

The package directory is `l-step_b200/` (repo layout); it is importable as `lstep_b200`
through the alias package at the repo root. Heavy submodules (torch, the CUDA library) are
loaded lazily so that the oracle and the CPU-only tests can import `lstep_b200.synth` alone.
"""
from importlib import import_module as _imp

__version__ = "0.1.0"

_LAZY = {
    "NeighborSampler": "sampler",
    "get_neighbor_sampler": "sampler",
    "LSTEP": "model",
    "TimeEncoder": "modules",
    "MergeLayer": "modules",
    "PEStream": "stream",
    "ChangeLogStream": "stream",
    "ShardRank": "shard",
    "LocalGroup": "shard",
    "DistGroup": "shard",
    "ShardedPEStream": "shard",
    "ReplicatedTableRank": "shard",
    "ReplicatedLocalGroup": "shard",
    "ReplicatedTableStream": "shard",
    "PeerRank": "peer",
    "PeerLocalGroup": "peer",
    "LaplacianPE": "pe_init",
    "RandomWalkPE": "pe_init",
    "save_pe": "pe_init",
    "load_pe": "pe_init",
    "NegativeEdgeSampler": "negative",
}


def __getattr__(name):
    if name in _LAZY:
        return getattr(_imp(f"{__name__}.{_LAZY[name]}"), name)
    raise AttributeError(name)
