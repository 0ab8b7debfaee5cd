"""Synthetic temporal graphs in the shapes of the datasets BASELINE.json names.

Datasets are not available offline, so every parity test and benchmark runs on a seeded
synthetic edge stream that follows the id conventions of the reference's preprocessing
(node ids start at 1, row 0 is the padding node, bipartite graphs put the destination side
after the source side: /root/reference/preprocess_data/preprocess_data.py:56-81,101-108) and
the chronological ordering the loaders rely on (utils/DataLoader.py:199,229-242).

Nothing here imports torch; arrays are plain numpy so the oracle, the tests and bench.py can
share the generator.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

# name -> (num_nodes, num_edges, bipartite split (src side count) or None, edge feature dim,
#          time span, integer-day timestamps?)
SHAPES = {
    # Enron-scale timestamps (1.1e8) exercise the fp32 rounding of current_time (Q4)
    "enron": dict(num_nodes=184, num_edges=125_235, src_side=None, edge_dim=32, t_span=1.1e8, ties=False),
    "wikipedia": dict(num_nodes=9_227, num_edges=157_474, src_side=8_227, edge_dim=172, t_span=2.68e6, ties=False),
    "reddit": dict(num_nodes=10_984, num_edges=672_447, src_side=10_000, edge_dim=172, t_span=2.68e6, ties=False),
    # Flights: integer day stamps, ~122 distinct values -> heavy ties exercise strict '<' semantics
    "flights": dict(num_nodes=13_169, num_edges=1_927_145, src_side=None, edge_dim=1, t_span=122.0, ties=True),
    # scale-out graph of config 5 (generated per shard, see shard.py); listed for completeness
    "scaleout": dict(num_nodes=10_000_000, num_edges=500_000_000, src_side=None, edge_dim=0, t_span=2.68e6, ties=False),
    # tiny graph for smoke tests / goldens
    "tiny": dict(num_nodes=60, num_edges=1_500, src_side=None, edge_dim=8, t_span=5.0e4, ties=False),
    "tiny_bip": dict(num_nodes=90, num_edges=2_000, src_side=60, edge_dim=8, t_span=9.0e7, ties=False),
    "tiny_ties": dict(num_nodes=50, num_edges=1_800, src_side=None, edge_dim=4, t_span=12.0, ties=True),
}


@dataclass
class TemporalGraph:
    """Edge stream in the reference's `Data` layout (utils/DataLoader.py:68-86)."""

    name: str
    num_nodes: int  # V (ids 1..V); tables have V+1 rows
    src_node_ids: np.ndarray  # int64 [E]
    dst_node_ids: np.ndarray  # int64 [E]
    node_interact_times: np.ndarray  # float64 [E], ascending
    edge_ids: np.ndarray  # int64 [E], 1..E
    edge_dim: int

    @property
    def num_edges(self) -> int:
        return int(self.src_node_ids.shape[0])

    @property
    def labels(self) -> np.ndarray:
        return np.zeros(self.num_edges, dtype=np.float64)

    def slice(self, lo: int, hi: int) -> "TemporalGraph":
        return TemporalGraph(self.name, self.num_nodes, self.src_node_ids[lo:hi], self.dst_node_ids[lo:hi],
                             self.node_interact_times[lo:hi], self.edge_ids[lo:hi], self.edge_dim)


def _zipf_ids(rng: np.random.Generator, n_ids: int, size: int, s: float) -> np.ndarray:
    """Truncated Zipf(s) over a random permutation of 0..n_ids-1."""
    ranks = np.arange(1, n_ids + 1, dtype=np.float64)
    p = ranks ** (-s)
    p /= p.sum()
    cdf = np.cumsum(p)
    draws = np.searchsorted(cdf, rng.random(size), side="left")
    np.minimum(draws, n_ids - 1, out=draws)
    perm = rng.permutation(n_ids)
    return perm[draws].astype(np.int64)


def make_graph(name: str, seed: int = 0, num_edges: int | None = None, num_nodes: int | None = None,
               zipf_s: float = 0.8) -> TemporalGraph:
    """Seeded synthetic edge stream of the named dataset shape.

    `num_edges` / `num_nodes` override the shape (used for bounded samples of the same workload).
    Endpoint popularity is a truncated Zipf; s=0.8 gives ~300 unique nodes per 200-edge batch on
    the Reddit shape and ~2500 per 2000-edge batch on the Flights shape, the N the survey's byte
    model assumes (s=1.2 collapses a batch onto ~165 nodes; SURVEY §8(d) flags that as too skewed).
    """
    spec = dict(SHAPES[name])
    V = int(num_nodes if num_nodes is not None else spec["num_nodes"])
    E = int(num_edges if num_edges is not None else spec["num_edges"])
    rng = np.random.default_rng(np.random.PCG64(seed))
    src_side = spec["src_side"]
    if src_side is not None and num_nodes is not None:
        src_side = max(1, int(round(src_side * V / spec["num_nodes"])))
    if src_side is None:
        src = _zipf_ids(rng, V, E, zipf_s) + 1
        dst = _zipf_ids(rng, V, E, zipf_s) + 1
    else:
        src = _zipf_ids(rng, src_side, E, zipf_s) + 1
        dst = _zipf_ids(rng, V - src_side, E, zipf_s) + 1 + src_side
    if spec["ties"]:
        t = np.sort(np.floor(rng.random(E) * spec["t_span"])).astype(np.float64)
    else:
        t = np.sort(rng.random(E) * spec["t_span"]).astype(np.float64)
    eid = np.arange(1, E + 1, dtype=np.int64)
    return TemporalGraph(name, V, src, dst, t, eid, int(spec["edge_dim"]))


def make_initial_pe(num_nodes: int, pe_dim: int, seed: int = 1, scale: float = 0.1) -> np.ndarray:
    """PE table [V+1, d] fp32 ~ N(0, scale^2) with the padding row zero."""
    rng = np.random.default_rng(np.random.PCG64(seed))
    pe = (rng.standard_normal((num_nodes + 1, pe_dim)) * scale).astype(np.float32)
    pe[0] = 0.0
    return pe


def batches(graph: TemporalGraph, batch_size: int, start: int = 0, stop: int | None = None):
    """Yield (lo, hi) index ranges in stream order, like get_idx_data_loader(shuffle=False)
    (utils/DataLoader.py:51-65)."""
    stop = graph.num_edges if stop is None else stop
    for lo in range(start, stop, batch_size):
        yield lo, min(lo + batch_size, stop)


def unique_batch_nodes(src: np.ndarray, dst: np.ndarray) -> np.ndarray:
    """Sorted unique ids of a batch, as the loops compute them
    (train_LSTEP_link_prediction.py:221-222, evaluate_model_utils.py:54-55)."""
    return np.unique(np.concatenate([src, dst])).astype(np.int64)


def make_scaleout_device(num_nodes: int, num_edges: int, device, seed: int = 0, zipf_s: float = 0.8, t_span: float = 2.68e6,
                         chunk: int = 50_000_000):
    """The scale-out edge stream (BASELINE config 5) generated ON the device — 5e8 edges are 12 GB and take minutes with
    numpy on the host, and every rank of a group needs the same stream: a seeded CUDA Philox generator gives every rank
    the same uniform draws (torch.randperm on the device is NOT reproducible across GPUs — measured: ranks disagreed — so
    the popularity ranks are scattered over the ids by an affine bijection instead of a random permutation). Otherwise the
    construction of make_graph: endpoints = truncated Zipf(s) over scrambled ids 1..V, timestamps = sorted uniform
    float64. Returns (src, dst, t) device tensors."""
    import torch
    dev = torch.device(device)
    gen = torch.Generator(device=dev).manual_seed(seed)
    # the CDF on the host (a parallel prefix sum on the device rounds differently from run to run / device to device —
    # measured: the low bits differed between two GPUs — and a draw next to a CDF step would then pick another node)
    p = np.arange(1, num_nodes + 1, dtype=np.float64) ** (-zipf_s)
    cdf_h = np.cumsum(p)
    cdf_h /= cdf_h[-1]
    cdf = torch.from_numpy(cdf_h).to(dev)
    del p, cdf_h
    out = []
    import math
    for side in range(2):
        mul = 7_368_787 + 2 * side  # affine bijection rank -> id: (mul * r + add) mod V with gcd(mul, V) = 1
        while math.gcd(mul, num_nodes) != 1:
            mul += 2
        add = (1_234_567 * (seed + 1) + 7_654_321 * side) % num_nodes
        ids = torch.empty(num_edges, dtype=torch.int64, device=dev)
        for lo in range(0, num_edges, chunk):
            hi = min(lo + chunk, num_edges)
            u = torch.rand(hi - lo, generator=gen, device=dev, dtype=torch.float64)
            r = torch.searchsorted(cdf, u).clamp_(max=num_nodes - 1)
            ids[lo:hi] = (r * mul + add) % num_nodes + 1
            del u, r
        out.append(ids)
    del cdf
    t = torch.rand(num_edges, generator=gen, device=dev, dtype=torch.float64).mul_(t_span)
    t = torch.sort(t).values
    return out[0], out[1], t
