"""Helpers shared by the golden-vector generator (make_golden.py, runs the unmodified reference
in the build container) and by the tests that consume the fixtures (which never read
/root/reference). Inputs that would be large to commit are regenerated from a seed here and
pinned by a checksum stored in the fixture."""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def golden_path(name: str) -> str:
    return os.path.join(HERE, name)


def seeded_normal(seed: int, shape, scale: float = 0.1) -> np.ndarray:
    rng = np.random.default_rng(np.random.PCG64(seed))
    return (rng.standard_normal(shape) * scale).astype(np.float32)


def seeded_edge_feats(num_edges: int, feat_dim: int, seed: int = 1) -> np.ndarray:
    """Edge features [E+1, feat_dim] ~ N(0,1), row 0 (padding edge) zero."""
    rng = np.random.default_rng(seed)
    f = rng.standard_normal((num_edges + 1, feat_dim)).astype(np.float32)
    f[0] = 0
    return f


def checksum(a: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(a)
    return np.array([np.float64(a.astype(np.float64).sum()), np.float64(np.abs(a.astype(np.float64)).sum())])


def load_params(name: str) -> dict:
    with np.load(golden_path(name)) as z:
        return {k: z[k] for k in z.files}


def pe_close(a: np.ndarray, b: np.ndarray, rtol: float = 1e-5):
    """The parity bar for fp32 PE values (BASELINE.json north_star: 1e-5 relative), in the form
    SURVEY §7 gives for values near zero: |a-b| <= rtol * max(|b|, rms(b)). Returns
    (ok, worst ratio)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    rms = float(np.sqrt(np.mean(b * b))) if b.size else 0.0
    denom = np.maximum(np.abs(b), max(rms, 1e-30))
    ratio = np.abs(a - b) / denom
    worst = float(ratio.max()) if ratio.size else 0.0
    return worst <= rtol, worst
