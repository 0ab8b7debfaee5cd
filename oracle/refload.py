"""Loader for the unmodified reference (TEST INFRASTRUCTURE; see oracle/make_ref.py).

`load()` puts oracle/shims and the reference root on sys.path and imports the reference's own modules;
the root is oracle/_ref/ when it was materialised (it travels to the GPU box), else /root/reference
(build container). Returns None when neither exists. Only tests/, bench.py's CPU legs and
__graft_entry__ may import this module."""
import importlib
import os
import sys
import types
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))


def root():
    for cand in (os.path.join(HERE, "_ref"), os.environ.get("LSTEP_REFERENCE", "/root/reference")):
        if cand and os.path.exists(os.path.join(cand, "models", "LSTEP.py")):
            return cand
    return None


_cache = None


def load():
    """-> namespace(root, LSTEP, MergeLayer, TimeEncoder, Data, get_idx_data_loader, NeighborSampler, get_neighbor_sampler,
    NegativeEdgeSampler, evaluate_model_link_prediction, evaluate_model_utils, EarlyStopping, PositionalEncoding) or None."""
    global _cache
    if _cache is not None:
        return _cache
    r = root()
    if r is None:
        return None
    for p in (r, os.path.join(HERE, "shims")):
        if p not in sys.path:
            sys.path.insert(0, p)
    # the reference's top-level packages are called `models` and `utils`: drop foreign modules of those names
    for name in ("models", "utils"):
        m = sys.modules.get(name)
        if m is not None and not str(getattr(m, "__file__", None) or getattr(m, "__path__", "")).count(r):
            for k in [k for k in sys.modules if k == name or k.startswith(name + ".")]:
                del sys.modules[k]
    warnings.filterwarnings("ignore")
    ns = types.SimpleNamespace(root=r)
    lstep = importlib.import_module("models.LSTEP")
    modules = importlib.import_module("models.modules")
    dl = importlib.import_module("utils.DataLoader")
    uu = importlib.import_module("utils.utils")
    emu = importlib.import_module("evaluate_model_utils")
    ns.LSTEP, ns.MergeLayer, ns.TimeEncoder = lstep.LSTEP, modules.MergeLayer, modules.TimeEncoder
    ns.Data, ns.get_idx_data_loader = dl.Data, dl.get_idx_data_loader
    ns.NeighborSampler, ns.get_neighbor_sampler, ns.NegativeEdgeSampler = uu.NeighborSampler, uu.get_neighbor_sampler, uu.NegativeEdgeSampler
    ns.evaluate_model_utils, ns.evaluate_model_link_prediction = emu, emu.evaluate_model_link_prediction
    ns.EarlyStopping = importlib.import_module("utils.EarlyStopping").EarlyStopping
    ns.PositionalEncoding = importlib.import_module("utils.PositionalEncoding")
    # silence the loop's progress bar (stdout hygiene for bench.py): same iteration, no output
    emu.tqdm = lambda it, **kw: type("Q", (), {"__iter__": lambda s: iter(it), "set_description": lambda s, *_: None})()
    _cache = ns
    return ns
