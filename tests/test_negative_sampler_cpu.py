"""SURVEY §8(f) f3 — lstep_b200.NegativeEdgeSampler (incremental history set, lazily built possible_edges) must be
RNG-bit-identical to the reference's utils/utils.py:304-494 for all three strategies: against fixtures written by the
unmodified reference (tests/golden/make_golden.py::gen_negatives) and, where oracle/_ref or /root/reference is present,
against the reference's own class live on a second graph (incl. a non-chronological stream: full-scan fall-back)."""
import numpy as np
import pytest

from common import checksum, golden_path
from lstep_b200 import synth
from lstep_b200.negative import NegativeEdgeSampler


def _sample(s, strat, g, lo, hi):
    if strat == "random":
        return s.sample(size=hi - lo)
    return s.sample(size=hi - lo, batch_src_node_ids=g.src_node_ids[lo:hi], batch_dst_node_ids=g.dst_node_ids[lo:hi],
                    current_batch_start_time=g.node_interact_times[lo], current_batch_end_time=g.node_interact_times[hi - 1])


@pytest.mark.parametrize("strat", ["random", "historical", "inductive"])
def test_negatives_match_reference_golden(strat):
    z = np.load(golden_path("negatives.npz"))
    g = synth.make_graph("tiny_bip", seed=3, num_nodes=300, num_edges=14000)
    assert np.allclose(checksum(g.node_interact_times), z["graph_ck"], rtol=0, atol=1e-6)
    E, B = g.num_edges, int(z["B"])
    s = NegativeEdgeSampler(g.src_node_ids, g.dst_node_ids, interact_times=g.node_interact_times,
                            last_observed_time=g.node_interact_times[int(E * 0.7)], negative_sample_strategy=strat, seed=2)
    assert s._possible_edges is None  # not materialised by the constructor
    for rep in range(2):
        s.reset_random_state()  # (pass 2 goes back in time: the history set is rebuilt)
        for i, lo in enumerate(z["starts"].tolist()):
            a, b = _sample(s, strat, g, lo, lo + B)
            assert a.dtype == np.int64 and b.dtype == np.int64
            assert np.array_equal(a, z[f"{strat}_{rep}_src"][i]) and np.array_equal(b, z[f"{strat}_{rep}_dst"][i]), (strat, rep, i)
    if strat != "random":
        assert s._possible_edges is not None and len(s.possible_edges) == len(np.unique(g.src_node_ids)) * len(np.unique(g.dst_node_ids))
        assert s._hist_n == E - B + np.searchsorted(g.node_interact_times[E - B:], g.node_interact_times[E - B], side="right")
    with pytest.raises(ValueError):
        NegativeEdgeSampler(g.src_node_ids, g.dst_node_ids, interact_times=g.node_interact_times, negative_sample_strategy="bogus", seed=1).sample(3)


@pytest.mark.parametrize("chronological", [True, False])
def test_negatives_match_reference_live(chronological):
    from oracle import refload
    ref = refload.load()
    if ref is None:
        pytest.skip("reference not available (oracle/_ref not materialised)")
    g = synth.make_graph("tiny_ties", seed=1)  # heavy timestamp ties: window boundaries fall inside runs of equal times
    src, dst, t = g.src_node_ids.copy(), g.dst_node_ids.copy(), g.node_interact_times.copy()
    if not chronological:
        p = np.random.default_rng(0).permutation(len(t))
        src, dst, t = src[p], dst[p], t[p]
    E, B = len(t), 40
    for strat in ("historical", "inductive"):
        kw = dict(interact_times=t, last_observed_time=float(np.sort(t)[int(E * 0.6)]), negative_sample_strategy=strat, seed=5)
        a, b = ref.NegativeEdgeSampler(src, dst, **kw), NegativeEdgeSampler(src, dst, **kw)
        order = np.argsort(t, kind="stable")
        for rep in range(2):
            a.reset_random_state()
            b.reset_random_state()
            for lo in list(range(0, 5 * B, B)) + list(range(E - 12 * B, E, B)):
                idx = order[lo:lo + B] if not chronological else np.arange(lo, lo + B)
                args = dict(size=B, batch_src_node_ids=src[idx], batch_dst_node_ids=dst[idx], current_batch_start_time=float(t[idx].min()),
                            current_batch_end_time=float(t[idx].max()))
                ra, rb = a.sample(**args), b.sample(**args)
                assert all(np.array_equal(x, y) and x.dtype == y.dtype for x, y in zip(ra, rb)), (strat, rep, lo)
        assert b._chronological == chronological
