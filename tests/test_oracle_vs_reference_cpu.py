"""The oracle against the UNMODIFIED reference itself (oracle/_ref, torch CPU) on a hub-heavy workload — CPU only.

The committed goldens and the bench use endpoint skew zipf_s = 0.8; SURVEY §8(d) names 1.2, which collapses a
200-edge Reddit-shaped batch onto ~160 nodes and ~730 sampled neighbours: hub rows collect several times more phase-A /
phase-B contributions and row 0 most padded slots. This pins the oracle (the checker of the `-m gpu` parity tests) on exactly
that regime; the CUDA path meets the same case in tests/test_reference_loop_gpu.py::...[reddit-hubs-200-20].
Skipped when oracle/_ref is absent (it is materialised by oracle/make_ref.py where /root/reference exists)."""
import numpy as np
import pytest

from common import check_updated_table, pe_close
from harness import lstep_params_np
from lstep_b200 import synth
from oracle import lstep_oracle as orc


@pytest.fixture(scope="module")
def ref():
    from oracle import refload
    ns = refload.load()
    if ns is None:
        pytest.skip("oracle/_ref not materialised (run `python oracle/make_ref.py` where /root/reference exists)")
    return ns


def test_oracle_matches_the_reference_on_a_hub_heavy_batch(ref):
    import torch
    gname, B, K, T, d = "reddit", 200, 20, 100, 172
    g = synth.make_graph(gname, seed=0, zipf_s=1.2)
    V1 = g.num_nodes + 1
    lo = g.num_edges - B
    src, dst, tt, ee = (a[lo:lo + B] for a in (g.src_node_ids, g.dst_node_ids, g.node_interact_times, g.edge_ids))
    ids = synth.unique_batch_nodes(src, dst)
    assert len(ids) < 200  # the skew is real: fewer batch nodes than edges (0.8 gives ~300)
    neg = np.random.default_rng(5).choice(g.dst_node_ids, B)
    queries = [(src, tt), (dst, tt), (neg, tt)]
    adj = orc.build_adjacency_fast(g.src_node_ids, g.dst_node_ids, g.edge_ids, g.node_interact_times, num_rows=V1)
    rs = ref.get_neighbor_sampler(ref.Data(g.src_node_ids, g.dst_node_ids, g.node_interact_times, g.edge_ids, g.labels), "recent", seed=1)
    # the lookups of the batch's own rows and of the negatives: bit-exact, incl. the hubs' long adjacency rows
    q_ids, q_t = np.concatenate([src, dst, neg]), np.concatenate([tt, tt, tt])
    for KK in (K, 2000):
        for x, y in zip(rs.get_historical_neighbors(q_ids, q_t, KK), orc.sample_recent(adj, q_ids, q_t, KK)):
            assert x.dtype == y.dtype and np.array_equal(x, y), KK
    torch.manual_seed(3)
    hist = torch.randn((V1, T, d)) * 0.3
    for tag in ("fullu", "full"):
        p = lstep_params_np(tag)
        rm = ref.LSTEP(np.zeros((V1, 172), np.float32), np.zeros((2, 172), np.float32), rs, rs, pe_dim=d, num_neighbors=K,
                       time_feat_dim=100, num_fft_batches=T, device="cpu")
        rm.load_state_dict({k: torch.from_numpy(v) for k, v in p.items()})
        rm.eval()
        with torch.no_grad():
            fft_r = rm.fourier_transform_pe(ids, hist, 50)
            cur_r = torch.clone(hist[:, -1, :])
            cur_r[torch.from_numpy(ids)] = fft_r
            outs_r = [rm.compute_neighborhood_pe(cur_r, qi, qt, num_neighbors=K).numpy() for qi, qt in queries]
            cur_in = cur_r.numpy().copy()
            assert rm.update_pe(cur_r, ids, ee, src, dst, tt, tt.max(), num_neighbors=K) is cur_r
        fft_o = orc.fourier_transform_pe(p, np.arange(len(ids)), hist[torch.from_numpy(ids)].numpy(), 50, T)
        ok, worst = pe_close(fft_o, fft_r.numpy())
        assert ok, (tag, "dft", worst)
        for c, ((qi, qt), want) in enumerate(zip(queries, outs_r)):
            ok, worst = pe_close(orc.compute_neighborhood_pe(p, adj, cur_in.copy(), qi, qt, K), want)
            assert ok, (tag, "nbr", c, worst)
        cur_o = orc.update_pe(p, adj, cur_in.copy(), ids, src, dst, tt, tt.max(), K)
        with orc.high_precision():
            cur_t = orc.update_pe(p, adj, cur_in.astype(np.float64), ids, src, dst, tt, tt.max(), K)
        rep = check_updated_table(cur_o, cur_r.numpy(), f"oracle-vs-reference/reddit-zipf1.2/{tag}/update", cur_t)
        # rows the batch did not touch are untouched, row 0 is rewritten by phase B (Q2)
        touched = np.zeros(V1, bool)
        touched[np.any(cur_o != cur_in, axis=1)] = True
        assert touched[0] and touched[ids].all() and rep["n"] == cur_o.size
