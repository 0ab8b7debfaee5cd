"""The drop-in claim, proven against the UNMODIFIED reference itself (oracle/_ref, materialised by oracle/make_ref.py
from /root/reference; it travels to the GPU box with the snapshot like the built .so):

  * the reference's own evaluate_model_link_prediction (evaluate_model_utils.py:19-144), untouched, drives
    lstep_b200.LSTEP + lstep_b200.get_neighbor_sampler on the B200 and must reproduce the reference's per-batch
    link probabilities, AP, AUC and losses (fixtures replay_*.npz were written by the same loop driving the reference's
    own modules on CPU);
  * one teacher-forced module-boundary step of every BASELINE dataset shape at FULL size (all edges, T = 100) with the
    reference's LSTEP running on the box's CPU cores beside the CUDA path on the same inputs and weights.

Skipped (with the reason) when oracle/_ref is absent.
"""
import numpy as np
import pytest

from common import check_updated_table, golden_path, pe_close, seeded_edge_feats, seeded_normal
from lstep_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref():
    from oracle import refload
    ns = refload.load()
    if ns is None:
        pytest.skip("oracle/_ref not materialised (run `python oracle/make_ref.py` where /root/reference exists)")
    return ns


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available()
    from lstep_b200 import _lib
    assert _lib.load().lstep_device_ok() == 1
    return torch


@pytest.mark.parametrize("tag", ["small", "fullu", "full"])
def test_untouched_reference_eval_loop_drives_the_dropin(torch_cuda, ref, tag, parity_log):
    torch = torch_cuda
    import torch.nn as nn
    from harness import build_dropin, check_rank_metrics, oracle_replay_f64
    from lstep_b200 import get_neighbor_sampler
    z = np.load(golden_path(f"replay_{tag}.npz"))
    d, T, K, t_dim, F, tg, B = (int(z[k]) for k in ("pe_dim", "T", "K", "time_dim", "feat_dim", "time_gap", "B"))
    V, E, e0 = int(z["V"]), int(z["E"]), int(z["e0"])
    g = synth.make_graph("tiny", seed=int(z["graph_seed"]), num_nodes=V, num_edges=E)
    full = ref.Data(g.src_node_ids, g.dst_node_ids, g.node_interact_times, g.edge_ids, g.labels)
    sampler = get_neighbor_sampler(full, "recent", seed=1)  # the drop-in for utils/utils.py:282-301
    model = build_dropin(tag, g, sampler, F, d, t_dim, T, K, edge_feats=seeded_edge_feats(E, F))
    ev = g.slice(e0, E)
    eval_data = ref.Data(ev.src_node_ids, ev.dst_node_ids, ev.node_interact_times, ev.edge_ids, ev.labels)
    neg = ref.NegativeEdgeSampler(g.src_node_ids, g.dst_node_ids, seed=2)  # the reference's own sampler and RNG
    loader = ref.get_idx_data_loader(list(range(ev.num_edges)), batch_size=B, shuffle=False)
    hist0 = seeded_normal(int(z["hist0_seed"]), (V + 1, 1, d), 0.3)
    hist0[0] = 0
    emu = ref.evaluate_model_utils
    preds, last = [], {}
    orig_metrics, orig_update = emu.get_link_prediction_metrics, model[0].update_pe

    def metrics_hook(predicts, labels):
        preds.append(predicts.detach().cpu().numpy().astype(np.float32))
        return orig_metrics(predicts=predicts, labels=labels)

    def update_hook(*a, **k):
        r = orig_update(*a, **k)
        last["pe"] = r
        return r

    emu.get_link_prediction_metrics = metrics_hook
    model[0].update_pe = update_hook
    try:
        losses, metrics = ref.evaluate_model_link_prediction("LSTEP", model, torch.from_numpy(hist0).cuda(), sampler, loader, neg,
                                                             eval_data, nn.BCELoss(), num_fft_batches=T, num_neighbors=K, time_gap=tg)
    finally:
        emu.get_link_prediction_metrics = orig_metrics
        del model[0].update_pe
    ap = np.array([m["average_precision"] for m in metrics])
    auc = np.array([m["roc_auc"] for m in metrics])
    assert len(ap) == len(z["ap"])
    rep = check_rank_metrics(preds, z["predicts"], B, ap, auc, z["ap"], z["auc"], f"reference_loop/{tag}/ap_auc", parity_log)
    assert rep["mean_ap_diff"] < 2e-5 and rep["mean_auc_diff"] < 2e-5, rep
    assert np.abs(np.array(losses) - z["losses"]).max() < 1e-4
    check_updated_table(last["pe"].cpu().numpy(), z["last_pe"], f"reference_loop/{tag}/final_table", oracle_replay_f64(tag), factor=2.0,
                        log=parity_log)
    sampler.check_errors()


def _ref_model(ref, rs, V1, tag, K, T):
    """The reference's LSTEP (CPU) carrying the golden parameters of `tag`, on the reference's own sampler `rs`."""
    import torch
    from harness import lstep_params_np
    m = ref.LSTEP(np.zeros((V1, 172), np.float32), np.zeros((2, 172), np.float32), rs, rs, pe_dim=172, num_neighbors=K,
                  time_feat_dim=100, num_fft_batches=T, device="cpu")
    m.load_state_dict({k: torch.from_numpy(v) for k, v in lstep_params_np(tag).items()})
    return m.eval()


# zipf_s = 1.2 is SURVEY §8(d)'s endpoint skew (the bench uses 0.8): a 200-edge batch collapses onto ~160 nodes and
# ~730 sampled neighbours, so hub rows collect several times more phase-A / phase-B contributions than at 0.8
@pytest.mark.parametrize("gname,B,K,zipf_s", [("enron", 200, 20, 0.8), ("wikipedia", 200, 20, 0.8), ("reddit", 200, 20, 0.8),
                                              ("flights", 2000, 20, 0.8), ("reddit", 200, 20, 1.2)],
                         ids=["enron-200-20", "wikipedia-200-20", "reddit-200-20", "flights-2000-20", "reddit-hubs-200-20"])
def test_full_size_step_vs_the_reference_itself(torch_cuda, ref, gname, B, K, zipf_s, parity_log):
    torch = torch_cuda
    from harness import build_dropin, lstep_params_np
    from lstep_b200 import NeighborSampler
    from oracle import lstep_oracle as orc
    g = synth.make_graph(gname, seed=0, zipf_s=zipf_s)
    if zipf_s != 0.8:
        gname = f"{gname}-zipf{zipf_s}"  # (key of the parity log)
    d, T = 172, 100
    V1 = g.num_nodes + 1
    lo = g.num_edges - B
    src, dst, tt, ee = (a[lo:lo + B] for a in (g.src_node_ids, g.dst_node_ids, g.node_interact_times, g.edge_ids))
    ids = synth.unique_batch_nodes(src, dst)
    neg = np.random.default_rng(5).choice(g.dst_node_ids, B)
    queries = [(src, tt), (dst, tt), (src, tt), (neg, tt)]
    torch.manual_seed(3)
    hist = torch.randn((V1, T, d)) * 0.3  # host copy for the reference, device copy for the CUDA path
    pe_h = hist.cuda()
    s = NeighborSampler.from_edges(g.src_node_ids, g.dst_node_ids, g.edge_ids, g.node_interact_times, "recent", num_rows=V1)
    adj = orc.build_adjacency_fast(g.src_node_ids, g.dst_node_ids, g.edge_ids, g.node_interact_times, num_rows=V1)
    rs = ref.get_neighbor_sampler(ref.Data(g.src_node_ids, g.dst_node_ids, g.node_interact_times, g.edge_ids, g.labels), "recent", seed=1)
    for tag in ("fullu", "full"):
        rm = _ref_model(ref, rs, V1, tag, K, T)
        lstep = build_dropin(tag, g, s, 172, d, 100, T, K)[0].eval()
        p = lstep_params_np(tag)
        with torch.no_grad():
            # ---- the reference, on the CPU
            fft_r = rm.fourier_transform_pe(ids, hist, 50)
            cur_r = torch.clone(hist[:, -1, :])
            cur_r[torch.from_numpy(ids)] = fft_r
            outs_r = [rm.compute_neighborhood_pe(cur_r, qi, qt, num_neighbors=K).numpy() for qi, qt in queries]
            cur_in = cur_r.numpy().copy()
            ret = rm.update_pe(cur_r, ids, ee, src, dst, tt, tt.max(), num_neighbors=K)
            assert ret is cur_r
            # ---- float64 evaluation of the same functions on the same inputs
            with orc.high_precision():
                outs_t = [orc.compute_neighborhood_pe(p, adj, cur_in.astype(np.float64), qi, qt, K) for qi, qt in queries]
                cur_t = orc.update_pe(p, adj, cur_in.astype(np.float64), ids, src, dst, tt, tt.max(), K)
            # ---- the CUDA path
            fft = lstep.fourier_transform_pe(ids, pe_h, 50)
            ok, worst = pe_close(fft.cpu().numpy(), fft_r.numpy())
            parity_log[f"reference/{gname}/{tag}/dft"] = {"max": worst, "N": int(len(ids))}
            assert ok, (gname, tag, "dft", worst)
            cur = torch.from_numpy(cur_in).cuda()  # teacher-forced: both sides continue from the reference's table
            for c, ((qi, qt), want, truth) in enumerate(zip(queries, outs_r, outs_t)):
                got = lstep.compute_neighborhood_pe(cur, qi, qt, num_neighbors=K)
                check_updated_table(got.cpu().numpy(), want, f"reference/{gname}/{tag}/nbr{c}", truth, log=parity_log)
            lstep.update_pe(cur, ids, ee, src, dst, tt, tt.max(), num_neighbors=K)
        check_updated_table(cur.cpu().numpy(), cur_r.numpy(), f"reference/{gname}/{tag}/update", cur_t, log=parity_log)
    # sampler: the reference's own lookups vs the device sampler on this full-size graph (bit-exact)
    q_ids = np.concatenate([src, dst, neg])
    q_t = np.concatenate([tt, tt, tt])
    for KK in (K, 2000):
        a = rs.get_historical_neighbors(q_ids, q_t, KK)
        b = s.get_historical_neighbors(q_ids, q_t, KK)
        for x, y in zip(a, b):
            assert x.dtype == y.dtype and np.array_equal(x, y), (gname, KK)
    del pe_h
    torch.cuda.empty_cache()
