"""Test-side helpers: load golden parameters into the drop-in modules and replay the eval
loop's per-batch sequence (the caller code of evaluate_model_utils.py:38-142, restated here
because the reference cannot be imported on the GPU box)."""
from __future__ import annotations

import numpy as np
import torch

from common import load_params


PE_MLPS = ("pe_mlp_1", "pe_mlp_2", "self_update_pe", "pe_neighbor_mlp_1", "pe_neighbor_mlp_2", "self_update_neighbor_pe")


def golden_params(tag: str) -> dict:
    """Parameters of a golden model. Tags ending in `u` are the UNSCALED regime: the same constructor call as the base
    tag (tests/golden/make_golden.py::build_model) without the x2 on the six PE-MLP weight matrices — halving is
    exact in fp32, so they are derived from the base file instead of being committed twice."""
    if tag.endswith("u"):
        p = load_params(f"params_{tag[:-1]}.npz")
        for k in list(p):
            if k.startswith("0.") and k.endswith(".weight") and k[2:-7] in PE_MLPS:
                p[k] = (p[k] * np.float32(0.5)).astype(np.float32)
        return p
    return load_params(f"params_{tag}.npz")


def lstep_params_np(tag: str) -> dict:
    return {k[2:]: v for k, v in golden_params(tag).items() if k.startswith("0.")}


def build_dropin(tag, graph, sampler, feat_dim, pe_dim, time_dim, T, K, edge_feats=None, device="cuda"):
    from lstep_b200 import LSTEP, MergeLayer
    node_feats = np.zeros((graph.num_nodes + 1, feat_dim), dtype=np.float32)
    if edge_feats is None:
        edge_feats = np.zeros((graph.num_edges + 1, feat_dim), dtype=np.float32)
    lstep = LSTEP(node_feats, edge_feats, sampler, sampler, pe_dim=pe_dim, num_neighbors=K, time_feat_dim=time_dim,
                  num_fft_batches=T, device=device)
    merge = MergeLayer(feat_dim, feat_dim, feat_dim, 1)
    p = golden_params(tag)
    lstep.load_state_dict({k[2:]: torch.from_numpy(v) for k, v in p.items() if k.startswith("0.")})
    merge.load_state_dict({k[2:]: torch.from_numpy(v) for k, v in p.items() if k.startswith("1.")})
    model = torch.nn.Sequential(lstep, merge).to(device)
    return model


def replay_eval(model, hist0, graph, e0, B, neg_dst, T, K, time_gap, n_batches=None, on_batch=None):
    """Free-running replay with 'random' negatives (negative source = positive source,
    evaluate_model_utils.py:51-52). Returns per-batch (probabilities, labels) and the final table."""
    from sklearn.metrics import average_precision_score, roc_auc_score
    lstep, merge = model[0], model[1]
    model.eval()
    dev = next(merge.parameters()).device
    pe = torch.from_numpy(hist0).to(dev)
    E = graph.num_edges
    nb = (E - e0 + B - 1) // B if n_batches is None else n_batches
    aps, aucs, losses, preds = [], [], [], []
    cur = None
    with torch.no_grad():
        for b in range(nb):
            lo, hi = e0 + b * B, min(e0 + (b + 1) * B, E)
            src, dst = graph.src_node_ids[lo:hi], graph.dst_node_ids[lo:hi]
            tt, ee = graph.node_interact_times[lo:hi], graph.edge_ids[lo:hi]
            nd = np.asarray(neg_dst[b][:hi - lo])
            ids = np.unique(np.concatenate([src, dst]))
            if pe.shape[1] > T:
                pe = pe[:, -T:, :].clone()
            fft_pe = lstep.fourier_transform_pe(ids, pe, b)
            cur = pe[:, -1, :].clone()
            cur[torch.from_numpy(ids).to(dev)] = fft_pe
            kw = dict(node_interact_times=tt, num_neighbors=K, time_gap=time_gap)
            pos_src = lstep.combining_pe_raw_feat(pe=cur, node_ids=src, **kw)
            pos_dst = lstep.combining_pe_raw_feat(pe=cur, node_ids=dst, **kw)
            neg_src = lstep.combining_pe_raw_feat(pe=cur, node_ids=src, **kw)
            neg_dst_e = lstep.combining_pe_raw_feat(pe=cur, node_ids=nd, **kw)
            pp = merge(pos_src, pos_dst).squeeze(-1).sigmoid().clamp(0, 1)
            npb = merge(neg_src, neg_dst_e).squeeze(-1).sigmoid().clamp(0, 1)
            predicts = torch.cat([pp, npb]).cpu().numpy()
            labels = np.concatenate([np.ones(len(pp)), np.zeros(len(npb))])
            ret = lstep.update_pe(pe=cur, node_ids=ids, edge_ids=ee, batch_src_node_ids=src, batch_dst_node_ids=dst,
                                  node_interact_times=tt, current_time=tt.max(), num_neighbors=K, time_gap=time_gap)
            assert ret is cur
            pe = torch.cat([pe, cur.unsqueeze(1)], dim=1)
            preds.append(predicts.astype(np.float32))
            aps.append(average_precision_score(labels, predicts))
            aucs.append(roc_auc_score(labels, predicts))
            pc = np.clip(predicts.astype(np.float64), 1e-12, 1 - 1e-12)
            losses.append(float(-(labels * np.log(pc) + (1 - labels) * np.log(1 - pc)).mean()))
            if on_batch is not None:
                on_batch(b, cur)
    replay_eval.last_predicts = preds
    return np.array(aps), np.array(aucs), np.array(losses), cur


def check_rank_metrics(preds_got, preds_ref, B, ap_got, auc_got, ap_ref, auc_ref, what, log=None, eps_bar=2e-5):
    """AP / AUC 'must match for fixed seeds' (BASELINE.json north_star). Both are functions of the ORDER of positive vs
    negative scores only. Per batch, with eps = max |score_got - score_ref| over the 2B link probabilities and gap = the
    smallest |positive score - negative score| of the reference: when gap > 2 * eps no positive / negative pair can have
    swapped, and AP and AUC must be EQUAL (to 1e-12, sklearn's own arithmetic); otherwise the batch holds a
    positive / negative pair the reference itself separates by less than the fp32 noise (a tie within noise): those
    batches are counted and reported, and may differ by the rank swaps of exactly those pairs.
    preds_ref: [n_batches, 2B] = [pos | neg] zero padded (tests/golden/make_golden.py); preds_got: list of [2b] arrays."""
    n = len(preds_got)
    eps_all, tied, worst_tied = [], 0, 0.0
    for i in range(n):
        g = np.asarray(preds_got[i], np.float64)
        h = len(g) // 2
        ref = np.concatenate([preds_ref[i, :h], preds_ref[i, B:B + h]]).astype(np.float64)
        eps = float(np.abs(g - ref).max())
        eps_all.append(eps)
        gap = float(np.abs(ref[:h, None] - ref[None, h:]).min())
        d_ap, d_auc = abs(ap_got[i] - ap_ref[i]), abs(auc_got[i] - auc_ref[i])
        if gap > 2 * eps:
            assert d_ap < 1e-12 and d_auc < 1e-12, (what, "batch", i, "AP/AUC differ without a tie", d_ap, d_auc, gap, eps)
        else:
            tied += 1
            n_pairs = int((np.abs(ref[:h, None] - ref[None, h:]) <= 2 * eps).sum())
            worst_tied = max(worst_tied, d_ap, d_auc)
            # one swapped pair moves AUC by 1/h^2 and AP by at most ~1/h
            assert d_auc <= n_pairs / (h * h) + 1e-12 and d_ap <= n_pairs / h + 1e-12, (what, "batch", i, d_ap, d_auc, n_pairs)
    rep = {"batches": n, "max_score_err": float(max(eps_all)), "batches_with_pos_neg_tie_within_noise": tied,
           "worst_ap_auc_diff_in_tied_batches": worst_tied,
           "mean_ap_diff": float(abs(np.mean(ap_got) - np.mean(ap_ref))), "mean_auc_diff": float(abs(np.mean(auc_got) - np.mean(auc_ref)))}
    if log is not None:
        log[str(what)] = rep
    assert rep["max_score_err"] <= eps_bar, (what, rep)
    return rep


_F64_REPLAY = {}


def oracle_replay_f64(tag: str):
    """Final PE table of the golden free-running replay `tag` evaluated by the oracle in float64 (the 'exact' recurrence
    on the same inputs), cached per process. The PE recurrence does not depend on the feature branch or the negatives."""
    if tag in _F64_REPLAY:
        return _F64_REPLAY[tag]
    from common import golden_path, seeded_normal
    from lstep_b200 import synth
    from oracle import lstep_oracle as orc
    z = np.load(golden_path(f"replay_{tag}.npz"))
    d, T, K, B = int(z["pe_dim"]), int(z["T"]), int(z["K"]), int(z["B"])
    V, E, e0 = int(z["V"]), int(z["E"]), int(z["e0"])
    g = synth.make_graph("tiny", seed=int(z["graph_seed"]), num_nodes=V, num_edges=E)
    adj = orc.build_adjacency(g.src_node_ids, g.dst_node_ids, g.edge_ids, g.node_interact_times)
    p = lstep_params_np(tag)
    hist = seeded_normal(int(z["hist0_seed"]), (V + 1, 1, d), 0.3).astype(np.float64)
    hist[0] = 0
    with orc.high_precision():
        for b in range(len(z["ap"])):
            lo, hi = e0 + b * B, min(e0 + (b + 1) * B, E)
            hist, _, cur = orc.pe_step(p, adj, hist, b, g.src_node_ids[lo:hi], g.dst_node_ids[lo:hi], g.node_interact_times[lo:hi], [], T, K)
    _F64_REPLAY[tag] = cur
    return cur
