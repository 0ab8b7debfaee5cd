"""Test-side helpers: load golden parameters into the drop-in modules and replay the eval
loop's per-batch sequence (the caller code of evaluate_model_utils.py:38-142, restated here
because the reference cannot be imported on the GPU box)."""
from __future__ import annotations

import numpy as np
import torch

from common import load_params


def lstep_params_np(tag: str) -> dict:
    p = load_params(f"params_{tag}.npz")
    return {k[2:]: v for k, v in p.items() if k.startswith("0.")}


def build_dropin(tag, graph, sampler, feat_dim, pe_dim, time_dim, T, K, edge_feats=None, device="cuda"):
    from lstep_b200 import LSTEP, MergeLayer
    node_feats = np.zeros((graph.num_nodes + 1, feat_dim), dtype=np.float32)
    if edge_feats is None:
        edge_feats = np.zeros((graph.num_edges + 1, feat_dim), dtype=np.float32)
    lstep = LSTEP(node_feats, edge_feats, sampler, sampler, pe_dim=pe_dim, num_neighbors=K, time_feat_dim=time_dim,
                  num_fft_batches=T, device=device)
    merge = MergeLayer(feat_dim, feat_dim, feat_dim, 1)
    p = load_params(f"params_{tag}.npz")
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
    aps, aucs, losses = [], [], []
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
            aps.append(average_precision_score(labels, predicts))
            aucs.append(roc_auc_score(labels, predicts))
            pc = np.clip(predicts.astype(np.float64), 1e-12, 1 - 1e-12)
            losses.append(float(-(labels * np.log(pc) + (1 - labels) * np.log(1 - pc)).mean()))
            if on_batch is not None:
                on_batch(b, cur)
    return np.array(aps), np.array(aucs), np.array(losses), cur
