"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on CPU.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

The reference imports three packages that are not installed offline; `shims/` supplies
`torch_scatter` (scatter == scatter_add_) and `tgb` (import-only stub). Nothing of the
reference is copied: it is imported from where it lies and driven through its public API
(NeighborSampler, LSTEP methods, evaluate_model_link_prediction). Outputs are small fixtures;
large inputs are regenerated from seeds by tests/golden/common.py and pinned by checksums.
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from common import ROOT, checksum, golden_path, seeded_edge_feats, seeded_normal  # noqa: E402

REF = os.environ.get("LSTEP_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(ROOT, "oracle", "shims"))
sys.path.insert(0, REF)

import torch  # noqa: E402
import torch.nn as nn  # noqa: E402

warnings.filterwarnings("ignore")
torch.set_num_threads(os.cpu_count())

from models.LSTEP import LSTEP as RefLSTEP  # noqa: E402
from models.modules import MergeLayer as RefMergeLayer  # noqa: E402
from utils.DataLoader import Data as RefData, get_idx_data_loader  # noqa: E402
from utils.utils import NegativeEdgeSampler, get_neighbor_sampler  # noqa: E402
from evaluate_model_utils import evaluate_model_link_prediction  # noqa: E402

from lstep_b200 import synth  # noqa: E402


def ref_data(g) -> RefData:
    return RefData(g.src_node_ids, g.dst_node_ids, g.node_interact_times, g.edge_ids, g.labels)


PE_MLPS = ["pe_mlp_1", "pe_mlp_2", "self_update_pe", "pe_neighbor_mlp_1", "pe_neighbor_mlp_2", "self_update_neighbor_pe"]


def build_model(g, pe_dim, time_dim, T, K, feat_dim, seed=0, scale=2.0):
    """Reference constructor under torch.manual_seed(seed). `scale` multiplies the six PE-MLP weight
    matrices: 2.0 = the stress regime (tanh exercised off zero, pre-activations up to |z| ~ 50), 1.0 = the torch
    default initialisation as it is (the realistic regime; fixtures tagged `...u`). Scaling by 2 is exact in fp32, so
    the unscaled parameters are params_<tag>.npz with those six matrices halved (tests/harness.py does that)."""
    torch.manual_seed(seed)
    node_feats = np.zeros((g.num_nodes + 1, feat_dim), dtype=np.float32)  # zeros, as the datasets' node feats
    edge_feats = seeded_edge_feats(g.num_edges, feat_dim)
    sampler = get_neighbor_sampler(ref_data(g), "recent", seed=1)
    lstep = RefLSTEP(node_feats, edge_feats, sampler, sampler, pe_dim=pe_dim, num_neighbors=K, time_feat_dim=time_dim,
                     num_fft_batches=T, device="cpu")
    merge = RefMergeLayer(feat_dim, feat_dim, feat_dim, 1)
    model = nn.Sequential(lstep, merge)
    # the torch default init leaves the PE MLPs small; scale them up so tanh() is exercised off zero
    with torch.no_grad():
        for name in PE_MLPS:
            getattr(lstep, name).weight.mul_(scale)
    return model, sampler, node_feats, edge_feats


def state_to_npz(model) -> dict:
    out = {}
    for k, v in model.state_dict().items():
        out[k] = v.detach().cpu().numpy()
    return out


# ---------------------------------------------------------------------------------------------
def gen_sampler():
    """(i) sampler triples incl. ties, empty histories, node 0, N>B truncation, several K."""
    for gname in ["tiny", "tiny_bip", "tiny_ties"]:
        g = synth.make_graph(gname, seed=0)
        sampler = get_neighbor_sampler(ref_data(g), "recent", seed=0)
        rng = np.random.default_rng(7)
        out = dict(src=g.src_node_ids, dst=g.dst_node_ids, eid=g.edge_ids, t=g.node_interact_times,
                   num_nodes=np.int64(g.num_nodes))
        cases = []
        # a) a mid-stream batch queried at its own edge times (what the model does)
        lo = g.num_edges // 2
        q_ids = np.concatenate([g.src_node_ids[lo:lo + 40], g.dst_node_ids[lo:lo + 40]])
        q_t = np.concatenate([g.node_interact_times[lo:lo + 40]] * 2)
        cases.append((q_ids, q_t))
        # b) random nodes (incl. padding node 0 and never-seen ids) at random times, incl. t before any edge
        q_ids = rng.integers(0, int(max(g.src_node_ids.max(), g.dst_node_ids.max())) + 1, size=64)
        q_ids[:3] = 0
        q_t = rng.random(64) * g.node_interact_times.max() * 1.2
        q_t[5:9] = -1.0
        cases.append((q_ids.astype(np.int64), q_t))
        # c) query times exactly equal to stored edge times (strict '<': ties excluded)
        pick = rng.integers(0, g.num_edges, size=48)
        cases.append((g.src_node_ids[pick].copy(), g.node_interact_times[pick].copy()))
        # d) zip truncation both ways (Q1): more ids than times, more times than ids
        ids_u = np.unique(np.concatenate([g.src_node_ids[lo:lo + 12], g.dst_node_ids[lo:lo + 12]]))
        cases.append((ids_u, g.node_interact_times[lo:lo + 12].copy()))
        cases.append((ids_u[:5], g.node_interact_times[lo:lo + 12].copy()))
        # e) empty query
        cases.append((np.zeros(0, np.int64), np.zeros(0, np.float64)))
        n_out = 0
        for ci, (qi, qt) in enumerate(cases):
            out[f"q{ci}_ids"], out[f"q{ci}_t"] = qi, qt
            for K in (1, 5, 20, 70):
                a, b, c = sampler.get_historical_neighbors(qi, qt, K)
                assert a.dtype == np.int64 and b.dtype == np.int64 and c.dtype == np.float32
                out[f"q{ci}_K{K}_nbr"], out[f"q{ci}_K{K}_eid"], out[f"q{ci}_K{K}_t"] = a, b, c
                n_out += 1
        out["num_cases"] = np.int64(len(cases))
        # also the adjacency itself (CSR) so the CSR builder is pinned (a1)
        deg = np.array([len(x) for x in sampler.nodes_neighbor_ids], dtype=np.int64)
        out["csr_indptr"] = np.concatenate([[0], np.cumsum(deg)]).astype(np.int64)
        out["csr_nbr"] = np.concatenate([np.asarray(x, dtype=np.int64) for x in sampler.nodes_neighbor_ids])
        out["csr_eid"] = np.concatenate([np.asarray(x, dtype=np.int64) for x in sampler.nodes_edge_ids])
        out["csr_t"] = np.concatenate([np.asarray(x, dtype=np.float64) for x in sampler.nodes_neighbor_times])
        np.savez_compressed(golden_path(f"sampler_{gname}.npz"), **out)
        print("sampler", gname, n_out, "outputs")


def gen_module_goldens(tag, pe_dim, time_dim, T, K, feat_dim, gname, scale=2.0):
    """(ii) fourier_transform_pe, (iii) compute_neighborhood_pe, (iv) update_pe, plus a training
    step's gradients, all on one seeded model; parameters saved as params_<tag>.npz.
    scale=1.0 (tag `<base>u`): only the parts that depend on the PE-MLP weights, no parameter file."""
    g = synth.make_graph(gname, seed=0)
    model, sampler, node_feats, edge_feats = build_model(g, pe_dim, time_dim, T, K, feat_dim, scale=scale)
    lstep = model[0]
    model.eval()
    mlp_only = scale != 2.0
    if not mlp_only:
        np.savez_compressed(golden_path(f"params_{tag}.npz"), **state_to_npz(model))
    V1 = g.num_nodes + 1
    out = dict(gname=np.array(gname), pe_dim=np.int64(pe_dim), time_dim=np.int64(time_dim), T=np.int64(T), K=np.int64(K),
               feat_dim=np.int64(feat_dim))

    # ---- (ii) DFT filter: history lengths around the mask boundary, several batch_idx (Q5)
    lo = g.num_edges // 2
    ids = synth.unique_batch_nodes(g.src_node_ids[lo:lo + 16], g.dst_node_ids[lo:lo + 16])
    out["dft_ids"] = ids
    cases = [] if mlp_only else [(1, 0), (1, 1), (min(5, T), 3), (min(5, T), 5), (T - 1, T - 1), (T - 1, 2 * T), (T, 0), (T, 7), (T, 5 * T)]
    if T > 40 and not mlp_only:
        cases += [(37, 37), (37, 12)]
    out["dft_cases"] = np.array(cases, dtype=np.int64)
    with torch.no_grad():
        for ci, (Th, bidx) in enumerate(cases):
            hist = seeded_normal(100 + ci, (V1, Th, pe_dim), 0.5)
            out[f"dft{ci}_in_ck"] = checksum(hist)
            y = lstep.fourier_transform_pe(ids, torch.from_numpy(hist), bidx)
            out[f"dft{ci}_out"] = y.numpy()
        if not mlp_only:  # single node -> .squeeze() drops the row dim
            hist = seeded_normal(99, (V1, T, pe_dim), 0.5)
            y = lstep.fourier_transform_pe(ids[:1], torch.from_numpy(hist), 3)
            out["dft_single_out"] = y.numpy()

    # ---- (iii) neighbourhood aggregate with nonzero pe[0] (Q2)
    pe = seeded_normal(5, (V1, pe_dim), 0.3)
    assert np.abs(pe[0]).sum() > 0
    out["pe_in_ck"] = checksum(pe)
    q_ids = np.concatenate([g.src_node_ids[lo:lo + 24], g.dst_node_ids[lo:lo + 24], g.src_node_ids[10:14]])
    q_t = np.concatenate([g.node_interact_times[lo:lo + 24]] * 2 + [g.node_interact_times[10:14]])
    out["nbr_q_ids"], out["nbr_q_t"] = q_ids, q_t
    with torch.no_grad():
        for KK in (K, 3):
            y = lstep.compute_neighborhood_pe(torch.from_numpy(pe.copy()), q_ids, q_t, num_neighbors=KK)
            out[f"nbr_out_K{KK}"] = y.numpy()

    # ---- (iv) update_pe: teacher-forced single steps; B edges, N unique nodes; incl. N > B (Q1/Q1b)
    upd_cases = [(lo, 16), (lo + 100, 1), (g.num_edges - 40, 40), (3, 8)]
    out["upd_cases"] = np.array(upd_cases, dtype=np.int64)
    with torch.no_grad():
        for ci, (s, B) in enumerate(upd_cases):
            src, dst = g.src_node_ids[s:s + B], g.dst_node_ids[s:s + B]
            tt, ee = g.node_interact_times[s:s + B], g.edge_ids[s:s + B]
            nids = synth.unique_batch_nodes(src, dst)
            pe_t = torch.from_numpy(seeded_normal(40 + ci, (V1, pe_dim), 0.3))
            ret = lstep.update_pe(pe_t, nids, ee, src, dst, tt, tt.max(), num_neighbors=K)
            assert ret is pe_t  # Q7
            out[f"upd{ci}_out"] = pe_t.numpy().copy()
            out[f"upd{ci}_N"] = np.int64(len(nids))
        # caller passes a subset of the batch nodes: contributions to other nodes are dropped
        s, B = upd_cases[0]
        src, dst = g.src_node_ids[s:s + B], g.dst_node_ids[s:s + B]
        tt, ee = g.node_interact_times[s:s + B], g.edge_ids[s:s + B]
        nids = synth.unique_batch_nodes(src, dst)[::2].copy()
        pe_t = torch.from_numpy(seeded_normal(49, (V1, pe_dim), 0.3))
        lstep.update_pe(pe_t, nids, ee, src, dst, tt, tt.max(), num_neighbors=K)
        out["upd_subset_ids"] = nids
        out["upd_subset_out"] = pe_t.numpy().copy()

    if mlp_only:
        np.savez_compressed(golden_path(f"module_{tag}.npz"), **out)
        print("module", tag, "done")
        return
    # ---- training step: gradients that reach the PE path
    model.train()
    hist = torch.from_numpy(seeded_normal(77, (V1, T, pe_dim), 0.5))
    src, dst, tt = g.src_node_ids[lo:lo + 16], g.dst_node_ids[lo:lo + 16], g.node_interact_times[lo:lo + 16]
    nids = synth.unique_batch_nodes(src, dst)
    fft_pe = lstep.fourier_transform_pe(nids, hist, 2 * T)
    cur = torch.clone(hist[:, -1, :])
    cur[torch.from_numpy(nids)] = fft_pe
    a = lstep.compute_neighborhood_pe(cur, src, tt, num_neighbors=K)
    b = lstep.compute_neighborhood_pe(cur, dst, tt, num_neighbors=K)
    loss = (a * b).sum() + (cur[torch.from_numpy(src)] - cur[torch.from_numpy(dst)]).pow(2).mean()
    loss.backward()
    out["train_loss"] = np.float64(loss.item())
    for name in ["fft_filter.weight", "fft_agg.weight", "pe_neighbor_mlp_1.weight", "pe_neighbor_mlp_1.bias",
                 "pe_neighbor_mlp_2.weight", "self_update_neighbor_pe.weight"]:
        gr = dict(lstep.named_parameters())[name].grad
        out["grad_" + name] = gr.detach().numpy()
    for name in ["self_update_pe.weight", "pe_mlp_1.weight", "pe_mlp_2.weight"]:
        assert dict(lstep.named_parameters())[name].grad is None  # update_pe is forward-only
    np.savez_compressed(golden_path(f"module_{tag}.npz"), **out)
    print("module", tag, "done")


def gen_replay(tag, pe_dim, time_dim, T, K, feat_dim, time_gap, V, E, B, n_eval_batches, scale=2.0, params_tag=None):
    """(v) free-running evaluate_model_link_prediction replay: per-batch AP/AUC/loss, the 2B link
    probabilities themselves and PE checksums. Hooks record what the untouched loop feeds the model."""
    g = synth.make_graph("tiny", seed=3, num_nodes=V, num_edges=E)
    model, sampler, node_feats, edge_feats = build_model(g, pe_dim, time_dim, T, K, feat_dim, scale=scale)
    lstep = model[0]
    e0 = E - n_eval_batches * B
    ev = g.slice(e0, E)
    eval_data = ref_data(ev)
    neg = NegativeEdgeSampler(g.src_node_ids, g.dst_node_ids, seed=2)
    loader = get_idx_data_loader(list(range(ev.num_edges)), batch_size=B, shuffle=False)
    hist0 = seeded_normal(11, (V + 1, 1, pe_dim), 0.3)
    hist0[0] = 0

    rec = dict(neg_dst=[], pe_ck=[], predicts=[])
    orig_sample = neg.sample
    orig_update = lstep.update_pe

    def sample_hook(*a, **k):
        r = orig_sample(*a, **k)
        rec["neg_dst"].append(np.asarray(r[1]).copy())
        return r

    def update_hook(*a, **k):
        r = orig_update(*a, **k)
        rec["pe_ck"].append(checksum(r.detach().numpy()))
        rec["last_pe"] = r.detach().numpy().copy()
        return r

    neg.sample = sample_hook
    lstep.update_pe = update_hook
    import evaluate_model_utils as emu
    orig_metrics = emu.get_link_prediction_metrics

    def metrics_hook(predicts, labels):
        rec["predicts"].append(predicts.detach().numpy().astype(np.float32).copy())
        return orig_metrics(predicts=predicts, labels=labels)

    emu.get_link_prediction_metrics = metrics_hook
    emu.tqdm = lambda it, **kw: type("Q", (), {"__iter__": lambda s: iter(it), "set_description": lambda s, *_: None})()
    losses, metrics = evaluate_model_link_prediction("LSTEP", model, torch.from_numpy(hist0), sampler, loader, neg, eval_data,
                                                     nn.BCELoss(), num_fft_batches=T, num_neighbors=K, time_gap=time_gap)
    emu.get_link_prediction_metrics = orig_metrics
    pred = np.zeros((len(rec["predicts"]), 2 * B), np.float32)  # [pos | neg] per batch; the ragged last batch is zero padded
    for i, p_ in enumerate(rec["predicts"]):
        h = len(p_) // 2
        pred[i, :h], pred[i, B:B + h] = p_[:h], p_[h:]
    out = dict(predicts=pred, mlp_scale=np.float64(scale),pe_dim=np.int64(pe_dim), time_dim=np.int64(time_dim), T=np.int64(T), K=np.int64(K), feat_dim=np.int64(feat_dim),
               time_gap=np.int64(time_gap), V=np.int64(V), E=np.int64(E), B=np.int64(B), e0=np.int64(e0),
               graph_seed=np.int64(3), hist0_seed=np.int64(11),
               losses=np.array(losses), ap=np.array([m["average_precision"] for m in metrics]),
               auc=np.array([m["roc_auc"] for m in metrics]),
               neg_dst=np.stack(rec["neg_dst"][:-1] + [np.resize(rec["neg_dst"][-1], B)]) if len(rec["neg_dst"]) else np.zeros(0),
               neg_last_len=np.int64(len(rec["neg_dst"][-1])),
               pe_ck=np.stack(rec["pe_ck"]), last_pe=rec["last_pe"],
               graph_ck=checksum(g.node_interact_times), edge_feats_ck=checksum(edge_feats))
    np.savez_compressed(golden_path(f"replay_{tag}.npz"), **out)
    # parameters: identical to params_<tag>.npz (same constructor under the same torch seed)
    ref_params = np.load(golden_path(f"params_{params_tag or tag}.npz"))
    for k, v in state_to_npz(model).items():
        is_mlp = k.startswith("0.") and k[2:].rsplit(".", 1)[0] in PE_MLPS and k.endswith(".weight")
        assert np.array_equal(ref_params[k] * np.float32(scale / 2.0 if is_mlp else 1.0), v), k
    print("replay", tag, "AP", float(np.mean(out["ap"])), "AUC", float(np.mean(out["auc"])), "batches", len(losses))


def gen_run_bracket():
    """(f4) the reference's PE initialisation on a first-batch graph (train_LSTEP_link_prediction.py:168-189) through the
    torch_geometric import shim, and a PE-history file written by the reference's own EarlyStopping.save_pe."""
    import logging
    import shutil
    import tempfile
    from utils.EarlyStopping import EarlyStopping
    from utils.PositionalEncoding import LaplacianPE, RandomWalkPE
    g = synth.make_graph("tiny", seed=3, num_nodes=300, num_edges=14000)
    B = 50
    src, dst = g.src_node_ids[:B], g.dst_node_ids[:B]
    edge_index = torch.from_numpy(np.array([src.tolist() + dst.tolist(), dst.tolist() + src.tolist()]))
    num_nodes = g.num_nodes + 1
    out = dict(edge_index=edge_index.numpy(), num_nodes=np.int64(num_nodes))
    out["rwpe_12"] = RandomWalkPE(edge_index, num_nodes, 12).numpy()
    out["rwpe_40"] = RandomWalkPE(edge_index, num_nodes, 40).numpy()
    torch.manual_seed(0)
    pe, ew = LaplacianPE(edge_index, num_nodes, 12)
    out["lappe_12"], out["lappe_edge_weight"] = pe.numpy(), ew.numpy()
    torch.manual_seed(0)
    out["lappe_sign"] = (-1 + 2 * torch.randint(0, 2, (12,))).numpy()
    np.savez_compressed(golden_path("run_bracket.npz"), **out)
    tmp = tempfile.mkdtemp()
    try:
        es = EarlyStopping(patience=1, save_model_folder=tmp, save_model_name="m", logger=logging.getLogger("golden"),
                           save_trained_pe="pe", save_spatial_ne="ne", model_name="LSTEP")
        hist = torch.from_numpy(seeded_normal(21, (61, 8, 12), 0.3))
        es.save_pe(hist)
        shutil.copyfile(es.save_trained_positional_encoding_path, golden_path("ref_saved_pe.pkl"))
        assert torch.equal(es.load_pe(), hist)
    finally:
        shutil.rmtree(tmp)
    print("run bracket done")


def gen_feature_branch():
    """(f2) LSTEP.aggregated_node_embeddings / combining_pe_raw_feat of the reference with non-zero edge features, once with the
    datasets' all-zero node features and once with non-zero ones (the time_gap masked mean then matters)."""
    g = synth.make_graph("tiny_bip", seed=0)
    V1 = g.num_nodes + 1
    lo = g.num_edges // 2
    q_ids = np.concatenate([g.src_node_ids[lo:lo + 30], g.dst_node_ids[lo:lo + 30], g.src_node_ids[5:9], np.zeros(2, np.int64)])
    q_t = np.concatenate([g.node_interact_times[lo:lo + 30]] * 2 + [g.node_interact_times[5:9], g.node_interact_times[lo:lo + 2]])
    out = dict(q_ids=q_ids, q_t=q_t)
    for name, nf_seed in (("zero_nf", None), ("rand_nf", 9)):
        model, sampler, node_feats, edge_feats = build_model(g, 172, 100, 100, 20, 172)
        lstep = model[0].eval()
        if nf_seed is not None:
            nf = seeded_normal(nf_seed, (V1, 172), 1.0)
            nf[0] = 0
            lstep.node_raw_features = torch.from_numpy(nf)
        pe = torch.from_numpy(seeded_normal(5, (V1, 172), 0.3))
        with torch.no_grad():
            out[f"{name}_emb"] = lstep.aggregated_node_embeddings(q_ids, q_t, num_neighbors=20, time_gap=50).numpy()
            out[f"{name}_comb"] = lstep.combining_pe_raw_feat(pe, q_ids, q_t, num_neighbors=20, time_gap=50).numpy()
    np.savez_compressed(golden_path("feature_branch.npz"), **out)
    print("feature branch done")


def negative_batches(E, B):
    """Batch starts used by the negative-sampler fixture: the first edges of the stream (few historical edges: the fall-back
    to random_sample_with_collision_check runs) and the evaluation tail."""
    return list(range(0, 6 * B, B)) + list(range(E - 120 * B, E, B))


def gen_negatives():
    """(f3) per-batch negatives of the reference's NegativeEdgeSampler for all three strategies, two passes each (the second
    after reset_random_state, as every evaluation run starts: evaluate_model_utils.py:25-26)."""
    g = synth.make_graph("tiny_bip", seed=3, num_nodes=300, num_edges=14000)
    E, B = g.num_edges, 50
    out = dict(B=np.int64(B), graph_ck=checksum(g.node_interact_times), starts=np.array(negative_batches(E, B), dtype=np.int64))
    for strat in ("random", "historical", "inductive"):
        s = NegativeEdgeSampler(g.src_node_ids, g.dst_node_ids, interact_times=g.node_interact_times,
                                last_observed_time=g.node_interact_times[int(E * 0.7)], negative_sample_strategy=strat, seed=2)
        for rep in range(2):
            s.reset_random_state()
            ns, nd = [], []
            for lo in negative_batches(E, B):
                hi = lo + B
                if strat == "random":
                    a, b = s.sample(size=B)
                else:
                    a, b = s.sample(size=B, batch_src_node_ids=g.src_node_ids[lo:hi], batch_dst_node_ids=g.dst_node_ids[lo:hi],
                                    current_batch_start_time=g.node_interact_times[lo], current_batch_end_time=g.node_interact_times[hi - 1])
                assert a.dtype == np.int64 and b.dtype == np.int64 and len(a) == B
                ns.append(a)
                nd.append(b)
            out[f"{strat}_{rep}_src"], out[f"{strat}_{rep}_dst"] = np.stack(ns).astype(np.int32), np.stack(nd).astype(np.int32)
    np.savez_compressed(golden_path("negatives.npz"), **out)
    print("negatives done")


if __name__ == "__main__":
    which = sys.argv[1:] or ["sampler", "module", "replay", "bracket", "negatives", "feature"]
    if "feature" in which:
        gen_feature_branch()
    if "negatives" in which:
        gen_negatives()
    if "bracket" in which:
        gen_run_bracket()
    if "sampler" in which:
        gen_sampler()
    if "module" in which:
        gen_module_goldens("small", pe_dim=12, time_dim=10, T=8, K=4, feat_dim=12, gname="tiny")
        gen_module_goldens("full", pe_dim=172, time_dim=100, T=100, K=20, feat_dim=172, gname="tiny_bip")
        gen_module_goldens("fullu", pe_dim=172, time_dim=100, T=100, K=20, feat_dim=172, gname="tiny_bip", scale=1.0)
    if "replay" in which:
        gen_replay("small", pe_dim=12, time_dim=10, T=8, K=4, feat_dim=12, time_gap=50, V=60, E=1500, B=10, n_eval_batches=60)
        gen_replay("full", pe_dim=172, time_dim=100, T=100, K=20, feat_dim=172, time_gap=2000, V=300, E=14000, B=50,
                   n_eval_batches=230)
        gen_replay("fullu", pe_dim=172, time_dim=100, T=100, K=20, feat_dim=172, time_gap=2000, V=300, E=14000, B=50,
                   n_eval_batches=230, scale=1.0, params_tag="full")
