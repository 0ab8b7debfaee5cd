"""Drop-in for the reference's LSTEP module (models/LSTEP.py:28-341).

Same constructor, method names, argument meaning, return values and state_dict keys, so the
reference's train / evaluate loops run unchanged with `from lstep_b200 import LSTEP`. The three
positional-encoding methods on the hot path run as sm_100a CUDA kernels through the C ABI:

    fourier_transform_pe      -> lstep_dft_filter        (K3; csrc/dft_filter.cu)
    compute_neighborhood_pe   -> lstep_sample_recent_compact + lstep_neighborhood_pe (K1 + K2 + MLP)
    update_pe                 -> lstep_update_pe         (K1 + K2 scatter side + K4 ordered write-back)

The feature branch (aggregated_node_embeddings): the K-neighbour edge-feature / time mixer runs as one fused
CUDA kernel (lstep_feature_aggregate: lookup + weighted gather + time features, with edge_mlp_1 / edge_agg collapsed by
linearity) followed by the two small Linear layers; the masked mean over `time_gap` neighbours is identically zero for
all-zero node features (every dataset of the reference) and runs as torch ops on the device lookup otherwise.

There is no CPU path: PE tensors must live on the CUDA device the module was moved to.
When autograd is recording (training), compute_neighborhood_pe and fourier_transform_pe use
CUDA kernels wrapped in autograd.Functions for the gather / filter stages and PyTorch fp32
Linear layers for the small MLPs, so gradients reach fft_filter, fft_agg and the neighbour MLP
parameters exactly as in the reference; update_pe is forward-only there too.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from ._staging import Stager
from .modules import TimeEncoder
from .sampler import NeighborSampler

I64, F64 = np.dtype(np.int64), np.dtype(np.float64)


# ---------------------------------------------------------------------------------------------
# autograd wrappers (training path)
# ---------------------------------------------------------------------------------------------
class _DftFilterFn(torch.autograd.Function):
    """out[n,c] = sum_s G[s,c] * hist[ids[n],s,c]; gradient w.r.t. G only (the history is data)."""

    @staticmethod
    def forward(ctx, G, hist, ids_dev, Th):
        lib = _lib.load()
        n, d = ids_dev.shape[0], hist.shape[2]
        out = torch.empty((n, d), dtype=torch.float32, device=hist.device)
        Gc = G.detach().contiguous()
        _lib.check(lib.lstep_dft_filter(_lib.ptr(hist), hist.stride(0), hist.stride(1), 0, max(Th, 1), Th, d, _lib.ptr(ids_dev), n,
                                        _lib.ptr(Gc), _lib.ptr(out), d, _lib.stream_ptr()), "lstep_dft_filter")
        ctx.save_for_backward(hist, ids_dev)
        ctx.Th, ctx.T = Th, G.shape[0]
        return out

    @staticmethod
    def backward(ctx, dout):
        hist, ids_dev = ctx.saved_tensors
        lib = _lib.load()
        d = hist.shape[2]
        dG = torch.zeros((ctx.T, d), dtype=torch.float32, device=hist.device)
        dout = dout.contiguous()
        _lib.check(lib.lstep_dft_filter_bwd(_lib.ptr(hist), hist.stride(0), hist.stride(1), 0, max(ctx.Th, 1), ctx.Th, d,
                                            _lib.ptr(ids_dev), ids_dev.shape[0], _lib.ptr(dout), _lib.ptr(dG), _lib.stream_ptr()),
                   "lstep_dft_filter_bwd")
        return dG, None, None, None


class _NbrAggregateFn(torch.autograd.Function):
    """S[i] = sum_k [pe[nbr[i,k]] || tf]; gradient: dpe[nbr[i,k]] += dS[i,:d]."""

    @staticmethod
    def forward(ctx, pe, nbr, nbr_t, q_time, tw, t_dim):
        lib = _lib.load()
        n, K = nbr.shape
        d = pe.shape[1]
        S = torch.empty((n, d + t_dim), dtype=torch.float32, device=pe.device)
        pec = pe.detach().contiguous()
        _lib.check(lib.lstep_nbr_aggregate(_lib.ptr(pec), pec.shape[0], _lib.ptr(q_time), _lib.ptr(nbr), _lib.ptr(nbr_t), n, K,
                                           _lib.ptr(tw), d, t_dim, _lib.ptr(S), _lib.stream_ptr()), "lstep_nbr_aggregate")
        ctx.save_for_backward(nbr)
        ctx.shape = (pe.shape[0], d, t_dim)
        return S

    @staticmethod
    def backward(ctx, dS):
        (nbr,) = ctx.saved_tensors
        rows, d, t_dim = ctx.shape
        lib = _lib.load()
        dpe = torch.zeros((rows, d), dtype=torch.float32, device=dS.device)
        dS = dS.contiguous()
        _lib.check(lib.lstep_nbr_aggregate_bwd(_lib.ptr(dS), _lib.ptr(nbr), nbr.shape[0], nbr.shape[1], d, t_dim, _lib.ptr(dpe),
                                               rows, _lib.stream_ptr()), "lstep_nbr_aggregate_bwd")
        return dpe, None, None, None, None, None


# ---------------------------------------------------------------------------------------------
class LSTEP(nn.Module):
    def __init__(self, node_raw_features: np.ndarray, edge_raw_features: np.ndarray, neighbor_sampler, full_neighbor_sampler,
                 pe_dim: int, num_neighbors: int, time_feat_dim: int, num_fft_batches: int, use_dropout=False, dropout: float = 0.1,
                 weighted_sum=False, concat_pe=True, device: str = 'cuda'):
        super().__init__()
        edge_feat_dim = edge_raw_features.shape[-1]
        node_feat_dim = node_raw_features.shape[-1]
        self.num_fft_batches = num_fft_batches
        self.num_nodes = node_raw_features.shape[0]
        self.pe_dim = pe_dim
        self.time_feat_dim = time_feat_dim
        self.use_dropout = use_dropout
        self.dropout = dropout
        self.concat_pe = concat_pe
        self.weighted_sum = weighted_sum
        self.device = device

        self.node_raw_features = torch.from_numpy(node_raw_features.astype(np.float32)).to(device)
        self.edge_raw_features = torch.from_numpy(edge_raw_features.astype(np.float32)).to(device)
        self._node_feats_all_zero = not bool(np.any(node_raw_features))

        self.neighbor_sampler = self._adopt(neighbor_sampler)
        self.full_neighbor_sampler = self._adopt(full_neighbor_sampler)

        # parameter names = the reference's state_dict keys (models/LSTEP.py:50-72)
        self.time_encoder = TimeEncoder(time_feat_dim, parameter_requires_grad=False)
        self.fft_filter = nn.Linear(pe_dim, num_fft_batches, bias=False).to(torch.complex64)  # used as a [T,d] table
        self.fft_dropout = nn.Dropout(p=dropout)
        self.fft_agg = nn.Linear(num_fft_batches, 1, bias=False)
        self.edge_mlp_1 = nn.Linear(edge_feat_dim + time_feat_dim, edge_feat_dim + time_feat_dim)
        self.edge_agg = nn.Linear(num_neighbors, 1)
        self.edge_mlp_2 = nn.Linear(edge_feat_dim + time_feat_dim, edge_feat_dim + time_feat_dim)
        self.node_mlp = nn.Linear(edge_feat_dim + node_feat_dim + time_feat_dim, node_feat_dim)
        self.self_update_pe = nn.Linear(pe_dim, pe_dim)
        self.pe_mlp_1 = nn.Linear(pe_dim + time_feat_dim, pe_dim)
        self.pe_mlp_2 = nn.Linear(pe_dim, pe_dim)
        self.self_update_neighbor_pe = nn.Linear(pe_dim, pe_dim)
        self.pe_neighbor_mlp_1 = nn.Linear(pe_dim + time_feat_dim, pe_dim)
        self.pe_neighbor_mlp_2 = nn.Linear(pe_dim, pe_dim)
        self.out_node_emb = nn.Linear(pe_dim + node_feat_dim, node_feat_dim)

        self._pack_cache = {}
        self._pack_fast = {}
        self._pack_calls = 0
        self._filter_cache = {}
        self._twiddle = {}
        self._stager = None
        self._update_ws = None
        self.h2d_bytes = 0  # bytes of per-call host inputs moved to the device (bench accounting)

    # ---- plumbing ---------------------------------------------------------------------------
    @staticmethod
    def _adopt(sampler):
        """Accept our device sampler, or an instance of the reference's host sampler."""
        if sampler is None or isinstance(sampler, NeighborSampler):
            return sampler
        if hasattr(sampler, "nodes_neighbor_ids"):
            return NeighborSampler.from_reference(sampler)
        return sampler

    def set_neighbor_sampler(self, neighbor_sampler):
        """models/LSTEP.py:76-85"""
        self.neighbor_sampler = self._adopt(neighbor_sampler)
        if self.neighbor_sampler.sample_neighbor_strategy in ['uniform', 'time_interval_aware']:
            assert self.neighbor_sampler.seed is not None
            self.neighbor_sampler.reset_random_state()

    def _dev(self) -> torch.device:
        return self.pe_mlp_1.weight.device

    def _require_cuda(self, x: torch.Tensor, what: str):
        dev = self._dev()
        if dev.type != "cuda":
            raise _lib.LstepError("LSTEP parameters are on the CPU; lstep_b200 has no CPU path — move the module to a CUDA device")
        if not isinstance(x, torch.Tensor) or x.device != dev:
            raise _lib.LstepError(f"{what} must be a tensor on {dev} (got {getattr(x, 'device', type(x))})")
        if x.dtype != torch.float32:
            raise _lib.LstepError(f"{what} must be float32")

    def _upload(self, arrays):
        if self._stager is None or self._stager.device != self._dev():
            self._stager = Stager(self._dev())
        before = self._stager.bytes_moved
        out = self._stager.upload(arrays)
        self.h2d_bytes += self._stager.bytes_moved - before
        return out

    @staticmethod
    def _check_ids(ids: np.ndarray, limit: int, what: str):
        if len(ids) and (int(ids.min()) < 0 or int(ids.max()) >= limit):
            raise IndexError(f"{what}: index out of range (table has {limit} rows)")

    def _packed_mlp(self, which: str) -> "_lib.PEMLP":
        names = self._MLP_NAMES[which]
        params = self._mlp_params(which)
        key = tuple((p.data_ptr(), p._version) for p in params)
        hit = self._pack_cache.get(which)
        if hit is not None and hit[0] == key:
            return hit[1]
        lib = _lib.load()
        dev = self._dev()
        d, t = self.pe_dim, self.time_feat_dim
        ldo = lib.lstep_packed_ld(d)
        if bool(torch.count_nonzero(self.time_encoder.w.bias)):
            raise _lib.LstepError("TimeEncoder bias must be zero (it is frozen at zero in the reference, models/modules.py:21)")
        bufs, tc = [], []
        with torch.cuda.device(dev):
            for n, in_f in zip(names, (d + t, d, d)):
                lin = getattr(self, n)
                pw = torch.empty((lib.lstep_packed_rows(in_f), ldo), dtype=torch.float32, device=dev)
                pb = torch.empty(ldo, dtype=torch.float32, device=dev)
                w = lin.weight.detach().contiguous()
                b = lin.bias.detach().contiguous()
                _lib.check(lib.lstep_pack_linear(_lib.ptr(w), _lib.ptr(b), d, in_f, _lib.ptr(pw), _lib.ptr(pb), _lib.stream_ptr()),
                           "lstep_pack_linear")
                bufs += [pw, pb]
                pt = torch.empty(lib.lstep_packed_tc_floats(d, in_f), dtype=torch.float32, device=dev)
                _lib.check(lib.lstep_pack_linear_tc(_lib.ptr(w), d, in_f, _lib.ptr(pt), _lib.stream_ptr()), "lstep_pack_linear_tc")
                tc.append(pt)
        tw = self.time_encoder.w.weight.detach().reshape(-1).contiguous()
        bufs.append(tw)
        st = _lib.PEMLP(*(b.data_ptr() for b in bufs), d, t, *(x.data_ptr() for x in tc))
        entry = (key, st, bufs + tc, C.byref(st))
        self._pack_cache[which] = entry
        return st

    _MLP_NAMES = {"update": ("pe_mlp_1", "pe_mlp_2", "self_update_pe"),
                  "nbr": ("pe_neighbor_mlp_1", "pe_neighbor_mlp_2", "self_update_neighbor_pe")}

    def invalidate_packed_weights(self):
        """Drop the packed copies of the PE-MLP weights and the collapsed DFT filters. Needed only after an in-place edit
        THROUGH `.data` (p.data.mul_(...), p.data.copy_(...)), which changes neither the tensor's version counter nor
        its address; every other change (optimizer steps, load_state_dict, .to(), replacing a submodule or a Parameter)
        is detected on the next call."""
        self._pack_cache.clear()
        self._pack_fast.clear()
        self._filter_cache.clear()

    def _mlp_params(self, which: str):
        # the LIVE Parameter objects, fetched without nn.Module.__getattr__'s fallback chain
        mods = self._modules
        out = []
        for n in self._MLP_NAMES[which]:
            p = mods[n]._parameters
            out += [p["weight"], p["bias"]]
        out.append(mods["time_encoder"]._modules["w"]._parameters["weight"])
        return out

    def _mlp_ref(self, which: str):
        # fast path (two calls per step on the streaming API): a packed copy is reused while every live parameter is the
        # same object at the same address with the same version counter — seven identity / integer compares per call, so
        # .to(), .data = assignment, load_state_dict, in-place optimizer steps and replaced submodules are all seen at once
        # (ADVICE r1). In-place edits through .data change none of those: call invalidate_packed_weights().
        fast = self._pack_fast.get(which)
        if fast is not None:
            params, sig, ref = fast
            live = self._mlp_params(which)
            if all(a is b for a, b in zip(live, params)) and tuple((p.data_ptr(), p._version) for p in live) == sig:
                return ref
        self._packed_mlp(which)
        entry = self._pack_cache[which]
        params = self._mlp_params(which)
        self._pack_fast[which] = (params, tuple((p.data_ptr(), p._version) for p in params), entry[3])
        return entry[3]

    # ---- a3: DFT filter ---------------------------------------------------------------------
    def _twiddles(self, T: int, dev):
        key = (T, dev)
        if key not in self._twiddle:
            idx = torch.arange(T, device=dev)
            ang = ((idx[:, None] * idx[None, :]) % T).to(torch.float64) * (2.0 * math.pi / T)
            self._twiddle[key] = torch.polar(torch.ones_like(ang), ang)  # E[f,t] = exp(+2*pi*i*f*t/T)
        return self._twiddle[key]

    def _collapsed_filter_autograd(self, b: int, residual: bool) -> torch.Tensor:
        """G[T,d] as differentiable torch ops (complex128), for the training path."""
        T = self.num_fft_batches
        dev = self._dev()
        E = self._twiddles(T, dev)
        a = self.fft_agg.weight.reshape(-1).to(torch.float64)
        m = (torch.arange(T, device=dev) < b).to(torch.float64)
        A = E @ (a * m).to(torch.complex128)  # A[f] = sum_{t<b} a[t] e^{+2 pi i f t / T}
        Bc = self.fft_filter.weight.to(torch.complex128) * (A * m)[:, None]
        G = (E.conj() @ Bc).real / T
        if residual:
            G = G + a[:, None]
        return G.to(torch.float32)

    def _parameters_ref(self, name: str):
        """The weight Parameter of submodule `name` without nn.Module.__getattr__'s fallback chain."""
        return self._modules[name]._parameters["weight"]

    def _collapsed_filter(self, b: int, residual: bool) -> torch.Tensor:
        """G[T,d] from the collapse kernel, cached until the parameters change."""
        hit = self._filter_cache.get((b, residual))
        if hit is not None:  # (parameters fetched through the cached references: nn.Module.__getattr__ is slow)
            Wc, ac = hit[2], hit[3]
            if (Wc.data_ptr(), Wc._version, ac.data_ptr(), ac._version, b, residual) == hit[0] and Wc is self._parameters_ref("fft_filter") \
                    and ac is self._parameters_ref("fft_agg"):
                return hit[1]
        W, a = self.fft_filter.weight, self.fft_agg.weight
        key = (W.data_ptr(), W._version, a.data_ptr(), a._version, b, residual)
        lib = _lib.load()
        T, d = self.num_fft_batches, self.pe_dim
        G = torch.empty((T, d), dtype=torch.float32, device=self._dev())
        Wr = torch.view_as_real(W.detach()).contiguous()
        ac = a.detach().reshape(-1).contiguous()
        _lib.check(lib.lstep_dft_collapse(_lib.ptr(Wr), _lib.ptr(ac), T, d, b, _lib.ptr(G), _lib.stream_ptr()), "lstep_dft_collapse")
        if residual:
            G += ac[:, None]
        if len(self._filter_cache) > 8:
            self._filter_cache.clear()
        self._filter_cache[(b, residual)] = (key, G, W, a)
        return G

    def fourier_transform_pe(self, node_ids, pe, batch_idx, use_dropout=False, use_mixer=False):
        """models/LSTEP.py:104-137. pe: [V1, Th, d] history, Th <= num_fft_batches. Returns [N, d]
        (squeezed like the reference)."""
        self._require_cuda(pe, "pe history")
        T, d = self.num_fft_batches, self.pe_dim
        if pe.dim() != 3 or pe.shape[2] != d:
            raise RuntimeError(f"pe history must be [num_nodes, T', {d}], got {tuple(pe.shape)}")
        Th = pe.shape[1]
        if Th > T:  # the [T,d] filter cannot broadcast against a longer history (LSTEP.py:121)
            raise RuntimeError(f"The size of tensor a ({T}) must match the size of tensor b ({Th}) at non-singleton dimension 1")
        if use_dropout and self.training:
            raise NotImplementedError("fft dropout in training mode is not collapsible; the reference's callers never enable it")
        if pe.stride(2) != 1:
            pe = pe.contiguous()
        node_ids = np.asarray(node_ids)
        self._check_ids(node_ids, pe.shape[0], "fourier_transform_pe")
        if Th < T:  # mask[:, :batch_idx] (python slice semantics), keyed on batch_idx (Q5)
            b = int(batch_idx)
            b = max(T + b, 0) if b < 0 else min(b, T)
        else:
            b = T
        (ids_dev,) = self._upload([(node_ids, I64)])
        return self.fourier_transform_pe_device(ids_dev, pe, b, bool(use_dropout)).squeeze()

    def fourier_transform_pe_device(self, ids_dev, pe, b: int, residual: bool = False, s0: int = 0, ring: int = None, Th: int = None,
                                    node_stride: int = None, time_stride: int = None, out=None):
        """Device-resident form: ids already in HBM, mask length b resolved. With `ring`/`s0`/strides the
        history may be a ring buffer (logical step s lives in slot (s0+s) % ring)."""
        T, d = self.num_fft_batches, self.pe_dim
        Th = pe.shape[1] if Th is None else Th
        ring = max(Th, 1) if ring is None else ring
        node_stride = pe.stride(0) if node_stride is None else node_stride
        time_stride = pe.stride(1) if time_stride is None else time_stride
        n = ids_dev.shape[0]
        lib = _lib.load()
        with torch.cuda.device(pe.device):
            needs_grad = torch.is_grad_enabled() and (self.fft_filter.weight.requires_grad or self.fft_agg.weight.requires_grad)
            if needs_grad:
                G = self._collapsed_filter_autograd(b, residual)
                out = _DftFilterFn.apply(G, pe.detach(), ids_dev, Th)
            else:
                G = self._collapsed_filter(b, residual)
                if out is None:
                    out = torch.empty((n, d), dtype=torch.float32, device=pe.device)
                _lib.check(lib.lstep_dft_filter(_lib.ptr(pe), node_stride, time_stride, s0, ring, Th, d, _lib.ptr(ids_dev), n,
                                                _lib.ptr(G), _lib.ptr(out), out.stride(0), _lib.stream_ptr()), "lstep_dft_filter")
        return out
    # ---- feature branch (PyTorch; lookups on the device sampler) ------------------------------
    def _sample_full(self, ids_dev, t_dev, n_rows, n_valid, K):
        s = self.neighbor_sampler
        dev = ids_dev.device
        out_n = torch.empty((n_rows, K), dtype=torch.int64, device=dev)
        out_e = torch.empty((n_rows, K), dtype=torch.int64, device=dev)
        out_t = torch.empty((n_rows, K), dtype=torch.float32, device=dev)
        _lib.check(_lib.load().lstep_sample_recent(s.csr_ref, _lib.ptr(ids_dev), _lib.ptr(t_dev), n_rows, n_valid, K, _lib.ptr(out_n),
                                                   _lib.ptr(out_e), _lib.ptr(out_t), _lib.ptr(s._err), _lib.stream_ptr()),
                   "lstep_sample_recent")
        return out_n, out_e, out_t

    def aggregated_node_embeddings(self, node_ids: np.ndarray, node_interact_times: np.ndarray, num_neighbors: int = 20,
                                   time_gap: int = 2000, testing=False):
        """models/LSTEP.py:139-220 — edge-feature / time mixer over the K most recent neighbours plus the
        masked mean of node features over the `time_gap` most recent neighbours."""
        if self.weighted_sum:
            raise NotImplementedError("weighted_sum ablation is outside the B200 hot path")
        assert num_neighbors > 0 and time_gap > 0, 'Number of sampled neighbors for each node should be greater than 0!'
        s = self.neighbor_sampler
        s._require_recent()
        node_ids = np.asarray(node_ids)
        self._check_ids(node_ids, s.num_rows, "neighbor sampler")
        n = len(node_ids)
        n_valid = min(n, len(node_interact_times))
        ids_dev, t_dev = self._upload([(node_ids, I64), (node_interact_times, F64)])
        with torch.cuda.device(ids_dev.device):
            # edge / time mixer (LSTEP.py:152-167): edge_mlp_1 then edge_agg are both linear, so the fused kernel forms
            # X = sum_k a_k [tf_k || edge_feat_k] and the 272x272 Linear runs once per row instead of once per neighbour
            t_dim, Fe = self.time_feat_dim, self.edge_raw_features.shape[1]
            K = int(num_neighbors)
            if K != self.edge_agg.weight.shape[1]:
                raise RuntimeError(f"mat1 and mat2 shapes cannot be multiplied (edge_agg expects {self.edge_agg.weight.shape[1]} neighbours, got {K})")
            if torch.is_grad_enabled() and (self.edge_agg.weight.requires_grad or self.edge_mlp_1.weight.requires_grad):
                # training: gradients must reach edge_agg / edge_mlp_1 through the per-neighbour form -> torch ops on the device lookup
                nbr, eid, nt = self._sample_full(ids_dev, t_dev, n, n_valid, K)
                dt = (t_dev[:, None] - nt.to(torch.float64)).float()  # f64 - f32 -> f64 -> float (LSTEP.py:153)
                tf = self.time_encoder(dt).masked_fill((nbr == 0).unsqueeze(-1), 0.0)
                x = self.edge_mlp_1(torch.cat([tf, self.edge_raw_features[eid]], dim=-1))
                x = self.edge_agg(x.permute(0, 2, 1)).squeeze()
            else:
                a_w = self.edge_agg.weight.detach().reshape(-1).contiguous()
                X = torch.empty((n, t_dim + Fe), dtype=torch.float32, device=ids_dev.device)
                tw = self.time_encoder.w.weight.detach().reshape(-1).contiguous()
                ef = self.edge_raw_features if self.edge_raw_features.is_contiguous() else self.edge_raw_features.contiguous()
                _lib.check(_lib.load().lstep_feature_aggregate(s.csr_ref, _lib.ptr(ids_dev), _lib.ptr(t_dev), n, n_valid, K, _lib.ptr(ef), ef.shape[0],
                                                               Fe, _lib.ptr(tw), t_dim, _lib.ptr(a_w), _lib.ptr(X), _lib.ptr(s._err), _lib.stream_ptr()),
                           "lstep_feature_aggregate")
                x = F.linear(X, self.edge_mlp_1.weight, self.edge_mlp_1.bias * a_w.sum()) + self.edge_agg.bias
                if n == 1:
                    x = x.squeeze()
            x = self.edge_mlp_2(F.relu(x))
            if self.use_dropout:
                x = F.dropout(x, p=self.dropout)
            own = self.node_raw_features[ids_dev]
            if self._node_feats_all_zero:
                agg = torch.zeros_like(own)  # every gathered feature row is zero: the masked mean is exactly zero
            else:
                g_nbr, _ = s.sample_device(ids_dev, t_dev, n, n_valid, int(time_gap))
                valid = g_nbr > 0
                cnt = valid.sum(dim=1, keepdim=True)
                # softmax over {1, -1e10}: 1/n_valid on valid slots, uniform when none is valid (LSTEP.py:183-186)
                scores = torch.where(cnt > 0, valid.float() / cnt.clamp(min=1).float(), torch.full_like(valid, 1.0 / time_gap, dtype=torch.float32))
                agg = torch.mean(self.node_raw_features[g_nbr.long()] * scores.unsqueeze(-1), dim=1)
            return self.node_mlp(torch.cat([agg + own, x], dim=-1))

    # ---- a6: neighbourhood PE ---------------------------------------------------------------
    def compute_neighborhood_pe(self, pe, node_ids: np.ndarray, node_interact_times: np.ndarray, num_neighbors: int = 30):
        """models/LSTEP.py:222-249."""
        assert num_neighbors > 0, 'Number of sampled neighbors for each node should be greater than 0!'
        self._require_cuda(pe, "pe")
        d, t = self.pe_dim, self.time_feat_dim
        if pe.dim() != 2 or pe.shape[1] != d:
            raise RuntimeError(f"pe must be [num_nodes, {d}], got {tuple(pe.shape)}")
        s = self.neighbor_sampler
        s._require_recent()
        node_ids = np.asarray(node_ids)
        n = len(node_ids)
        if len(node_interact_times) != n:
            raise RuntimeError(f"The size of tensor a ({len(node_interact_times)}) must match the size of tensor b ({n}) at non-singleton dimension 0")
        self._check_ids(node_ids, s.num_rows, "neighbor sampler")
        self._check_ids(node_ids, pe.shape[0], "pe")
        if s.num_rows > pe.shape[0]:
            raise IndexError(f"pe has {pe.shape[0]} rows but the sampler knows node ids up to {s.num_rows - 1}")
        ids_dev, t_dev = self._upload([(node_ids, I64), (node_interact_times, F64)])
        return self.compute_neighborhood_pe_device(pe, ids_dev, t_dev, int(num_neighbors))

    def compute_neighborhood_pe_device(self, pe, ids_dev, t_dev, K: int, out=None):
        """Device-resident form (ids int64 / times float64 already in HBM, validated by the caller)."""
        s = self.neighbor_sampler
        d, t = self.pe_dim, self.time_feat_dim
        n = ids_dev.shape[0]
        lib = _lib.load()
        with torch.cuda.device(pe.device):
            nbr, nt = s.sample_device(ids_dev, t_dev, n, n, K)
            if torch.is_grad_enabled() and (pe.requires_grad or self.pe_neighbor_mlp_1.weight.requires_grad):
                tw = self.time_encoder.w.weight.detach().reshape(-1).contiguous()
                S = _NbrAggregateFn.apply(pe, nbr, nt, t_dev, tw, t)
                node_pe = pe[ids_dev]
                h = self.pe_neighbor_mlp_2(F.relu(self.pe_neighbor_mlp_1(S)))
                return node_pe + torch.tanh(self.self_update_neighbor_pe(node_pe) + h)
            pec = pe if pe.is_contiguous() else pe.contiguous()
            if out is None:
                out = torch.empty((n, d), dtype=torch.float32, device=pe.device)
            ws = torch.empty(max(n, 1) * (d + t + 3), dtype=torch.float32, device=pe.device)
            _lib.check(lib.lstep_neighborhood_pe(_lib.ptr(pec), pec.shape[0], _lib.ptr(ids_dev), _lib.ptr(t_dev), _lib.ptr(nbr),
                                                 _lib.ptr(nt), n, K, self._mlp_ref("nbr"), _lib.ptr(out), _lib.ptr(ws), ws.numel() * 4,
                                                 _lib.stream_ptr()), "lstep_neighborhood_pe")
            return out

    def combining_pe_raw_feat(self, pe, node_ids: np.ndarray, node_interact_times: np.ndarray, num_neighbors: int = 30,
                              time_gap: int = 2000, testing=False):
        """models/LSTEP.py:251-266."""
        emb = self.aggregated_node_embeddings(node_ids=node_ids, node_interact_times=node_interact_times,
                                              num_neighbors=num_neighbors, time_gap=time_gap, testing=testing)
        nbr_pe = self.compute_neighborhood_pe(pe, node_ids=node_ids, node_interact_times=node_interact_times,
                                              num_neighbors=num_neighbors)
        return self.out_node_emb(torch.cat([emb, nbr_pe], dim=-1))

    # ---- a7 + a8: PE update -------------------------------------------------------------------
    def update_pe(self, pe, node_ids: np.ndarray, edge_ids: np.ndarray, batch_src_node_ids: np.ndarray,
                  batch_dst_node_ids: np.ndarray, node_interact_times: np.ndarray, current_time, num_neighbors: int = 30,
                  time_gap: int = 2000):
        """models/LSTEP.py:268-341 — updates the caller's `pe` [V1, d] in place and returns it."""
        assert num_neighbors > 0, 'Number of sampled neighbors for each node should be greater than 0!'
        self._require_cuda(pe, "pe")
        d, t = self.pe_dim, self.time_feat_dim
        if pe.dim() != 2 or pe.shape[1] != d or not pe.is_contiguous():
            raise _lib.LstepError(f"update_pe writes in place and needs a contiguous [num_nodes, {d}] table")
        s = self.neighbor_sampler
        s._require_recent()
        V1 = pe.shape[0]
        node_ids = np.asarray(node_ids)
        src, dst = np.asarray(batch_src_node_ids), np.asarray(batch_dst_node_ids)
        times = np.asarray(node_interact_times)
        n_ids, n_edges = len(node_ids), len(src)
        if n_ids > 1 and not bool(np.all(node_ids[1:] > node_ids[:-1])) and len(np.unique(node_ids)) != n_ids:
            # phase A runs in place on pe[node_ids]: a duplicate id would let one row be read after another block rewrote it
            # (the reference's callers always pass torch.unique()'d ids: evaluate_model_utils.py:54-55, train:221-222)
            raise ValueError("update_pe: node_ids must be unique")
        if len(dst) != n_edges or len(times) != n_edges:
            raise RuntimeError("batch_src_node_ids, batch_dst_node_ids and node_interact_times must have the same length")
        for a, what in ((node_ids, "node_ids"), (src, "batch_src_node_ids"), (dst, "batch_dst_node_ids")):
            self._check_ids(a, V1, what)
        self._check_ids(node_ids[:min(n_ids, n_edges)], s.num_rows, "neighbor sampler")
        if s.num_rows > V1:
            raise IndexError(f"pe has {V1} rows but the sampler knows node ids up to {s.num_rows - 1}")
        ids_dev, src_dev, dst_dev, t_dev = self._upload([(node_ids, I64), (src, I64), (dst, I64), (times, F64)])
        return self.update_pe_device(pe, ids_dev, src_dev, dst_dev, t_dev, float(current_time), int(num_neighbors))

    def update_pe_device(self, pe, ids_dev, src_dev, dst_dev, t_dev, current_time: float, K: int):
        """Device-resident form of update_pe (inputs validated by the caller)."""
        s = self.neighbor_sampler
        d, t = self.pe_dim, self.time_feat_dim
        V1 = pe.shape[0]
        n_ids, n_edges = ids_dev.shape[0], src_dev.shape[0]
        lib = _lib.load()
        need = lib.lstep_update_pe_workspace_bytes(n_ids, n_edges, K, d, t, V1)
        with torch.cuda.device(pe.device):
            ws = self._update_ws
            if ws is None or ws[0].numel() < need or ws[1] != V1 or ws[0].device != pe.device:
                buf = torch.empty(int(need * 1.5) + 4096, dtype=torch.uint8, device=pe.device)
                _lib.check(lib.lstep_update_pe_workspace_init(_lib.ptr(buf), buf.numel(), V1, _lib.stream_ptr()), "workspace_init")
                ws = self._update_ws = (buf, V1)
            _lib.check(lib.lstep_update_pe(_lib.ptr(pe), V1, s.csr_ref, _lib.ptr(ids_dev), n_ids, _lib.ptr(src_dev), _lib.ptr(dst_dev),
                                           _lib.ptr(t_dev), n_edges, current_time, K, self._mlp_ref("update"), _lib.ptr(ws[0]),
                                           ws[0].numel(), _lib.ptr(s._err), _lib.stream_ptr()), "lstep_update_pe")
            with torch.no_grad():
                pe.narrow(0, 0, 0).zero_()  # in-place no-op: bumps the autograd version counter of the caller's tensor
        return pe
