"""Drop-in for the reference's temporal neighbour sampler (utils/utils.py:70-301), backed by a
time-sorted CSR in HBM and the warp-cooperative lookup kernel (csrc/sampler.cu).

Same names, argument meaning and error behaviour as the reference:

    get_neighbor_sampler(data, sample_neighbor_strategy, time_scaling_factor, seed)
    NeighborSampler(adj_list, sample_neighbor_strategy, time_scaling_factor, seed)
        .get_historical_neighbors(node_ids, node_interact_times, num_neighbors=20)
            -> (int64[n,K], int64[n,K], float32[n,K]) numpy arrays
        .find_neighbors_before / .get_multi_hop_neighbors / .get_all_first_hop_neighbors
        .reset_random_state(), .sample_neighbor_strategy, .seed

Only the 'recent' strategy — the one LSTEP's defaults and best-configs use
(utils/load_configs.py:22,81-96) — runs on the device; 'uniform' / 'time_interval_aware' draw from
numpy's RandomState on the host in the reference and are outside the hot path:
asking for them raises NotImplementedError instead of silently computing on the CPU.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib

_STRATEGIES = ("uniform", "recent", "time_interval_aware")


def _dev(device=None) -> torch.device:
    if device is None:
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device(device)


class NeighborSampler:
    def __init__(self, adj_list: list = None, sample_neighbor_strategy: str = "uniform", time_scaling_factor: float = 0.0,
                 seed: int = None, device=None, _csr_tensors=None):
        """adj_list: list (index = node id) of lists of (neighbor_id, edge_id, timestamp) tuples, as in the
        reference (utils/utils.py:72-80). Each list is stably sorted by timestamp on the device."""
        self.sample_neighbor_strategy = sample_neighbor_strategy
        self.time_scaling_factor = time_scaling_factor
        self.seed = seed
        if self.seed is not None:
            self.random_state = np.random.RandomState(self.seed)
        self._lib = _lib.load()
        self.device = _dev(device)
        self._err = torch.zeros(1, dtype=torch.int32, device=self.device)
        if _csr_tensors is not None:
            self._set_csr(*_csr_tensors)
            return
        owner = np.fromiter((i for i, lst in enumerate(adj_list) for _ in lst), dtype=np.int64)
        n = owner.shape[0]
        flat = [x for lst in adj_list for x in lst]
        nbr = np.fromiter((x[0] for x in flat), dtype=np.int64, count=n)
        eid = np.fromiter((x[1] for x in flat), dtype=np.int64, count=n)
        t = np.fromiter((x[2] for x in flat), dtype=np.float64, count=n)
        self._build("lstep_csr_build_from_entries", (owner, nbr, eid, t), n, n, len(adj_list))

    # ---- construction -----------------------------------------------------------------------
    @classmethod
    def from_edges(cls, src_node_ids, dst_node_ids, edge_ids, node_interact_times, sample_neighbor_strategy="recent",
                   time_scaling_factor=0.0, seed=None, device=None, num_rows=None):
        """Build straight from an edge stream (what get_neighbor_sampler does, utils/utils.py:292-299)
        without materialising python adjacency lists. Arrays may be numpy or CUDA tensors."""
        self = cls.__new__(cls)
        NeighborSampler.__init__(self, None, sample_neighbor_strategy, time_scaling_factor, seed, device,
                                 _csr_tensors=())  # placeholder, replaced below
        E = int(len(src_node_ids))
        if num_rows is None:
            mx = max(int(src_node_ids.max()), int(dst_node_ids.max())) if E else 0
            num_rows = mx + 1
        self._build("lstep_csr_build_from_edges", (src_node_ids, dst_node_ids, edge_ids, node_interact_times), E, 2 * E,
                    int(num_rows))
        return self

    @classmethod
    def from_reference(cls, ref_sampler, device=None):
        """Adopt an instance of the reference's own NeighborSampler (already-sorted per-node arrays)."""
        adj = [list(zip(np.asarray(a).tolist(), np.asarray(b).tolist(), np.asarray(c).tolist()))
               for a, b, c in zip(ref_sampler.nodes_neighbor_ids, ref_sampler.nodes_edge_ids, ref_sampler.nodes_neighbor_times)]
        return cls(adj, ref_sampler.sample_neighbor_strategy, getattr(ref_sampler, "time_scaling_factor", 0.0),
                   ref_sampler.seed, device)

    def _to_dev(self, a, dtype):
        if isinstance(a, torch.Tensor):
            return a.to(device=self.device, dtype=dtype).contiguous()
        return torch.from_numpy(np.ascontiguousarray(a, dtype={torch.int64: np.int64, torch.float64: np.float64}[dtype])).to(self.device)

    def _build(self, fn_name, arrays, n_in, n_entries, num_rows):
        lib, dev = self._lib, self.device
        a0, a1, a2 = (self._to_dev(a, torch.int64) for a in arrays[:3])
        a3 = self._to_dev(arrays[3], torch.float64)
        indptr = torch.empty(num_rows + 1, dtype=torch.int64, device=dev)
        nbr = torch.empty(max(n_entries, 1), dtype=torch.int32, device=dev)
        eid = torch.empty(max(n_entries, 1), dtype=torch.int32, device=dev)
        t = torch.empty(max(n_entries, 1), dtype=torch.float64, device=dev)
        ws_bytes = lib.lstep_csr_build_workspace_bytes(n_entries, num_rows)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        self._err.zero_()
        with torch.cuda.device(dev):
            rc = getattr(lib, fn_name)(_lib.ptr(a0), _lib.ptr(a1), _lib.ptr(a2), _lib.ptr(a3), n_in, num_rows,
                                       _lib.ptr(indptr), _lib.ptr(nbr), _lib.ptr(eid), _lib.ptr(t), _lib.ptr(ws), ws_bytes,
                                       _lib.ptr(self._err), _lib.stream_ptr())
        _lib.check(rc, fn_name)
        if int(self._err.item()) & _lib.FLAG_NODE_OUT_OF_RANGE:
            raise IndexError("node / edge id out of range while building the temporal adjacency")
        del ws
        self._set_csr(indptr, nbr[:n_entries], eid[:n_entries], t[:n_entries])

    def _set_csr(self, *tensors):
        if not tensors:
            return
        self.indptr, self.nbr, self.eid, self.t = tensors
        self.num_rows = int(self.indptr.shape[0] - 1)
        self.nnz = int(self.nbr.shape[0])
        self.csr = _lib.CSR(self.indptr.data_ptr(), self.nbr.data_ptr(), self.eid.data_ptr(), self.t.data_ptr(), self.num_rows,
                            self.nnz)
        self.csr_ref = C.byref(self.csr)

    # ---- reference API ----------------------------------------------------------------------
    def reset_random_state(self):
        self.random_state = np.random.RandomState(self.seed)

    def _require_recent(self):
        if self.sample_neighbor_strategy == "recent":
            return
        if self.sample_neighbor_strategy in _STRATEGIES:
            raise NotImplementedError(f"sample_neighbor_strategy '{self.sample_neighbor_strategy}' draws from a host RNG in the "
                                      "reference and is outside the B200 hot path; use 'recent'")
        raise ValueError(f"Not implemented error for sample_neighbor_strategy {self.sample_neighbor_strategy}!")

    def sample_device(self, node_ids_dev: torch.Tensor, times_dev: torch.Tensor, n_rows: int, n_valid: int, num_neighbors: int):
        """Device-resident lookup for the kernels downstream: (int32[n,K] ids, float32[n,K] times).
        Errors (out-of-range ids) accumulate in a device flag; see check_errors()."""
        K = int(num_neighbors)
        out_n = torch.empty((n_rows, K), dtype=torch.int32, device=self.device)
        out_t = torch.empty((n_rows, K), dtype=torch.float32, device=self.device)
        rc = self._lib.lstep_sample_recent_compact(self.csr_ref, _lib.ptr(node_ids_dev), _lib.ptr(times_dev), n_rows, n_valid, K,
                                                   _lib.ptr(out_n), _lib.ptr(out_t), _lib.ptr(self._err), _lib.stream_ptr())
        _lib.check(rc, "lstep_sample_recent_compact")
        return out_n, out_t

    def check_errors(self):
        """Synchronising check of the device error flag; raises what the reference would."""
        flag = int(self._err.item())
        if flag:
            self._err.zero_()
        if flag & _lib.FLAG_NODE_OUT_OF_RANGE:
            raise IndexError("list index out of range")  # utils/utils.py:140 on an unknown node id

    def get_historical_neighbors(self, node_ids: np.ndarray, node_interact_times: np.ndarray, num_neighbors: int = 20):
        assert num_neighbors > 0, 'Number of sampled neighbors for each node should be greater than 0!'
        self._require_recent()
        K = int(num_neighbors)
        n_rows = len(node_ids)
        n_valid = min(n_rows, len(node_interact_times))  # zip() truncation (utils/utils.py:169)
        dev = self.device
        if n_rows == 0:
            return (np.zeros((0, K), np.int64), np.zeros((0, K), np.int64), np.zeros((0, K), np.float32))
        q_n = torch.from_numpy(np.ascontiguousarray(node_ids, dtype=np.int64)).to(dev, non_blocking=True)
        q_t = torch.from_numpy(np.ascontiguousarray(node_interact_times, dtype=np.float64)).to(dev, non_blocking=True)
        out_n = torch.empty((n_rows, K), dtype=torch.int64, device=dev)
        out_e = torch.empty((n_rows, K), dtype=torch.int64, device=dev)
        out_t = torch.empty((n_rows, K), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            rc = self._lib.lstep_sample_recent(self.csr_ref, _lib.ptr(q_n), _lib.ptr(q_t), n_rows, n_valid, K, _lib.ptr(out_n),
                                               _lib.ptr(out_e), _lib.ptr(out_t), _lib.ptr(self._err), _lib.stream_ptr())
        _lib.check(rc, "lstep_sample_recent")
        res = (out_n.cpu().numpy(), out_e.cpu().numpy(), out_t.cpu().numpy())
        self.check_errors()
        return res

    def find_neighbors_before(self, node_id: int, interact_time: float, return_sampled_probabilities: bool = False):
        """All interactions of node_id strictly before interact_time, sorted by time (utils/utils.py:129-146)."""
        if return_sampled_probabilities:
            raise NotImplementedError("sampled probabilities belong to the 'time_interval_aware' strategy")
        node_id = int(node_id)
        if node_id >= self.num_rows or node_id < -self.num_rows:
            raise IndexError("list index out of range")
        node_id %= self.num_rows
        lo, hi = (int(x) for x in self.indptr[node_id:node_id + 2].tolist())
        t = self.t[lo:hi].cpu().numpy()
        i = int(np.searchsorted(t, interact_time))
        return (self.nbr[lo:lo + i].cpu().numpy().astype(np.int64), self.eid[lo:lo + i].cpu().numpy().astype(np.int64), t[:i], None)

    def get_multi_hop_neighbors(self, num_hops: int, node_ids: np.ndarray, node_interact_times: np.ndarray, num_neighbors: int = 20):
        assert num_hops > 0, 'Number of sampled hops should be greater than 0!'
        n = len(node_ids)
        ids, eids, times = self.get_historical_neighbors(node_ids, node_interact_times, num_neighbors)
        out = ([ids], [eids], [times])
        for _ in range(1, num_hops):
            ids, eids, times = self.get_historical_neighbors(out[0][-1].flatten(), out[2][-1].flatten(), num_neighbors)
            for lst, a in zip(out, (ids, eids, times)):
                lst.append(a.reshape(n, -1))
        return out

    def get_all_first_hop_neighbors(self, node_ids: np.ndarray, node_interact_times: np.ndarray):
        res = ([], [], [])
        for node_id, tq in zip(node_ids, node_interact_times):
            a, b, c, _ = self.find_neighbors_before(node_id, tq)
            res[0].append(a), res[1].append(b), res[2].append(c)
        return res


def get_neighbor_sampler(data, sample_neighbor_strategy: str = "uniform", time_scaling_factor: float = 0.0, seed: int = None,
                         device=None) -> NeighborSampler:
    """Same contract as utils/utils.py:282-301: `data` has src_node_ids, dst_node_ids, edge_ids,
    node_interact_times; the adjacency is undirected and each node's list is sorted by time."""
    return NeighborSampler.from_edges(data.src_node_ids, data.dst_node_ids, data.edge_ids, data.node_interact_times,
                                      sample_neighbor_strategy, time_scaling_factor, seed, device)
