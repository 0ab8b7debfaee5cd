"""Streaming PE-step API with a device-resident edge stream and PE-history ring buffer
(SURVEY §8(f) f1).

The reference's loops (train_LSTEP_link_prediction.py:205,224-230,301-306;
evaluate_model_utils.py:57-63,131-135) keep the PE history as a [V1, <=T+1, d] tensor that they
trim (full clone), slice, and re-concatenate every batch — 3 GB of copies per batch at Reddit
size — and hand every batch to the model as host numpy arrays. `PEStream` is an additional,
numerically identical way to run the same per-batch sequence

    a3  fft = fourier_transform_pe(ids, hist[:, -T:], batch_idx)
        cur = clone(hist[:, -1]); cur[ids] = fft
    a6  C x compute_neighborhood_pe(cur, q_ids, times)
    a7/a8 update_pe(cur, ids, ..., current_time = max(times))
        hist = cat(hist, cur)

with the whole edge stream uploaded once and the history held in a ring of T slots, so a step
moves only the bytes the algorithm needs. `export_history()` materialises the reference layout
[V1, Th, d] (chronological, oldest first) for torch.save / parity checks, `import_history()`
adopts one (utils/EarlyStopping.py:79-82,100-104).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .model import LSTEP


class PEStream:
    def __init__(self, model: LSTEP, src_node_ids: np.ndarray, dst_node_ids: np.ndarray, node_interact_times: np.ndarray,
                 batch_size: int, num_neighbors: int = 20, initial_pe: torch.Tensor = None, history: torch.Tensor = None,
                 start: int = 0, stop: int = None):
        self.model = model
        self.K = int(num_neighbors)
        self.B = int(batch_size)
        self.T = model.num_fft_batches
        self.d = model.pe_dim
        dev = self.dev = model._dev()
        if dev.type != "cuda":
            raise _lib.LstepError("PEStream needs the model on a CUDA device (no CPU path)")
        stop = len(src_node_ids) if stop is None else stop
        self.start, self.stop = int(start), int(stop)
        self.src_np = np.ascontiguousarray(src_node_ids, dtype=np.int64)
        self.dst_np = np.ascontiguousarray(dst_node_ids, dtype=np.int64)
        self.t_np = np.ascontiguousarray(node_interact_times, dtype=np.float64)
        # whole stream resident in HBM (24 B/edge)
        self.src = torch.from_numpy(self.src_np).to(dev)
        self.dst = torch.from_numpy(self.dst_np).to(dev)
        self.t = torch.from_numpy(self.t_np).to(dev)
        # per-batch sorted unique node ids (the loops compute them on the host with torch.unique:
        # evaluate_model_utils.py:54-55); done once for the whole stream here
        self.batch_lo = list(range(self.start, self.stop, self.B))
        ids, off = [], [0]
        for lo in self.batch_lo:
            hi = min(lo + self.B, self.stop)
            u = np.unique(np.concatenate([self.src_np[lo:hi], self.dst_np[lo:hi]]))
            ids.append(u)
            off.append(off[-1] + len(u))
        self.ids_off = off
        self.ids_np = np.concatenate(ids) if ids else np.zeros(0, np.int64)
        self.ids = torch.from_numpy(self.ids_np).to(dev)
        self.num_batches = len(self.batch_lo)
        self.V1 = None
        self.batch_idx = 0
        if history is not None:
            self.import_history(history)
        elif initial_pe is not None:
            self.import_history(initial_pe.unsqueeze(1))

    # ---- history ring -----------------------------------------------------------------------
    def import_history(self, history: torch.Tensor):
        """Adopt a reference-layout history [V1, Th, d] (chronological)."""
        history = history.to(self.dev, torch.float32)
        V1, Th, d = history.shape
        assert d == self.d
        if Th > self.T:
            history = history[:, -self.T:, :]
            Th = self.T
        self.V1 = V1
        self.ring = torch.zeros((V1, self.T, d), dtype=torch.float32, device=self.dev)  # node-major: 68.8 KB / node
        self.ring[:, :Th, :] = history
        self.len = Th  # number of valid steps
        self.head = 0  # slot of the oldest valid step
        self.cur = torch.empty((V1, d), dtype=torch.float32, device=self.dev)

    def export_history(self) -> torch.Tensor:
        idx = (self.head + torch.arange(self.len, device=self.dev)) % self.T
        return self.ring.index_select(1, idx).contiguous()

    def _last_slot(self) -> int:
        return (self.head + self.len - 1) % self.T

    # ---- one batch --------------------------------------------------------------------------
    def batch_arrays(self, b: int):
        lo = self.batch_lo[b]
        hi = min(lo + self.B, self.stop)
        return lo, hi, self.ids_off[b], self.ids_off[b + 1]

    def step(self, b: int, queries, outs=None, batch_idx: int = None):
        """Run batch b. `queries`: list of device int64 tensors [B_b] of node ids for the
        compute_neighborhood_pe calls (positive sources / destinations, negatives), all queried at
        the batch's edge times. Returns the list of [B_b, d] outputs; the updated table is in
        self.cur and has been appended to the ring."""
        m = self.model
        lo, hi, io, ie = self.batch_arrays(b)
        ids = self.ids[io:ie]
        src, dst, t = self.src[lo:hi], self.dst[lo:hi], self.t[lo:hi]
        bi = self.batch_idx if batch_idx is None else batch_idx
        T, d = self.T, self.d
        masked = self.len < T
        bmask = min(max(bi, 0), T) if masked else T
        with torch.no_grad():
            fft = m.fourier_transform_pe_device(ids, self.ring, bmask, False, s0=self.head, ring=T, Th=self.len,
                                                node_stride=T * d, time_stride=d)
            self.cur.copy_(self.ring[:, self._last_slot(), :])  # current table = last snapshot ...
            self.cur.index_copy_(0, ids, fft)  # ... with the batch nodes replaced by the filtered history
            res = []
            for qi, q in enumerate(queries):
                res.append(m.compute_neighborhood_pe_device(self.cur, q, t, self.K, out=None if outs is None else outs[qi][:hi - lo]))
            m.update_pe_device(self.cur, ids, src, dst, t, float(self.t_np[lo:hi].max()), self.K)
            # append: overwrite the oldest slot once the ring is full
            if self.len < T:
                slot = (self.head + self.len) % T
                self.len += 1
            else:
                slot = self.head
                self.head = (self.head + 1) % T
            self.ring[:, slot, :] = self.cur
        self.batch_idx = bi + 1
        return res

    def step_host(self, src: np.ndarray, dst: np.ndarray, times: np.ndarray, query_ids, batch_idx: int = None):
        """Same step fed from HOST arrays (what a loop holding numpy batches calls): uploads the batch,
        runs it, and returns per-query row sums [len(query_ids), B] on the host (a small per-batch
        result, like the predictions the eval loop reads back)."""
        m = self.model
        ids_np = np.unique(np.concatenate([src, dst]))
        I64, F64 = np.dtype(np.int64), np.dtype(np.float64)
        up = m._upload([(ids_np, I64), (src, I64), (dst, I64), (times, F64)] + [(q, I64) for q in query_ids])
        ids, s_dev, d_dev, t_dev = up[:4]
        queries = up[4:]
        bi = self.batch_idx if batch_idx is None else batch_idx
        T, d = self.T, self.d
        bmask = min(max(bi, 0), T) if self.len < T else T
        with torch.no_grad():
            fft = m.fourier_transform_pe_device(ids, self.ring, bmask, False, s0=self.head, ring=T, Th=self.len,
                                                node_stride=T * d, time_stride=d)
            self.cur.copy_(self.ring[:, self._last_slot(), :])
            self.cur.index_copy_(0, ids, fft)
            sums = torch.stack([m.compute_neighborhood_pe_device(self.cur, q, t_dev, self.K).sum(dim=1) for q in queries])
            m.update_pe_device(self.cur, ids, s_dev, d_dev, t_dev, float(times.max()), self.K)
            if self.len < T:
                slot = (self.head + self.len) % T
                self.len += 1
            else:
                slot = self.head
                self.head = (self.head + 1) % T
            self.ring[:, slot, :] = self.cur
            host = sums.cpu()  # D2H of the step's result (synchronises)
        self.batch_idx = bi + 1
        return host.numpy()
