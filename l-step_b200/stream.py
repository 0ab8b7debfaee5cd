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

import ctypes

import numpy as np
import torch

from . import _lib
from .model import LSTEP

C_void_p = ctypes.c_void_p
C_byref = ctypes.byref
_I64, _F64 = np.dtype(np.int64), np.dtype(np.float64)


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
        self.batch_tmax = [float(self.t_np[lo:min(lo + self.B, self.stop)].max()) for lo in self.batch_lo]
        self._host_stepper = None
        self._ticket_shape = {}
        self._drained = {}  # public ticket -> result read out before its stepper was replaced
        self._epoch = 0
        self._libc = _lib.load()
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
        # ids of the resident stream against the table and the sampler's CSR, once (the device-resident step reads and WRITES
        # table / ring rows by these ids; the reference raises IndexError for an unknown node at its first lookup, Q8)
        if len(self.ids_np):
            lo_id, hi_id = int(self.ids_np.min()), int(self.ids_np.max())
            if lo_id < 0 or hi_id >= V1:
                raise IndexError(f"edge stream holds node id {hi_id if hi_id >= V1 else lo_id} but the PE table has {V1} rows")
            samp = self.model.neighbor_sampler
            if samp is not None and hi_id >= samp.num_rows:
                raise IndexError(f"list index out of range: node id {hi_id} is unknown to the neighbor sampler ({samp.num_rows} rows)")
            if samp is not None and samp.num_rows > V1:
                raise IndexError(f"pe has {V1} rows but the sampler knows node ids up to {samp.num_rows - 1}")
        self.V1 = V1
        self.ring = torch.zeros((V1, self.T, d), dtype=torch.float32, device=self.dev)  # node-major: 68.8 KB / node
        self.ring[:, :Th, :] = history
        self.len = Th  # number of valid steps
        self.head = 0  # slot of the oldest valid step
        self.cur = torch.empty((V1, d), dtype=torch.float32, device=self.dev)
        lib = _lib.load()
        with torch.cuda.device(self.dev):
            _lib.check(lib.lstep_ring_load(_lib.ptr(self.ring), _lib.ptr(self.cur), V1, self.T, d, self._last_slot(),
                                           _lib.stream_ptr()), "lstep_ring_load")
        self.desc = _lib.PEStreamDesc(self.src.data_ptr(), self.dst.data_ptr(), self.t.data_ptr(), self.ring.data_ptr(),
                                      self.cur.data_ptr(), V1, self.T, d)
        self._ws = None
        self._ws_cap = (0, 0, 0)
        self.steps_done = 0

    def check_errors(self):
        """Synchronising read of the device error flag the step's lookups raise for query ids outside the sampler's
        rows (such a query yields an empty neighbourhood and reads table row 0; nothing is written by query ids).
        Raises IndexError like the reference's lookup (utils/utils.py:140, Q8). Called by export_history(), result()
        and run(check=True); the host-fed calls validate ids on the host before anything is enqueued. Call it yourself
        after a loop of step()."""
        self.model.neighbor_sampler.check_errors()

    def export_history(self) -> torch.Tensor:
        """The history in the reference's layout [V1, Th, d], oldest first (for save_pe / torch.cat)."""
        self.check_errors()
        idx = (self.head + torch.arange(self.len, device=self.dev)) % self.T
        return self.ring.index_select(1, idx).contiguous()

    def _last_slot(self) -> int:
        return (self.head + self.len - 1) % self.T

    # ---- one batch --------------------------------------------------------------------------
    def batch_arrays(self, b: int):
        lo = self.batch_lo[b]
        hi = min(lo + self.B, self.stop)
        return lo, hi, self.ids_off[b], self.ids_off[b + 1]

    def _workspace(self, n_ids, n_edges, C):
        lib = _lib.load()
        m = self.model
        need = lib.lstep_pe_step_workspace_bytes(n_ids, n_edges, C, self.K, self.d, m.time_feat_dim, self.V1)
        if self._ws is None or self._ws.numel() < need:
            self._ws_cap = (max(2 * self.B, n_ids), max(self.B, n_edges), max(C, 4))  # (ids, edges, query sets) it is sized for
            cap = lib.lstep_pe_step_workspace_bytes(self._ws_cap[0], self._ws_cap[1], self._ws_cap[2], self.K, self.d,
                                                    m.time_feat_dim, self.V1)
            self._ws = torch.empty(max(cap, need) + 4096, dtype=torch.uint8, device=self.dev)
            _lib.check(lib.lstep_update_pe_workspace_init(_lib.ptr(self._ws), self._ws.numel(), self.V1, _lib.stream_ptr()),
                       "workspace_init")
        return self._ws

    def _run(self, lo, n_edges, ids, tmax, queries, out, batch_idx):
        """Device-resident core shared by step() and step_host()."""
        lib = _lib.load()
        m = self.model
        T = self.T
        bi = self.batch_idx if batch_idx is None else batch_idx
        bmask = min(max(bi, 0), T) if self.len < T else T  # mask keyed on batch_idx while the history is short (Q5)
        C = len(queries)
        with torch.cuda.device(self.dev), torch.no_grad():
            G = m._collapsed_filter(bmask, False)
            ws = self._workspace(ids.shape[0], n_edges, C)
            if self.len < T:
                slot, new_head, new_len = (self.head + self.len) % T, self.head, self.len + 1
            else:
                slot, new_head, new_len = self.head, (self.head + 1) % T, T
            qptrs = (C_void_p * max(C, 1))(*[q.data_ptr() for q in queries])
            _lib.check(lib.lstep_pe_step(self._desc_ref, m.neighbor_sampler.csr_ref, lo, n_edges, _lib.ptr(ids), ids.shape[0],
                                         float(tmax), self.head, self.len, slot, _lib.ptr(G), qptrs, C, _lib.ptr(out), self.K,
                                         m._mlp_ref("nbr"), m._mlp_ref("update"), _lib.ptr(ws), ws.numel(),
                                         _lib.ptr(m.neighbor_sampler._err), _lib.stream_ptr()), "lstep_pe_step")
            self.head, self.len = new_head, new_len
        self.batch_idx = bi + 1
        self.steps_done += 1

    @property
    def _desc_ref(self):
        return C_byref(self.desc)

    def step(self, b: int, queries, out: torch.Tensor = None, batch_idx: int = None):
        """Run batch b of the resident stream. `queries`: list of device int64 tensors [B_b] of node ids
        for the compute_neighborhood_pe calls (positive sources / destinations, negatives), all queried
        at the batch's edge times. Returns a [C, B_b, d] tensor of neighbourhood PEs; the updated
        table is self.cur and has been appended to the ring."""
        lo, hi, io, ie = self.batch_arrays(b)
        n = hi - lo
        C = len(queries)
        if out is None:
            out = torch.empty((max(C, 1), n, self.d), dtype=torch.float32, device=self.dev)
        else:
            out = out.view(-1)[:max(C, 1) * n * self.d].view(max(C, 1), n, self.d)
        self._run(lo, n, self.ids[io:ie], self.batch_tmax[b], queries, out, batch_idx)
        return out

    def run(self, b0: int, n_steps: int, queries, out: torch.Tensor = None, check: bool = False) -> torch.Tensor:
        """Batches b0 .. b0+n_steps-1 of the resident stream in ONE native call (lstep_pe_steps): the per-batch loop
        runs in C, so the host pays six kernel launches per step and nothing else. `queries`: list of device int64
        tensors holding, for every query set, the node ids of ALL edges of those batches back to back (entry e belongs
        to edge lo(b0) + e). Returns the neighbourhood PEs of the LAST step [C, n_last, d] (or fills `out`
        [n_steps, C, B, d] with every step's). While the history ring is still filling, falls back to step().
        check=True: read the device error flag afterwards (synchronises; IndexError for a query id unknown to the sampler)."""
        if n_steps <= 0:
            return None
        lo0 = self.batch_lo[b0]
        C = len(queries)
        if self.len < self.T:
            last = None
            for i in range(n_steps):
                lo, hi, _, _ = self.batch_arrays(b0 + i)
                last = self.step(b0 + i, [q[lo - lo0:hi - lo0] for q in queries], None if out is None else out[i])
            return last
        lib, m = self._libc, self.model
        B, d = self.B, self.d
        los = np.asarray(self.batch_lo[b0:b0 + n_steps], dtype=np.int64)
        his = np.minimum(los + B, self.stop)
        ne = (his - los).astype(np.int64)
        ids_off = np.asarray(self.ids_off[b0:b0 + n_steps + 1], dtype=np.int64)
        tmax = np.asarray(self.batch_tmax[b0:b0 + n_steps], dtype=np.float64)
        q_off = (los - lo0).astype(np.int64)
        n_ids_max = int(np.diff(ids_off).max())
        with torch.cuda.device(self.dev), torch.no_grad():
            G = m._collapsed_filter(self.T, False)
            ws = self._workspace(n_ids_max, int(ne.max()), C)
            if out is None:
                buf = torch.empty((max(C, 1), B, d), dtype=torch.float32, device=self.dev)
                stride = 0
            else:
                buf = out
                stride = max(C, 1) * B * d
                if ne.min() != B:
                    raise ValueError("run(out=...) needs full batches (a ragged last batch changes the [C, n, d] layout)")
            qptrs = (C_void_p * max(C, 1))(*[q.data_ptr() for q in queries])
            head, ln = ctypes.c_int(self.head), ctypes.c_int(self.len)
            rc = lib.lstep_pe_steps(self._desc_ref, m.neighbor_sampler.csr_ref, n_steps, los.ctypes.data, ne.ctypes.data,
                                    self.ids.data_ptr(), ids_off.ctypes.data, tmax.ctypes.data, C_byref(head), C_byref(ln), G.data_ptr(),
                                    qptrs, q_off.ctypes.data, C, buf.data_ptr(), stride, self.K, m._mlp_ref("nbr"), m._mlp_ref("update"),
                                    ws.data_ptr(), ws.numel(), m.neighbor_sampler._err.data_ptr(),
                                    torch.cuda.current_stream().cuda_stream)
            done = (head.value - self.head) % self.T if rc != 0 else n_steps
            self.head = head.value
            self.batch_idx += done
            self.steps_done += done
            if rc != 0:
                _lib.check(rc, "lstep_pe_steps")
        if check:
            self.check_errors()
        if out is not None:
            return out
        n_last = int(ne[-1])
        return buf.view(-1)[:max(C, 1) * n_last * d].view(max(C, 1), n_last, d)

    # ---- host-fed steps (native stager: csrc/host_step.cu) -----------------------------------------
    def _stepper(self, n_edges: int, C: int):
        lib = _lib.load()
        st = self._host_stepper
        if st is None or st[1] < n_edges or st[2] < C:
            if st is not None:
                # a larger batch / more query sets than the stepper was sized for: results of tickets still outstanding
                # are read out first (the native ticket numbering restarts with the new stepper; public tickets carry the
                # stepper's epoch so an old one can never alias a new one)
                for tk in sorted(self._ticket_shape):
                    try:
                        self._drained[tk] = self._result_now(tk)
                    except ValueError:  # result slot already reused by later steps (e.g. a run_host call): nothing to keep
                        pass
                self._ticket_shape.clear()
                lib.lstep_host_stepper_destroy(st[0])
                self._epoch += 1
            h = C_void_p()
            cap_e, cap_c = max(self.B, n_edges), max(C, 4)
            with torch.cuda.device(self.dev):
                _lib.check(lib.lstep_host_stepper_create(self.HOST_SLOTS, cap_e, cap_c, self.d, C_byref(h)), "lstep_host_stepper_create")
            self._host_stepper = st = (h, cap_e, cap_c)
        return st[0]

    HOST_SLOTS = 4  # steps that may be in flight before step_host_async blocks on the oldest

    def __del__(self):
        st = getattr(self, "_host_stepper", None)
        if st is not None:
            try:
                _lib.load(False).lstep_host_stepper_destroy(st[0])
            except Exception:
                pass
            self._host_stepper = None

    def step_host_async(self, src: np.ndarray, dst: np.ndarray, times: np.ndarray, query_ids, batch_idx: int = None,
                        ids: np.ndarray = None, out: torch.Tensor = None) -> int:
        """One step fed from HOST arrays (what a loop holding numpy batches calls): the batch's endpoints,
        times and the query id sets are packed into a pinned slot and uploaded by one async copy inside the
        native call, the step's kernels are enqueued behind it, and the per-query row sums [C, n] of the
        neighbourhood PEs are copied back to a pinned result slot. Returns a ticket at once (nothing
        synchronises); `result(ticket)` waits for that step only, so a loop can read results one step behind.
        `ids` (sorted unique batch nodes) is computed natively when omitted; `out` [C, n, d] keeps the full
        neighbourhood PEs on the device. This is the per-batch host path: it avoids every avoidable Python
        call (≈ 25 µs of interpreter time per step)."""
        lib = self._libc
        m = self.model
        n = len(src)
        C = len(query_ids)
        if src.dtype != _I64 or not src.flags.c_contiguous:
            src = np.ascontiguousarray(src, dtype=_I64)
        if dst.dtype != _I64 or not dst.flags.c_contiguous:
            dst = np.ascontiguousarray(dst, dtype=_I64)
        if times.dtype != _F64 or not times.flags.c_contiguous:
            times = np.ascontiguousarray(times, dtype=_F64)
        qs = [q if (q.dtype == _I64 and q.flags.c_contiguous) else np.ascontiguousarray(q, dtype=_I64) for q in query_ids]
        if len(dst) != n or len(times) != n or any(len(q) != n for q in qs):
            raise ValueError("step_host: src, dst, times and every query set must have the same length")
        if ids is not None:
            ids = np.ascontiguousarray(ids, dtype=_I64)
        T = self.T
        bi = self.batch_idx if batch_idx is None else batch_idx
        bmask = min(max(bi, 0), T) if self.len < T else T
        if torch.cuda.current_device() != self.dev.index:
            torch.cuda.set_device(self.dev)
        st = self._host_stepper
        if st is None or st[1] < n or st[2] < C:
            h = self._stepper(n, C)
        else:
            h = st[0]
        G = m._collapsed_filter(bmask, False)
        ws = self._ws
        if ws is None or n > self._ws_cap[1] or C > self._ws_cap[2] or 2 * n > self._ws_cap[0]:
            ws = self._workspace(len(ids) if ids is not None else 2 * n, n, C)
        if self.len < T:
            slot, new_head, new_len = (self.head + self.len) % T, self.head, self.len + 1
        else:
            slot, new_head, new_len = self.head, (self.head + 1) % T, T
        qptrs = (C_void_p * max(C, 1))(*[q.ctypes.data for q in qs])
        ticket = ctypes.c_int64(-1)
        samp = m.neighbor_sampler
        rc = lib.lstep_pe_step_host(h, self._desc_ref, samp.csr_ref, n, src.ctypes.data, dst.ctypes.data, times.ctypes.data,
                                    ids.ctypes.data if ids is not None else None, len(ids) if ids is not None else 0, self.head,
                                    self.len, slot, G.data_ptr(), qptrs, C, out.data_ptr() if out is not None else None, self.K,
                                    m._mlp_ref("nbr"), m._mlp_ref("update"), ws.data_ptr(), ws.numel(), samp._err.data_ptr(),
                                    torch.cuda.current_stream().cuda_stream, C_byref(ticket))
        if rc != 0:
            _lib.check(rc, "lstep_pe_step_host")
        self.head, self.len = new_head, new_len
        self.batch_idx = bi + 1
        self.steps_done += 1
        tk = (self._epoch << 32) | ticket.value
        self._ticket_shape[tk] = (C, n)
        self._ticket_shape.pop(tk - self.HOST_SLOTS, None)
        self._drained.pop(tk - 2 * self.HOST_SLOTS, None)
        m.h2d_bytes += 8 * n * (5 + C)  # what lstep_pe_step_host copies: src, dst, t, <= 2n ids, C query sets
        return tk

    def run_host(self, src: np.ndarray, dst: np.ndarray, times: np.ndarray, query_ids) -> np.ndarray:
        """A run of consecutive batches fed from HOST arrays in ONE native call (an evaluation split): batch b is edges
        [b*B, (b+1)*B) of the given arrays; every step copies its own inputs in and its own per-query row sums out
        (read one step behind inside the call). Returns [n_batches, C, B] float32 (tail of a ragged last batch zero).
        While the history ring is still filling (the filter changes with every step) the batches go through
        step_host_async one by one."""
        n_total, B, C = len(src), self.B, len(query_ids)
        src = np.ascontiguousarray(src, dtype=_I64)
        dst = np.ascontiguousarray(dst, dtype=_I64)
        times = np.ascontiguousarray(times, dtype=_F64)
        qs = [np.ascontiguousarray(q, dtype=_I64) for q in query_ids]
        if len(dst) != n_total or len(times) != n_total or any(len(q) != n_total for q in qs):
            raise ValueError("run_host: src, dst, times and every query set must have the same length")
        nb = (n_total + B - 1) // B
        res = np.zeros((nb, C, B), dtype=np.float32)
        lo = 0
        pend = []
        while lo < n_total and self.len < self.T:  # masked regime: per-step calls
            hi = min(lo + B, n_total)
            pend.append((lo // B, hi - lo, self.step_host_async(src[lo:hi], dst[lo:hi], times[lo:hi], [q[lo:hi] for q in qs])))
            if len(pend) > 1:
                b, n, tk = pend.pop(0)
                res[b, :, :n] = self.result(tk)
            lo = hi
        for b, n, tk in pend:
            res[b, :, :n] = self.result(tk)
        if lo >= n_total:
            return res
        m, lib = self.model, self._libc
        if torch.cuda.current_device() != self.dev.index:
            torch.cuda.set_device(self.dev)
        h = self._stepper(B, C)
        G = m._collapsed_filter(self.T, False)
        ws = self._workspace(2 * B, B, C)
        qptrs = (C_void_p * max(C, 1))(*[q[lo:].ctypes.data for q in qs])
        head, ln, done = ctypes.c_int(self.head), ctypes.c_int(self.len), ctypes.c_int64(0)
        sub = res[lo // B:]
        samp = m.neighbor_sampler
        rc = lib.lstep_pe_steps_host(h, self._desc_ref, samp.csr_ref, n_total - lo, B, src[lo:].ctypes.data, dst[lo:].ctypes.data,
                                     times[lo:].ctypes.data, qptrs, C, C_byref(head), C_byref(ln), G.data_ptr(), self.K,
                                     m._mlp_ref("nbr"), m._mlp_ref("update"), ws.data_ptr(), ws.numel(), samp._err.data_ptr(),
                                     torch.cuda.current_stream().cuda_stream, sub.ctypes.data, C_byref(done))
        # batches applied before an error stay applied: the host-side ring position follows the device ring either way
        steps = done.value
        self.head = head.value
        self.batch_idx += steps
        self.steps_done += steps
        m.h2d_bytes += 8 * min(steps * B, n_total - lo) * (5 + C)
        if rc != 0:
            _lib.check(rc, "lstep_pe_steps_host")
        return res

    def result(self, ticket: int) -> np.ndarray:
        """Per-query row sums [C, n] of step `ticket` (waits for that step's device-to-host copy)."""
        r = self._drained.pop(ticket, None)
        if r is not None:
            return r
        if ticket >> 32 != self._epoch or ticket not in self._ticket_shape:
            raise KeyError(f"result: ticket {ticket} is not outstanding (already read, or overwritten: the last {self.HOST_SLOTS} are kept)")
        return self._result_now(ticket)

    def _result_now(self, ticket: int) -> np.ndarray:
        p = ctypes.POINTER(ctypes.c_float)()
        nf = ctypes.c_int64(0)
        rc = self._libc.lstep_host_step_result(self._host_stepper[0], ticket & 0xFFFFFFFF, C_byref(p), C_byref(nf))
        if rc != 0:
            _lib.check(rc, "lstep_host_step_result")
        C, n = self._ticket_shape[ticket]
        if nf.value == 0:
            return np.zeros((C, n), np.float32)
        return np.ctypeslib.as_array(p, shape=(C, n)).copy()

    def step_host(self, src: np.ndarray, dst: np.ndarray, times: np.ndarray, query_ids, batch_idx: int = None, ids: np.ndarray = None):
        """Synchronous form: run the step and return its per-query row sums [len(query_ids), n]."""
        return self.result(self.step_host_async(src, dst, times, query_ids, batch_idx, ids))


class ChangeLogStream(PEStream):
    """PEStream on a CHANGE-LOG history (csrc/changelog.cu) instead of the dense ring: base rows + the rows every step
    changed, V1 + T * (rows changed per step) rows instead of V1 * T — 7.8 GB instead of 688 GB for the 10 M-node graph at
    T = 100, and ~9 MB written per step instead of 13.8 GB (the dense ring's append copies one row per node per step).
    Same API (step, cur, export_history / import_history); the DFT filter sums (events in window + 1) rows per node with
    span sums of G as weights. `event_capacity`: events one step may record (default: every row a batch can touch,
    2B + 2BK + 2, capped at V1); importing a dense history needs capacity for the rows that differ between consecutive
    snapshots. `row_mul` / `row_add`: node-id sharded groups keep only the nodes v % row_mul == row_add."""

    def __init__(self, *args, event_capacity: int = None, **kw):
        self._cap_arg = event_capacity
        super().__init__(*args, **kw)

    def import_history(self, history: torch.Tensor):
        history = history.to(self.dev, torch.float32)
        V1, Th, d = history.shape
        assert d == self.d
        if Th > self.T:
            history = history[:, -self.T:, :]
            Th = self.T
        if len(self.ids_np):
            lo_id, hi_id = int(self.ids_np.min()), int(self.ids_np.max())
            if lo_id < 0 or hi_id >= V1:
                raise IndexError(f"edge stream holds node id {hi_id if hi_id >= V1 else lo_id} but the PE table has {V1} rows")
            samp = self.model.neighbor_sampler
            if samp is not None and (hi_id >= samp.num_rows or samp.num_rows > V1):
                raise IndexError("list index out of range: the neighbor sampler's rows do not cover the stream / exceed the table")
        if self.T > 128:
            raise _lib.LstepError("the change-log history supports T <= 128 window steps")
        self.V1 = V1
        T = self.T
        cap = self._cap_arg if self._cap_arg is not None else min(V1, 2 * self.B * (self.K + 1) + 2)
        self.cap = cap = int(max(cap, 1))
        H = 1
        while H < 2 * cap:
            H *= 2
        dev = self.dev
        self.base = history[:, 0, :].contiguous().clone()
        self.ev_node = torch.zeros((T, cap), dtype=torch.int32, device=dev)
        self.ev_row = torch.empty((T, cap, d), dtype=torch.float32, device=dev)
        self.ev_cnt = torch.zeros(T, dtype=torch.int32, device=dev)
        self.ev_hash = torch.full((T, H), -1, dtype=torch.int64, device=dev)
        self.ev_mask = torch.zeros((V1, 4), dtype=torch.int32, device=dev)
        self.cl = _lib.ChangeLog(self.base.data_ptr(), self.ev_node.data_ptr(), self.ev_row.data_ptr(), self.ev_cnt.data_ptr(),
                                 self.ev_hash.data_ptr(), self.ev_mask.data_ptr(), V1, T, cap, H, d, 1, 0)
        self.ring = None
        self.head, self.len = 0, 1
        self.cur = self.base.clone()
        lib = _lib.load()
        err = torch.zeros(1, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            for s in range(1, Th):  # every later snapshot: the rows that differ from the previous one are that step's events
                snap = history[:, s, :].contiguous()
                changed = torch.nonzero((snap != self.cur).any(dim=1)).flatten()
                if changed.numel() > cap:
                    raise _lib.LstepError(f"snapshot {s} differs from the previous one in {changed.numel()} rows but event_capacity is {cap}")
                _lib.check(lib.lstep_changelog_append(ctypes.byref(self.cl), s, 0, _lib.ptr(snap), None, None, 0, _lib.ptr(changed), changed.numel(),
                                                      None, 0, 0, _lib.ptr(err), _lib.stream_ptr()), "lstep_changelog_append")
                self.cur = snap
                self.len = s + 1
            self.cur = self.cur.clone()
        self.desc = _lib.PEStreamDesc(self.src.data_ptr(), self.dst.data_ptr(), self.t.data_ptr(), None, self.cur.data_ptr(), V1, T, d)
        self._ws = None
        self._ws_cap = (0, 0, 0)
        self.steps_done = 0

    def history_bytes(self) -> int:
        """Bytes the history occupies (base + event rows + index structures) — against V1 * T * d * 4 for the dense ring."""
        return sum(t.numel() * t.element_size() for t in (self.base, self.ev_node, self.ev_row, self.ev_cnt, self.ev_hash, self.ev_mask))

    def export_history(self) -> torch.Tensor:
        """The dense layout [V1, len, d] (oldest first) replayed from the log (tests / save_pe; needs V1 * len * d * 4 bytes)."""
        self.check_errors()
        snap = self.base.clone()
        out = torch.empty((self.V1, self.len, self.d), dtype=torch.float32, device=self.dev)
        cnt = self.ev_cnt.cpu().tolist()
        for f in range(self.len):
            slot = (self.head + f) % self.T
            n = cnt[slot]
            if n:
                snap[self.ev_node[slot, :n].long()] = self.ev_row[slot, :n]
            out[:, f, :] = snap
        return out

    def check_errors(self):
        samp = self.model.neighbor_sampler
        flag = int(samp._err.item())
        if flag & _lib.FLAG_CHANGELOG_FULL:
            samp._err.zero_()
            raise _lib.LstepError(f"a step changed more rows than the change-log history's event capacity ({self.cap})")
        samp.check_errors()

    def _run(self, lo, n_edges, ids, tmax, queries, out, batch_idx):
        lib = _lib.load()
        m = self.model
        T = self.T
        bi = self.batch_idx if batch_idx is None else batch_idx
        bmask = min(max(bi, 0), T) if self.len < T else T
        C = len(queries)
        with torch.cuda.device(self.dev), torch.no_grad():
            G = m._collapsed_filter(bmask, False)
            ws = self._workspace(ids.shape[0], n_edges, C)
            qptrs = (C_void_p * max(C, 1))(*[q.data_ptr() for q in queries])
            _lib.check(lib.lstep_pe_step_changelog(self._desc_ref, ctypes.byref(self.cl), m.neighbor_sampler.csr_ref, lo, n_edges, _lib.ptr(ids),
                                                   ids.shape[0], float(tmax), self.head, self.len, _lib.ptr(G), qptrs, C, 0, -1, _lib.ptr(out),
                                                   self.K, m._mlp_ref("nbr"), m._mlp_ref("update"), _lib.ptr(ws), ws.numel(),
                                                   _lib.ptr(m.neighbor_sampler._err), _lib.stream_ptr(), 0), "lstep_pe_step_changelog")
            if self.len < T:
                self.len += 1
            else:
                self.head = (self.head + 1) % T
        self.batch_idx = bi + 1
        self.steps_done += 1

    def run(self, b0: int, n_steps: int, queries, out: torch.Tensor = None, check: bool = False):
        lo0 = self.batch_lo[b0]
        last = None
        for i in range(n_steps):  # (no native multi-step call for this history: one C call per step)
            lo, hi, _, _ = self.batch_arrays(b0 + i)
            last = self.step(b0 + i, [q[lo - lo0:hi - lo0] for q in queries], None if out is None else out[i])
        if check:
            self.check_errors()
        return out if out is not None else last

    def step_host_async(self, *a, **k):
        raise NotImplementedError("host-fed steps run on the dense-ring PEStream")

    run_host = step_host = step_host_async
