"""Peer group: the scale-out PE step over NVLink peer memory (BASELINE config 5: 10 M nodes on 2 / 4 / 8 B200s).

One process per GPU. Every rank keeps a replica of the current PE table and of the temporal CSR and OWNS the nodes v with
v % world == rank: their PE history (change log) and every piece of the step whose result is a row of an owned node —
1 / world of the DFT filter, of the a6 query rows, of update_pe's phase A and phase B. Owners copy the rows they change, as
contiguous blocks, into small fixed inbox regions of the other ranks through peer pointers (CUDA IPC over NVLink / NVSwitch);
every rank scatters its inbox into its replica right after the first of the two flag barriers of a step (also in peer memory).
No NCCL call, no host synchronisation and no data-dependent message size is on the step's path. The protocol and the argument why two barriers suffice are in csrc/peer.cu; the C entry points
are lstep_pe_step_peer / lstep_pe_steps_peer (include/lstep_b200.h).

    PeerRank         this rank's state: table replica, change log of the owned nodes, per-batch plan, peer-shared buffers
    PeerLocalGroup   all ranks of a group in ONE process on one device (tests): peer pointers are plain device pointers and the
                     step is issued phase by phase for every rank
    connect_ipc      one process per GPU (torchrun): exchanges the CUDA IPC handles through torch.distributed and opens them
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib
from .model import LSTEP
from .sampler import NeighborSampler


class _DevMem:
    """A cudaMalloc'ed block (lstep_ipc_alloc) viewed as a torch tensor through __cuda_array_interface__."""

    def __init__(self, lib, nbytes: int):
        p = ctypes.c_void_p()
        _lib.check(lib.lstep_ipc_alloc(int(nbytes), ctypes.byref(p)), "ipc_alloc")
        self.lib, self.ptr, self.nbytes = lib, int(p.value), int(nbytes)

    def tensor(self, shape, dtype, device):
        typestr = {torch.float32: "<f4", torch.int32: "<i4", torch.uint8: "|u1"}[dtype]

        class _View:
            pass

        v = _View()
        v.__cuda_array_interface__ = {"shape": tuple(int(x) for x in shape), "typestr": typestr, "data": (self.ptr, False), "version": 2,
                                      "strides": None}
        v._keep = self
        return torch.as_tensor(v, device=device)

    def handle(self) -> bytes:
        h = (ctypes.c_ubyte * 64)()
        _lib.check(self.lib.lstep_ipc_export(ctypes.c_void_p(self.ptr), h), "ipc_export")
        return bytes(h)

    def free(self):
        if self.ptr:
            self.lib.lstep_ipc_free(ctypes.c_void_p(self.ptr))
            self.ptr = 0


def peer_batch_plan(src: np.ndarray, dst: np.ndarray, batch_size: int, world: int, rank: int) -> dict:
    """Host-side plan of a resident edge stream for one rank of a peer group (pure numpy: CPU-testable). Per batch b of
    `batch_size` edges: `ids` = the sorted unique batch nodes (replicated on every rank), `mine` = the ones this rank owns
    (v % world == rank), `pos` = their positions in `ids` (where the owner stores a node's row in the peers' buffers), and the
    share [q_off, q_off + q_rows) of the batch's edges whose a6 queries this rank computes. The owners' lists partition `ids`, the
    shares partition the batch. Returns the concatenated lists with their offsets."""
    n = len(src)
    nb = (n + batch_size - 1) // batch_size
    ids_l, mine_l, pos_l, share = [], [], [], []
    ids_off, mine_off = [0], [0]
    for b in range(nb):
        lo, hi = b * batch_size, min((b + 1) * batch_size, n)
        ids = np.unique(np.concatenate([src[lo:hi], dst[lo:hi]]))
        pos = np.nonzero(ids % world == rank)[0].astype(np.int64)
        ids_l.append(ids)
        mine_l.append(ids[pos])
        pos_l.append(pos)
        ids_off.append(ids_off[-1] + len(ids))
        mine_off.append(mine_off[-1] + len(pos))
        m = hi - lo
        share.append((rank * m // world, (rank + 1) * m // world - rank * m // world))
    cat = lambda xs: np.concatenate(xs).astype(np.int64) if xs else np.zeros(0, np.int64)
    return dict(num_batches=nb, ids=cat(ids_l), mine=cat(mine_l), pos=cat(pos_l), ids_off=np.asarray(ids_off, np.int64),
                mine_off=np.asarray(mine_off, np.int64), share=share)


class PeerRank:
    """Rank `rank` of `world` (see the module docstring). src / dst / t: the whole edge stream (replicated, device or host
    arrays); the per-batch plan (sorted unique batch nodes, the owned ones and their positions) of the resident stream
    [start, stop) is computed once, like PEStream's."""

    def __init__(self, model: LSTEP, rank: int, world: int, src, dst, t, num_nodes: int, batch_size: int, num_neighbors: int,
                 initial_pe: torch.Tensor, start: int = 0, stop: int = None, device=None, sampler: NeighborSampler = None,
                 event_capacity: int = None, timeout_ms: int = 4000):
        assert 1 <= world <= 16 and 0 <= rank < world
        self.m, self.rank, self.G = model, int(rank), int(world)
        self.dev = dev = torch.device(device) if device is not None else model._dev()
        self.B, self.K, self.T = int(batch_size), int(num_neighbors), model.num_fft_batches
        self.d, self.t_dim = model.pe_dim, model.time_feat_dim
        self.V1 = int(num_nodes) + 1
        self.timeout_ms = int(timeout_ms)
        if self.T > 128:
            raise _lib.LstepError("the change-log history supports T <= 128 window steps")
        lib = self.lib = _lib.load()
        as_dev = lambda a, dt: (a.to(dev, dt) if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a)).to(dev, dt))
        T, d, G, B = self.T, self.d, self.G, self.B
        with torch.cuda.device(dev):
            self.src, self.dst, self.tt = as_dev(src, torch.int64), as_dev(dst, torch.int64), as_dev(t, torch.float64)
            E = int(self.src.shape[0])
            stop = E if stop is None else int(stop)
            self.start, self.stop = int(start), stop
            if sampler is None:
                eid = torch.arange(1, E + 1, dtype=torch.int64, device=dev)
                sampler = NeighborSampler.from_edges(self.src, self.dst, eid, self.tt, "recent", device=dev, num_rows=self.V1)
                del eid
            self.sampler = model.neighbor_sampler = sampler
            # ---- memory the other ranks WRITE (cudaMalloc + CUDA IPC; small fixed regions): phase-A row buffer, filtered-row buffer,
            # inbox (one block per source rank for the rows it changed in a step), barrier flags
            self.rows_local = (self.V1 - self.rank + G - 1) // G
            # (the same on every rank: the inbox blocks are laid out by it)
            cap = event_capacity if event_capacity is not None else min((self.V1 + G - 1) // G, 2 * (2 * B * (self.K + 1) + 2) // G + 1024)
            self.cap = cap = int(max(cap, 1))
            H = 1
            while H < 2 * cap:
                H *= 2
            self._mem_rows = _DevMem(lib, 2 * B * d * 4)
            self._mem_filt = _DevMem(lib, 2 * B * d * 4)
            self._mem_inbox = _DevMem(lib, max(int(lib.lstep_peer_inbox_bytes(G, cap, d)), 256))
            self._mem_flags = _DevMem(lib, 16 * 4)
            self.cur = torch.empty((self.V1, d), dtype=torch.float32, device=dev)
            self.cur.copy_(initial_pe.to(dev, torch.float32))
            self.new_rows = self._mem_rows.tensor((2 * B, d), torch.float32, dev)
            # ---- change log of the owned nodes
            self.base = self.cur[self.rank::G].contiguous().clone()
            self.ev_node = torch.zeros((T, cap), dtype=torch.int32, device=dev)
            self.ev_row = torch.empty((T, cap, d), dtype=torch.float32, device=dev)
            self.ev_cnt = torch.zeros(T, dtype=torch.int32, device=dev)
            self.ev_hash = torch.full((T, H), -1, dtype=torch.int64, device=dev)
            self.ev_mask = torch.zeros((max(self.rows_local, 1), 4), dtype=torch.int32, device=dev)
            self.cl = _lib.ChangeLog(self.base.data_ptr(), self.ev_node.data_ptr(), self.ev_row.data_ptr(), self.ev_cnt.data_ptr(),
                                     self.ev_hash.data_ptr(), self.ev_mask.data_ptr(), max(self.rows_local, 1), T, cap, H, d, G, self.rank)
            self.desc = _lib.PEStreamDesc(self.src.data_ptr(), self.dst.data_ptr(), self.tt.data_ptr(), None, self.cur.data_ptr(), self.V1, T, d)
            # ---- per-batch plan of the resident stream: one pass on the host, uploaded once
            src_h = self.src[self.start:stop].cpu().numpy()
            dst_h = self.dst[self.start:stop].cpu().numpy()
            t_h = self.tt[self.start:stop].cpu().numpy()
            self.num_batches = nb = (stop - self.start + B - 1) // B
            plan = peer_batch_plan(src_h, dst_h, B, G, self.rank)
            assert plan["num_batches"] == nb
            lo_l, n_l, tmax = [], [], []
            for b in range(nb):
                lo, hi = b * B, min((b + 1) * B, stop - self.start)
                lo_l.append(self.start + lo)
                n_l.append(hi - lo)
                tmax.append(float(t_h[lo:hi].max()))
            if len(plan["ids"]) and int(plan["ids"].max()) >= self.V1:
                raise IndexError("edge stream holds a node id outside the PE table")
            up = lambda a: torch.from_numpy(a).to(dev)
            self.ids, self.ids_mine, self.pos_mine = up(plan["ids"]), up(plan["mine"]), up(plan["pos"])
            self.ids_off, self.mine_off = plan["ids_off"], plan["mine_off"]
            self.lo, self.n_edges, self.tmax = np.asarray(lo_l, np.int64), np.asarray(n_l, np.int64), np.asarray(tmax, np.float64)
            need = lib.lstep_pe_step_workspace_bytes(2 * B, B, 8, self.K, d, self.t_dim, self.V1)
            self.ws = torch.empty(need + 4096, dtype=torch.uint8, device=dev)
            _lib.check(lib.lstep_update_pe_workspace_init(_lib.ptr(self.ws), self.ws.numel(), self.V1, _lib.stream_ptr()), "ws init")
        self.head, self.len, self.batch_idx = 0, 1, 0
        self.epoch = ctypes.c_uint32(0)
        self.grp = None
        self._opened = []

    # ---- group wiring ---------------------------------------------------------------------------------------------------
    _SHARED = ("_mem_rows", "_mem_filt", "_mem_inbox", "_mem_flags")

    def set_group(self, rows, filts, inboxes, flags):
        """Device pointers (ints) of every rank's new_rows buffer / filt buffer / inbox / flag block, indexed by rank."""
        g = _lib.PeerGroup()
        g.rank, g.world, g.cap = self.rank, self.G, self.cap
        for i in range(self.G):
            g.new_rows[i], g.filt[i], g.inbox[i], g.flags[i] = rows[i], filts[i], inboxes[i], flags[i]
        g.table[self.rank] = self.cur.data_ptr()
        self.grp = g

    def local_ptrs(self):
        return tuple(getattr(self, m).ptr for m in self._SHARED)

    def connect_ipc(self, group=None):
        """One process per GPU: exchange the CUDA IPC handles of the four peer-written blocks through torch.distributed and open
        the other ranks' (cudaIpcOpenMemHandle enables peer access over NVLink)."""
        import torch.distributed as dist
        mine = (self.rank,) + tuple(getattr(self, m).handle() for m in self._SHARED)
        allh = [None] * self.G
        if self.G > 1:
            dist.all_gather_object(allh, mine, group=group)
        else:
            allh = [mine]
        ptrs = [[0] * self.G for _ in range(len(self._SHARED))]
        for r, *hs in allh:
            for k, h in enumerate(hs):
                if r == self.rank:
                    ptrs[k][r] = self.local_ptrs()[k]
                else:
                    p = ctypes.c_void_p()
                    buf = (ctypes.c_ubyte * 64).from_buffer_copy(h)
                    with torch.cuda.device(self.dev):
                        _lib.check(self.lib.lstep_ipc_open(buf, ctypes.byref(p)), "ipc_open")
                    self._opened.append(int(p.value))
                    ptrs[k][r] = int(p.value)
        self.set_group(*ptrs)
        if self.G > 1:
            dist.barrier(group=group)

    def close(self):
        torch.cuda.synchronize(self.dev)
        for p in self._opened:
            self.lib.lstep_ipc_close(ctypes.c_void_p(p))
        self._opened = []

    # ---- the step -------------------------------------------------------------------------------------------------------
    def batch(self, b: int):
        return int(self.lo[b]), int(self.lo[b] + self.n_edges[b])

    def share(self, b: int):
        """(q_off, q_rows): this rank's share of the batch's edges for the a6 queries."""
        n = int(self.n_edges[b])
        return self.rank * n // self.G, (self.rank + 1) * n // self.G - self.rank * n // self.G

    def step(self, b: int, queries, out: torch.Tensor = None, phases: int = 7) -> torch.Tensor:
        """One step on batch b. queries: device int64 tensors with the ids of ALL edges of the batch (this rank reads its
        share). Returns [C, q_rows, d]. phases: see lstep_pe_step_peer (7 = whole step; the state advances with bit 2)."""
        assert self.grp is not None, "set_group / connect_ipc first"
        lib, T, d = self.lib, self.T, self.d
        lo, hi = self.batch(b)
        q_off, q_rows = self.share(b)
        C = len(queries)
        i0, i1 = int(self.ids_off[b]), int(self.ids_off[b + 1])
        m0, m1 = int(self.mine_off[b]), int(self.mine_off[b + 1])
        bmask = min(max(self.batch_idx, 0), T) if self.len < T else T
        with torch.cuda.device(self.dev), torch.no_grad():
            Gt = self.m._collapsed_filter(bmask, False)
            if out is None:
                out = torch.empty((max(C, 1), q_rows, d), dtype=torch.float32, device=self.dev)
            qptrs = (ctypes.c_void_p * max(C, 1))(*[q.data_ptr() + 8 * q_off for q in queries])
            i64p = lambda t_, o: ctypes.c_void_p(t_.data_ptr() + 8 * o)
            _lib.check(lib.lstep_pe_step_peer(ctypes.byref(self.desc), ctypes.byref(self.cl), self.sampler.csr_ref, ctypes.byref(self.grp), lo, hi - lo,
                                              i64p(self.ids, i0), i1 - i0, i64p(self.ids_mine, m0), i64p(self.pos_mine, m0), m1 - m0,
                                              float(self.tmax[b]), self.head, self.len, _lib.ptr(Gt), qptrs, C, q_off, q_rows, _lib.ptr(out), self.K,
                                              self.m._mlp_ref("nbr"), self.m._mlp_ref("update"), _lib.ptr(self.ws), self.ws.numel(),
                                              _lib.ptr(self.sampler._err), self.epoch.value, self.timeout_ms, phases, _lib.stream_ptr()),
                       "lstep_pe_step_peer")
        if phases & 4:
            self._advance(1)
        return out

    def _advance(self, n: int):
        for _ in range(n):
            if self.len < self.T:
                self.len += 1
            else:
                self.head = (self.head + 1) % self.T
        self.batch_idx += n
        self.epoch = ctypes.c_uint32((self.epoch.value + 2 * n) & 0xffffffff)

    def run(self, b0: int, n_steps: int, query_arrays, out: torch.Tensor = None, out_step_stride: int = 0):
        """n_steps consecutive steady-state steps (full history) starting at batch b0 in ONE native call (lstep_pe_steps_peer).
        query_arrays: device int64 tensors indexed by GLOBAL edge position (query set c of a step = query_arrays[c][lo:lo+n]).
        out: [C, max q_rows, d] (every step overwrites it) or, with out_step_stride, one block per step."""
        assert self.grp is not None and self.len == self.T and b0 + n_steps <= self.num_batches
        lib, d = self.lib, self.d
        C = len(query_arrays)
        if out is None:
            out = torch.empty((max(C, 1), self.B // self.G + 1, d), dtype=torch.float32, device=self.dev)
        qptrs = (ctypes.c_void_p * max(C, 1))(*[q.data_ptr() for q in query_arrays])
        ap = lambda a, o: a[o:].ctypes.data_as(ctypes.c_void_p)
        head, epoch = ctypes.c_int(self.head), ctypes.c_uint32(self.epoch.value)
        ids_off = np.ascontiguousarray(self.ids_off[b0:b0 + n_steps + 1])
        mine_off = np.ascontiguousarray(self.mine_off[b0:b0 + n_steps + 1])
        lo = np.ascontiguousarray(self.lo[b0:b0 + n_steps])
        with torch.cuda.device(self.dev), torch.no_grad():
            Gt = self.m._collapsed_filter(self.T, False)
            rc = lib.lstep_pe_steps_peer(ctypes.byref(self.desc), ctypes.byref(self.cl), self.sampler.csr_ref, ctypes.byref(self.grp), n_steps,
                                         ap(lo, 0), ap(self.n_edges, b0), ap(self.tmax, b0), _lib.ptr(self.ids), ap(ids_off, 0),
                                         _lib.ptr(self.ids_mine), _lib.ptr(self.pos_mine), ap(mine_off, 0), ctypes.byref(head), _lib.ptr(Gt), qptrs,
                                         ap(lo, 0), C, _lib.ptr(out), int(out_step_stride), self.K, self.m._mlp_ref("nbr"),
                                         self.m._mlp_ref("update"), _lib.ptr(self.ws), self.ws.numel(), _lib.ptr(self.sampler._err),
                                         ctypes.byref(epoch), self.timeout_ms, _lib.stream_ptr())
        done = ((epoch.value - self.epoch.value) & 0xffffffff) // 2
        self.head, self.batch_idx = head.value, self.batch_idx + done
        self.epoch = epoch
        _lib.check(rc, "lstep_pe_steps_peer")
        return out

    def barrier(self):
        """Barrier + application of the inbox: once the kernels this enqueues have run, this replica holds the rows every rank changed
        in the last step."""
        e = (self.epoch.value + 1) & 0xffffffff
        with torch.cuda.device(self.dev):
            _lib.check(self.lib.lstep_peer_sync_tables(ctypes.byref(self.grp), e, self.timeout_ms, _lib.ptr(self.sampler._err), self.d, self.V1,
                                                       _lib.stream_ptr()), "peer_sync_tables")
        self.epoch = ctypes.c_uint32((self.epoch.value + 2) & 0xffffffff)

    def check_errors(self):
        flag = int(self.sampler._err.item())
        if flag & _lib.FLAG_PEER_TIMEOUT:
            self.sampler._err.zero_()
            raise _lib.LstepError("a peer GPU did not reach a step barrier in time (peer group step)")
        if flag & _lib.FLAG_CHANGELOG_FULL:
            self.sampler._err.zero_()
            raise _lib.LstepError(f"a step changed more owned rows than the change-log history's event capacity ({self.cap})")
        self.sampler.check_errors()

    def history_bytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in (self.base, self.ev_node, self.ev_row, self.ev_cnt, self.ev_hash, self.ev_mask))

    def export_history_rows(self) -> torch.Tensor:
        """[rows_local, len, d], oldest first: row l = node l * G + rank."""
        snap = self.base.clone()
        out = torch.empty((self.rows_local, self.len, self.d), dtype=torch.float32, device=self.dev)
        cnt = self.ev_cnt.cpu().tolist()
        for f in range(self.len):
            slot = (self.head + f) % self.T
            if cnt[slot]:
                snap[(self.ev_node[slot, :cnt[slot]].long() - self.rank) // self.G] = self.ev_row[slot, :cnt[slot]]
            out[:, f, :] = snap[:self.rows_local]
        return out


class PeerLocalGroup:
    """All ranks of a group in one process on one device (tests): the peer pointers are the rank states' own device pointers
    and every phase of a step is issued for all ranks before the next one (the barriers are then already satisfied)."""

    def __init__(self, ranks):
        self.ranks, self.G = ranks, len(ranks)
        assert len({rk.cap for rk in ranks}) == 1
        ptrs = [[rk.local_ptrs()[k] for rk in ranks] for k in range(len(PeerRank._SHARED))]
        for rk in ranks:
            rk.set_group(*ptrs)

    def step(self, b: int, queries):
        outs = [None] * self.G
        for ph in (1, 2, 4):
            for i, rk in enumerate(self.ranks):
                o = rk.step(b, queries, out=outs[i], phases=ph)
                outs[i] = o
        return outs

    def barrier(self):
        for rk in self.ranks:
            e = (rk.epoch.value + 1) & 0xffffffff
            _lib.check(rk.lib.lstep_peer_signal(ctypes.byref(rk.grp), e, _lib.stream_ptr()), "peer_signal")
        for rk in self.ranks:
            rk.barrier()
