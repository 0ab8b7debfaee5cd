"""Host -> device staging for the small per-batch arrays (ids, times) the reference API hands
over as numpy. All arrays of one call are packed into one pinned slot and moved with a single
async copy; a ring of slots with events keeps the host from overwriting a slot whose copy is
still in flight."""
from __future__ import annotations

import numpy as np
import torch

_ALIGN = 16


class Stager:
    def __init__(self, device, slots: int = 8, initial_bytes: int = 1 << 16):
        self.device = torch.device(device)
        self.slots = slots
        self._host = [None] * slots
        self._events = [None] * slots
        self._cap = initial_bytes
        self._i = 0
        self.bytes_moved = 0

    def upload(self, arrays):
        """arrays: list of (ndarray, np.dtype). Returns a list of device tensors (views into one
        device buffer that stays alive as long as the views do)."""
        sizes, total = [], 0
        conv = []
        for a, dt in arrays:
            a = np.ascontiguousarray(a, dtype=dt)
            conv.append(a)
            total = (total + _ALIGN - 1) // _ALIGN * _ALIGN
            sizes.append((total, a.nbytes))
            total += a.nbytes
        total = max(total, _ALIGN)
        i = self._i
        self._i = (i + 1) % self.slots
        if self._events[i] is not None:
            self._events[i].synchronize()
        if self._host[i] is None or self._host[i].numel() < total:
            cap = max(self._cap, 1 << (total - 1).bit_length())
            self._host[i] = torch.empty(cap, dtype=torch.uint8, pin_memory=True)
        host = self._host[i]
        hview = host.numpy()
        for a, (off, nb) in zip(conv, sizes):
            if nb:
                hview[off:off + nb] = a.view(np.uint8).reshape(-1)
        dev = torch.empty(total, dtype=torch.uint8, device=self.device)
        dev.copy_(host[:total], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._events[i] = ev
        self.bytes_moved += total
        out = []
        for a, (off, nb) in zip(conv, sizes):
            tdt = {np.dtype(np.int64): torch.int64, np.dtype(np.float64): torch.float64, np.dtype(np.float32): torch.float32,
                   np.dtype(np.int32): torch.int32}[a.dtype]
            out.append(dev[off:off + nb].view(tdt) if nb else torch.empty(0, dtype=tdt, device=self.device))
        return out
