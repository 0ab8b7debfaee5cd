"""Peer-store bandwidth of the row-publication kernel (csrc/peer.cu::peer_rows_bcast_kernel) between two GPUs of one box, next
to a device-to-device cudaMemcpy and an NCCL all-gather of the same bytes. Run with torchrun, 2 ranks:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 profiles/peer_bw.py
"""
import ctypes
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lstep_b200 import _lib  # noqa: E402
from lstep_b200.peer import _DevMem  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    d, R = 172, int(os.environ.get('PEER_BW_ROWS', 1 << 20))
    mem = _DevMem(lib, R * d * 4)
    rows = _DevMem(lib, 4096)
    flags = _DevMem(lib, 64)
    hs = [None] * world
    dist.all_gather_object(hs, (rank, mem.handle(), rows.handle(), flags.handle()))
    g = _lib.PeerGroup()
    g.rank, g.world = rank, world
    for r, h0, h1, h2 in hs:
        for k, h in enumerate((h0, h1, h2)):
            if r == rank:
                p = (mem.ptr, rows.ptr, flags.ptr)[k]
            else:
                pp = ctypes.c_void_p()
                _lib.check(lib.lstep_ipc_open((ctypes.c_ubyte * 64).from_buffer_copy(h), ctypes.byref(pp)), "open")
                p = pp.value
            (g.table, g.new_rows, g.flags)[k][r] = p
    src = torch.randn((min(R, 1 << 20), d), device=dev)
    perm = torch.randperm(R, device=dev)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    peer_mask = ((1 << world) - 1) & ~(1 << rank)
    res = {}
    for n in (1000, 4000, 24000, 100000, 1000000):
        for name, mask, idx in (("peer_contig", peer_mask, None), ("peer_scattered", peer_mask, perm[:n].contiguous()), ("self_scattered", 1 << rank, perm[:n].contiguous())):
            ts = []
            for it in range(6):
                dist.barrier()
                torch.cuda.synchronize()
                ev0.record()
                _lib.check(lib.lstep_peer_rows_bcast(_lib.ptr(src), n, d, _lib.ptr(idx), ctypes.byref(g), 0, mask, _lib.stream_ptr()), "bcast")
                ev1.record()
                torch.cuda.synchronize()
                ts.append(ev0.elapsed_time(ev1))
            t = sorted(ts[1:])[len(ts[1:]) // 2]
            res[(name, n)] = (t * 1e3, n * d * 4 / (t * 1e-3) / 1e9)
    # NCCL all-gather of the same bytes
    for n in (4000, 24000, 1000000):
        out = torch.empty((world * n, d), device=dev)
        ts = []
        for it in range(6):
            dist.barrier()
            torch.cuda.synchronize()
            ev0.record()
            dist.all_gather_into_tensor(out, src[:n])
            ev1.record()
            torch.cuda.synchronize()
            ts.append(ev0.elapsed_time(ev1))
        t = sorted(ts[1:])[len(ts[1:]) // 2]
        res[("nccl_all_gather", n)] = (t * 1e3, n * d * 4 / (t * 1e-3) / 1e9)
    if rank == 0:
        for k, (us, gbs) in res.items():
            print(f"{k[0]:>16} rows={k[1]:>8}  {us:9.1f} us  {gbs:8.1f} GB/s per destination", flush=True)
    dist.barrier()
    torch.cuda.synchronize()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
