"""Multi-GPU check of the sharded path over NCCL (run with torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/run_sharded_nccl_check.py

Every rank also runs the single-GPU PEStream on the same small graph and compares the rows it owns and
the (all-reduced) neighbourhood outputs. Prints one line per rank and exits non-zero on mismatch."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")]
from common import golden_path, pe_close, seeded_normal, update_error_report  # noqa: E402
from harness import build_dropin  # noqa: E402
from lstep_b200 import DistGroup, NeighborSampler, PEStream, ShardedPEStream, ShardRank, synth  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    tag = "full"
    z = np.load(golden_path(f"replay_{tag}.npz"))
    d, T, K, t_dim, F, B = (int(z[k]) for k in ("pe_dim", "T", "K", "time_dim", "feat_dim", "B"))
    V, E, e0 = int(z["V"]), int(z["E"]), int(z["e0"])
    g = synth.make_graph("tiny", seed=int(z["graph_seed"]), num_nodes=V, num_edges=E)
    s = NeighborSampler.from_edges(g.src_node_ids, g.dst_node_ids, g.edge_ids, g.node_interact_times, "recent", num_rows=V + 1)
    lstep = build_dropin(tag, g, s, F, d, t_dim, T, K)[0].eval()
    hist0 = seeded_normal(int(z["hist0_seed"]), (V + 1, 1, d), 0.3)
    hist0[0] = 0
    init = torch.from_numpy(hist0[:, 0, :]).cuda()
    single = PEStream(lstep, g.src_node_ids, g.dst_node_ids, g.node_interact_times, B, K, initial_pe=init, start=e0)
    rk = ShardRank(lstep, rank, world, g.src_node_ids, g.dst_node_ids, g.node_interact_times, g.edge_ids, V, B, K, init, start=e0)
    sh = ShardedPEStream(rk, DistGroup())
    worst_out = 0.0
    for b in range(min(single.num_batches, 80)):
        lo, hi, _, _ = single.batch_arrays(b)
        nd = z["neg_dst"][b][:hi - lo].astype(np.int64)
        want = single.step(b, [single.src[lo:hi], single.dst[lo:hi], torch.from_numpy(nd).cuda()])
        got = sh.step(b, [g.src_node_ids[lo:hi], g.dst_node_ids[lo:hi], nd])
        ok, w = pe_close(got.cpu().numpy(), want.cpu().numpy(), 1e-4)
        worst_out = max(worst_out, w)
        assert ok, (rank, b, w)
    mine, ref = rk.owned_table().cpu().numpy(), single.cur[rank::world].cpu().numpy()
    rep = update_error_report(mine, ref)
    assert rep["max"] <= 1e-4, (rank, rep)  # same bar as the emulated-ranks test (float partial rows combined per rank)
    print(f"rank {rank}/{world}: sharded == single GPU (outputs worst {worst_out:.2e}); X1 {rk.bytes_x1 / 1e6:.1f} MB, X2 {rk.bytes_x2 / 1e6:.1f} MB",
          flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
