"""A few streaming PE steps (tiny graph with timestamp ties + a Reddit-shaped graph of 60 k edges, full history T = 100; B = 200 and
one B = 2000 step so that both MLP kernels run) for compute-sanitizer:

    compute-sanitizer --tool memcheck  python profiles/sanitize_step.py          > profiles/r02_sanitizer_memcheck.txt
    compute-sanitizer --tool racecheck python profiles/sanitize_step.py          > profiles/r02_sanitizer_racecheck.txt
    ... python profiles/sanitize_step.py --no-pdl        (plain stream launches)

One tool per GPU call (B200_PROFILING.md)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")]
from common import seeded_normal  # noqa: E402
from harness import build_dropin  # noqa: E402
from lstep_b200 import NeighborSampler, PEStream, _lib, synth  # noqa: E402

lib = _lib.load()
if "--no-pdl" in sys.argv:
    lib.lstep_set_option(b"pdl", 0)
for gname, n_edges, B, steps in (("tiny_ties", None, 40, 4), ("reddit", 60_000, 200, 3), ("reddit", 60_000, 2000, 2)):
    g = synth.make_graph(gname, seed=0, num_edges=n_edges)
    V, d, T, K = g.num_nodes, 172, 100, 20
    s = NeighborSampler.from_edges(g.src_node_ids, g.dst_node_ids, g.edge_ids, g.node_interact_times, "recent", num_rows=V + 1)
    lstep = build_dropin("full", g, s, 172, d, 100, T, K)[0].eval()
    hist = torch.from_numpy(seeded_normal(3, (V + 1, T, d), 0.3)).cuda()
    e0 = g.num_edges - steps * B
    st = PEStream(lstep, g.src_node_ids, g.dst_node_ids, g.node_interact_times, B, K, history=hist, start=e0)
    neg = torch.from_numpy(np.random.default_rng(0).integers(1, V + 1, g.num_edges - e0).astype(np.int64)).cuda()
    for b in range(st.num_batches):
        lo, hi, _, _ = st.batch_arrays(b)
        out = st.step(b, [st.src[lo:hi], st.dst[lo:hi], st.src[lo:hi], neg[lo - e0:hi - e0]])
    st.check_errors()
    torch.cuda.synchronize()
    print(gname, "B", B, "steps", st.num_batches, "checksum", float(st.cur.double().abs().sum()), flush=True)
    # host-fed path too
    q = [g.src_node_ids[e0:e0 + B], g.dst_node_ids[e0:e0 + B]]
    st.step_host(g.src_node_ids[e0:e0 + B], g.dst_node_ids[e0:e0 + B], g.node_interact_times[e0:e0 + B], q)
    torch.cuda.synchronize()
print("done")
