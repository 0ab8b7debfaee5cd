"""Per-kernel timeline of one streaming step (build with LSTEP_NVCC_EXTRA=-DLSTEP_TIMELINE)."""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from lstep_b200 import NeighborSampler, PEStream, synth, _lib
lib = _lib.load()
dev = torch.device("cuda")
g = synth.make_graph("reddit", seed=0)
V1 = g.num_nodes + 1
s = NeighborSampler.from_edges(g.src_node_ids, g.dst_node_ids, g.edge_ids, g.node_interact_times, "recent", num_rows=V1)
m = bench.make_params_model(g, s, dev)
init = torch.from_numpy(synth.make_initial_pe(g.num_nodes, 172, seed=1)).to(dev)
B = 200
e0 = int(g.num_edges * 0.7) // B * B
st = PEStream(m, g.src_node_ids, g.dst_node_ids, g.node_interact_times, B, 20, initial_pe=init, start=e0)
neg = torch.from_numpy(np.random.default_rng(2).choice(np.unique(g.dst_node_ids), size=g.num_edges - e0).astype(np.int64)).to(dev)
def q(b):
    lo, hi, _, _ = st.batch_arrays(b)
    return [st.src[lo:hi], st.dst[lo:hi], st.src[lo:hi], neg[lo - e0:hi - e0]]
outs = torch.empty((4, B, 172), device=dev)
for b in range(130): st.step(b, q(b), outs)
torch.cuda.synchronize()
tus = ["dft", "step", "mlp", "push"]
fns = {t: getattr(lib, "lstep_debug_timeline_" + t) for t in tus}
names = {0: "dft", 1: "gather", 2: "mlp_pair", 3: "push", 4: "mlp_B", 5: "append", 6: "g:lookup", 7: "g:cos", 8: "g:waitret", 9: "g:start_q", 10: "g:start_e", 11: "g:exit_q", 12: "g:exit_e"}
def reset():
    for f in fns.values(): f(0, None)
def read():
    tl = {}
    for t, f in fns.items():
        buf = (ctypes.c_ulonglong * 64)()
        f(1, buf)
        for k in range(13):
            a, w, e = buf[4 * k], buf[4 * k + 1], buf[4 * k + 2]
            if e != 0: tl[k] = (a, w, e)
    return tl
rows = []
b = 130
for trial in range(12):
    # three steps back to back; the timeline records the LAST one only if we reset between: instead run 2 untimed, sync-free, then reset cannot be
    # stream-ordered, so: sync, reset, run ONE step, sync
    torch.cuda.synchronize(); reset()
    st.step(b, q(b), outs); b += 1
    torch.cuda.synchronize()
    tl = read()
    t0 = min(v[0] for k, v in tl.items() if k < 6)
    rows.append({k: tuple(((x - t0) / 1e3 if 0 < x < 2**63 else float('nan')) for x in v) for k, v in tl.items()})
med = {k: tuple(float(np.median([r[k][i] for r in rows if k in r])) for i in range(3)) for k in names}
print("kernel      entry   waited     exit   (us since first entry; single isolated step)")
for k in sorted(med): print(f"{names[k]:9s} {med[k][0]:8.1f} {med[k][1]:8.1f} {med[k][2]:8.1f}")
