"""Run bracket of the PE path: the initial positional encoding and the PE-history checkpoint files.

    LaplacianPE(edge_index, num_nodes, k)          utils/PositionalEncoding.py:42-61
    RandomWalkPE(edge_index, num_nodes, walk_length)  utils/PositionalEncoding.py:68-91
    save_pe(history, path) / load_pe(path)         utils/EarlyStopping.py:79-82, 100-104 (plain torch.save / torch.load of
                                                   the [V1, Th, d] history tensor: files are interchangeable both ways)

The reference computes the initial encoding ONCE per run from the edges of the FIRST training batch only
(train_LSTEP_link_prediction.py:168-189): `edge_index` holds <= 2B undirected edges over num_nodes = V1 nodes, so all but
a few hundred nodes are isolated. Both functions therefore work on the compacted block of non-isolated nodes (dense, on
the GPU, float64) and fill the isolated rows in closed form; no CPU ARPACK, no V1 x V1 matrix.

Numerical contract. RandomWalkPE is deterministic: pe[i, s] = (P^(s+1))[i, i] with P = D^-1 A (duplicate edges count
with their multiplicity) — matches the reference to fp32 rounding. LaplacianPE returns eigenvectors of the symmetric
normalised Laplacian for its k+1 smallest eigenvalues, the first dropped, each column multiplied by a random sign drawn
exactly as the reference draws it (`-1 + 2 * torch.randint(0, 2, (k,))` on the CPU generator). Eigenvectors are only
defined up to the basis of each eigenspace — isolated nodes make the eigenvalue 1 massively degenerate and ARPACK's
choice inside it is arbitrary — so parity is on the eigenvalues and on the invariant subspaces, not elementwise.
"""
from __future__ import annotations

import torch


def _device(edge_index, device):
    if device is not None:
        return torch.device(device)
    if edge_index.is_cuda:
        return edge_index.device
    if not torch.cuda.is_available():
        raise RuntimeError("lstep_b200.pe_init runs on a CUDA device (no CPU path)")
    return torch.device("cuda", torch.cuda.current_device())


def _block(edge_index, num_nodes, dev):
    """Compact the nodes that carry at least one edge: (sorted node ids [n], A [n, n] float64 with edge multiplicities)."""
    ei = edge_index.to(dev).long()
    nodes = torch.unique(ei)
    if nodes.numel() and (int(nodes.min()) < 0 or int(nodes.max()) >= num_nodes):
        raise IndexError("edge_index refers to a node outside [0, num_nodes)")
    n = nodes.numel()
    if n > 20000:
        raise RuntimeError(f"{n} non-isolated nodes: the dense block path is meant for the first-batch graph the reference uses")
    loc = torch.searchsorted(nodes, ei)
    A = torch.zeros((n, n), dtype=torch.float64, device=dev)
    A.index_put_((loc[0], loc[1]), torch.ones(ei.shape[1], dtype=torch.float64, device=dev), accumulate=True)
    return nodes, A


def RandomWalkPE(edge_index: torch.Tensor, num_nodes: int, walk_length: int, device=None) -> torch.Tensor:
    """[num_nodes, walk_length] float32: return probabilities of 1..walk_length step random walks (isolated nodes: 0)."""
    dev = _device(edge_index, device)
    nodes, A = _block(edge_index, num_nodes, dev)
    pe = torch.zeros((num_nodes, walk_length), dtype=torch.float32, device=dev)
    if nodes.numel() == 0:
        return pe
    deg = A.sum(dim=1).clamp(min=1.0)  # out-degree with multiplicity (tg_scatter of ones over `row`, clamp(min=1))
    P = (A / deg[:, None]).to(torch.float32)  # the reference works in float32 (`value = 1.0 / value` on a float32 tensor)
    out = P.clone()
    cols = [out.diagonal().clone()]
    for _ in range(walk_length - 1):
        out = out @ P
        cols.append(out.diagonal().clone())
    pe[nodes] = torch.stack(cols, dim=-1)
    return pe


def LaplacianPE(edge_index: torch.Tensor, num_nodes: int, k: int, device=None):
    """(pe [num_nodes, k] float32, edge_weight) — eigenvectors 1..k (by ascending eigenvalue) of L = I - D^-1/2 A D^-1/2 with
    the reference's random column signs; edge_weight = the Laplacian's entries in torch_geometric.get_laplacian order
    (off-diagonal entries of the non-loop edges, then one diagonal entry per node)."""
    dev = _device(edge_index, device)
    ei = edge_index.to(dev).long()
    keep = ei[0] != ei[1]  # get_laplacian removes self loops first
    nodes, A = _block(ei[:, keep], num_nodes, dev)
    n = nodes.numel()
    deg = A.sum(dim=1)
    dis = deg.pow(-0.5)
    dis[torch.isinf(dis)] = 0
    L = torch.eye(n, dtype=torch.float64, device=dev) - dis[:, None] * A * dis[None, :]
    # spectrum of the whole graph = spectrum of the block + eigenvalue 1 (unit vectors) for every isolated node
    ev_b, vec_b = torch.linalg.eigh(L) if n else (torch.zeros(0, dtype=torch.float64, device=dev), torch.zeros((0, 0), dtype=torch.float64, device=dev))
    n_iso = num_nodes - n
    iso_mask = torch.ones(num_nodes, dtype=torch.bool, device=dev)
    iso_mask[nodes] = False
    iso_nodes = torch.nonzero(iso_mask).flatten()
    take = min(k + 1, num_nodes)
    ev_all = torch.cat([ev_b, torch.ones(n_iso, dtype=torch.float64, device=dev)])
    order = torch.argsort(ev_all, stable=True)[:take]  # block eigenvalues first among ties
    vecs = torch.zeros((num_nodes, take), dtype=torch.float64, device=dev)
    for j, o in enumerate(order.tolist()):
        if o < n:
            vecs[nodes, j] = vec_b[:, o]
        else:
            vecs[iso_nodes[o - n], j] = 1.0
    pe = vecs[:, 1:k + 1].to(torch.float32)
    sign = -1 + 2 * torch.randint(0, 2, (pe.shape[1],))  # the reference's draw, on the CPU generator
    pe = pe * sign.to(dev)
    # edge_weight as get_laplacian(normalization='sym') returns it
    row, col = ei[0][keep], ei[1][keep]
    deg_full = torch.zeros(num_nodes, dtype=torch.float32, device=dev).index_add_(0, row, torch.ones(row.numel(), dtype=torch.float32, device=dev))
    dis_f = deg_full.pow(-0.5)
    dis_f[torch.isinf(dis_f)] = 0
    edge_weight = torch.cat([-(dis_f[row] * dis_f[col]), torch.ones(num_nodes, dtype=torch.float32, device=dev)])
    return pe, edge_weight


def save_pe(final_trained_positional_encoding: torch.Tensor, path: str):
    """utils/EarlyStopping.py:79-82 — the history [V1, Th, d] as a plain tensor file (`PEStream.export_history()` gives it)."""
    if final_trained_positional_encoding is not None:
        torch.save(final_trained_positional_encoding, path)


def load_pe(path: str, map_location=None) -> torch.Tensor:
    """utils/EarlyStopping.py:100-104; feed the result to PEStream(history=...) / PEStream.import_history()."""
    return torch.load(path, map_location=map_location)
