"""The seven `torch_geometric.utils` functions utils/PositionalEncoding.py imports, restated from
their documented behaviour (PyG is not installed offline). TEST INFRASTRUCTURE: lets the unmodified
reference's LaplacianPE / RandomWalkPE run on CPU as the checker of lstep_b200.pe_init."""
import numpy as np
import scipy.sparse as sp
import torch


def scatter(src, index, dim=0, dim_size=None, reduce="sum"):
    assert reduce in ("sum", "add") and dim == 0
    n = int(dim_size if dim_size is not None else (int(index.max()) + 1 if index.numel() else 0))
    out = torch.zeros((n,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    return out.index_add_(0, index, src)


def get_laplacian(edge_index, edge_weight=None, normalization=None, dtype=None, num_nodes=None):
    """L = D - A (None), I - D^-1/2 A D^-1/2 ('sym') or I - D^-1 A ('rw'); self loops removed first, the
    diagonal appended after the off-diagonal entries (PyG's order: edges, then one loop per node)."""
    row, col = edge_index[0], edge_index[1]
    keep = row != col
    row, col = row[keep], col[keep]
    if edge_weight is None:
        edge_weight = torch.ones(row.numel(), dtype=dtype or torch.float32, device=edge_index.device)
    else:
        edge_weight = edge_weight[keep]
    n = int(num_nodes if num_nodes is not None else int(edge_index.max()) + 1)
    deg = scatter(edge_weight, row, 0, dim_size=n)
    loop = torch.arange(n, device=row.device)
    if normalization is None:
        w = torch.cat([-edge_weight, deg])
    elif normalization == "sym":
        dis = deg.pow(-0.5)
        dis.masked_fill_(dis == float("inf"), 0)
        w = torch.cat([-(dis[row] * edge_weight * dis[col]), torch.ones(n, dtype=edge_weight.dtype, device=row.device)])
    else:
        dinv = 1.0 / deg
        dinv.masked_fill_(dinv == float("inf"), 0)
        w = torch.cat([-(dinv[row] * edge_weight), torch.ones(n, dtype=edge_weight.dtype, device=row.device)])
    return torch.stack([torch.cat([row, loop]), torch.cat([col, loop])]), w


def to_scipy_sparse_matrix(edge_index, edge_attr=None, num_nodes=None):
    row, col = edge_index.cpu().numpy()
    if edge_attr is None:
        edge_attr = torch.ones(row.shape[0])
    n = int(num_nodes if num_nodes is not None else int(edge_index.max()) + 1)
    return sp.coo_matrix((edge_attr.view(-1).cpu().numpy(), (row, col)), (n, n))  # duplicates are summed on conversion


def is_torch_sparse_tensor(x):
    return isinstance(x, torch.Tensor) and x.layout in (torch.sparse_coo, torch.sparse_csr, torch.sparse_csc)


def to_torch_csr_tensor(edge_index, edge_attr=None, size=None, is_coalesced=False):
    n = int(size if size is not None else int(edge_index.max()) + 1)
    if edge_attr is None:
        edge_attr = torch.ones(edge_index.shape[1])
    coo = torch.sparse_coo_tensor(edge_index, edge_attr, (n, n)).coalesce()  # sorts by (row, col), sums duplicates
    return coo.to_sparse_csr()


def to_edge_index(adj):
    if adj.layout == torch.sparse_csr:
        adj = adj.to_sparse_coo()
    adj = adj.coalesce()
    return adj.indices(), adj.values()


def get_self_loop_attr(edge_index, edge_attr=None, num_nodes=None):
    mask = edge_index[0] == edge_index[1]
    idx = edge_index[0][mask]
    attr = edge_attr[mask] if edge_attr is not None else torch.ones_like(idx, dtype=torch.float)
    n = int(num_nodes if num_nodes is not None else int(edge_index.max()) + 1)
    out = attr.new_zeros((n,) + tuple(attr.shape[1:]))
    out[idx] = attr
    return out
