"""Import shim for the un-vendored `torch_scatter` dependency of the reference
(models/LSTEP.py:10). Only used by tests/golden/make_golden.py to run the unmodified reference
on CPU. Semantics follow torch_scatter's documented behaviour: `scatter(..., reduce='sum')` is
`out.scatter_add_` along `dim`; `scatter_mean` divides the sums by the clamped counts."""
import torch


def _expand(index, src, dim):
    if dim < 0:
        dim += src.dim()
    if index.dim() == 1:
        for _ in range(dim):
            index = index.unsqueeze(0)
    while index.dim() < src.dim():
        index = index.unsqueeze(-1)
    return index.expand_as(src), dim


def scatter(src, index, dim=-1, out=None, dim_size=None, reduce="sum"):
    index, dim = _expand(index, src, dim)
    if out is None:
        size = list(src.shape)
        size[dim] = int(dim_size if dim_size is not None else (int(index.max()) + 1 if index.numel() else 0))
        out = torch.zeros(size, dtype=src.dtype, device=src.device)
    if reduce in ("sum", "add"):
        return out.scatter_add_(dim, index, src)
    raise NotImplementedError(reduce)


def scatter_mean(src, index, dim=-1, out=None, dim_size=None):
    out = scatter(src, index, dim, out, dim_size, "sum")
    idx, d = _expand(index, src, dim)
    cnt = torch.zeros_like(out).scatter_add_(d, idx, torch.ones_like(src)).clamp_(min=1)
    return out.div_(cnt)
