"""Stand-in for torch_scatter 1.x (reference call site gcn_conv.py:4,66)."""
import torch


def scatter_add(src, index, dim=0, out=None, dim_size=None, fill_value=0):
    assert dim == 0 and out is None
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() else 0
    res = src.new_full((dim_size,) + tuple(src.shape[1:]), fill_value)
    return res.index_add_(0, index, src)


def scatter_max(src, index, dim=0, out=None, dim_size=None, fill_value=None):
    assert dim == 0 and out is None
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() else 0
    idx = index.view(-1, *([1] * (src.dim() - 1))).expand_as(src)
    res = src.new_full((dim_size,) + tuple(src.shape[1:]), float("-inf"))
    res = res.scatter_reduce(0, idx, src, reduce="amax", include_self=True)
    return res, None
