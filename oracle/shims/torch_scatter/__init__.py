"""Shim for torch_scatter (absent from this image). TEST INFRASTRUCTURE ONLY.

Published semantics of ``scatter_add(src, index, dim=-1, out=None, dim_size=None)``:
out[index[i]] += src[i] along ``dim``; output has ``dim_size`` entries along ``dim``.
Reference call sites: model.py:75, data_loader.py:126.
"""
import torch


def scatter_add(src, index, dim=-1, out=None, dim_size=None, fill_value=0):
    if dim < 0:
        dim = src.dim() + dim
    if out is None:
        size = list(src.size())
        size[dim] = int(index.max()) + 1 if dim_size is None else int(dim_size)
        out = src.new_full(size, fill_value)
    return out.scatter_add_(dim, index, src)
