"""torch_scatter.composite (2.0.9): softmax / log_softmax / logsumexp / std over groups, composed
from the B200 scatter kernels plus element-wise torch ops on the GPU.  Float tensors only."""
from typing import Optional

import torch

from gno_b200 import ops as _ops


def _expand(t, index, src, dim):
    """Values of the reduced tensor `t` gathered back to src's shape along dim."""
    if index.dim() == 1 and src.dim() > 1:
        return t.index_select(dim, index)
    return t.gather(dim, index)


def scatter_logsumexp(src: torch.Tensor, index: torch.Tensor, dim: int = -1,
                      out: Optional[torch.Tensor] = None, dim_size: Optional[int] = None,
                      eps: float = 1e-12) -> torch.Tensor:
    if not torch.is_floating_point(src):
        raise ValueError("`scatter_logsumexp` can only be computed over tensors with floating point data types.")
    if out is not None:
        raise NotImplementedError("gno_b200 scatter_logsumexp: out= is not supported")
    if dim < 0:
        dim += src.dim()
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() else 0
    mx = _ops.scatter(src, index, dim, None, dim_size, "max")
    rec = (src - _expand(mx, index, src, dim)).exp_()
    s = _ops.scatter(rec, index, dim, None, dim_size, "sum")
    return s.add_(eps).log_().add_(mx)


def scatter_softmax(src: torch.Tensor, index: torch.Tensor, dim: int = -1,
                    dim_size: Optional[int] = None) -> torch.Tensor:
    if not torch.is_floating_point(src):
        raise ValueError("`scatter_softmax` can only be computed over tensors with floating point data types.")
    if dim < 0:
        dim += src.dim()
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() else 0
    mx = _ops.scatter(src, index, dim, None, dim_size, "max")
    rec = (src - _expand(mx, index, src, dim)).exp_()
    s = _ops.scatter(rec, index, dim, None, dim_size, "sum")
    return rec.div_(_expand(s, index, src, dim))


def scatter_log_softmax(src: torch.Tensor, index: torch.Tensor, dim: int = -1, eps: float = 1e-12,
                        dim_size: Optional[int] = None) -> torch.Tensor:
    if not torch.is_floating_point(src):
        raise ValueError("`scatter_log_softmax` can only be computed over tensors with floating point data types.")
    if dim < 0:
        dim += src.dim()
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() else 0
    mx = _ops.scatter(src, index, dim, None, dim_size, "max")
    rec = src - _expand(mx, index, src, dim)
    s = _ops.scatter(rec.exp(), index, dim, None, dim_size, "sum")
    return rec.sub_(_expand(s.add_(eps).log_(), index, src, dim))


def scatter_std(src: torch.Tensor, index: torch.Tensor, dim: int = -1,
                out: Optional[torch.Tensor] = None, dim_size: Optional[int] = None,
                unbiased: bool = True) -> torch.Tensor:
    if out is not None:
        raise NotImplementedError("gno_b200 scatter_std: out= is not supported")
    if dim < 0:
        dim += src.dim()
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() else 0
    ones = torch.ones_like(src)
    count = _ops.scatter(ones, index, dim, None, dim_size, "sum")
    mean = _ops.scatter(src, index, dim, None, dim_size, "sum") / count.clamp(min=1)
    var = src - _expand(mean, index, src, dim)
    var = _ops.scatter(var * var, index, dim, None, dim_size, "sum")
    if unbiased:
        count = count.sub(1).clamp_(min=1)
    return var.div_(count.add(1e-6)).sqrt_()
