"""torch_scatter.composite (2.0.9): softmax / log_softmax / logsumexp / std over groups, composed
from the B200 scatter kernels plus element-wise torch ops on the GPU.  Float tensors only."""
from typing import Optional

import torch

from gno_b200 import autograd as _ag
from gno_b200 import ops as _ops


def _needs_grad(t):
    return t.requires_grad and torch.is_grad_enabled()


def _sc(src, index, dim, dim_size, reduce):
    """One scatter of the composite: through the autograd Function when src needs a gradient
    (upstream's composites are differentiable), straight to the kernel otherwise."""
    if _needs_grad(src):
        return _ag.scatter(src, index, dim, dim_size, reduce)
    return _ops.scatter(src, index, dim, None, dim_size, reduce)


def _group_max(src, index, dim, dim_size):
    """Per-group maximum used as the shift of softmax / logsumexp.  The shift cancels in the
    result, so it is a constant for autograd (no gradient flows through it)."""
    return _ops.scatter(src.detach(), index, dim, None, dim_size, "max")


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
    mx = _group_max(src, index, dim, dim_size)
    rec = (src - _expand(mx, index, src, dim)).exp()
    s = _sc(rec, index, dim, dim_size, "sum")
    return (s + eps).log() + mx


def scatter_softmax(src: torch.Tensor, index: torch.Tensor, dim: int = -1,
                    dim_size: Optional[int] = None) -> torch.Tensor:
    if not torch.is_floating_point(src):
        raise ValueError("`scatter_softmax` can only be computed over tensors with floating point data types.")
    if dim < 0:
        dim += src.dim()
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() else 0
    mx = _group_max(src, index, dim, dim_size)
    rec = (src - _expand(mx, index, src, dim)).exp()
    s = _sc(rec, index, dim, dim_size, "sum")
    if _needs_grad(src):
        return rec / _expand(s, index, src, dim)
    return rec.div_(_expand(s, index, src, dim))


def scatter_log_softmax(src: torch.Tensor, index: torch.Tensor, dim: int = -1, eps: float = 1e-12,
                        dim_size: Optional[int] = None) -> torch.Tensor:
    if not torch.is_floating_point(src):
        raise ValueError("`scatter_log_softmax` can only be computed over tensors with floating point data types.")
    if dim < 0:
        dim += src.dim()
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() else 0
    mx = _group_max(src, index, dim, dim_size)
    rec = src - _expand(mx, index, src, dim)
    s = _sc(rec.exp(), index, dim, dim_size, "sum")
    return rec - _expand((s + eps).log(), index, src, dim)


def scatter_std(src: torch.Tensor, index: torch.Tensor, dim: int = -1,
                out: Optional[torch.Tensor] = None, dim_size: Optional[int] = None,
                unbiased: bool = True) -> torch.Tensor:
    if out is not None:
        raise NotImplementedError("gno_b200 scatter_std: out= is not supported")
    if dim < 0:
        dim += src.dim()
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() else 0
    ones = torch.ones_like(src, requires_grad=False)
    count = _ops.scatter(ones, index, dim, None, dim_size, "sum")
    mean = _sc(src, index, dim, dim_size, "sum") / count.clamp(min=1)
    var = src - _expand(mean, index, src, dim)
    var = _sc(var * var, index, dim, dim_size, "sum")
    if unbiased:
        count = count.sub(1).clamp_(min=1)
    return (var / (count + 1e-6)).sqrt()
