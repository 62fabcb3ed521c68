"""Drop-in `torch_scatter` backed by the B200 library (gno_b200).

Keeps the torch-scatter 2.0.9 call signatures the reference imports
(op_bm_scripts/benchmark_scatter_add.py:5-7, benchmark_scatter_max.py:5-7,
benchmark_scatter_min.py:5-7, benchmark_scatter_mean.py:5-7) and the ones
PyG's MessagePassing.aggregate / global_mean_pool reach
(graph_benchmark/models/ptg_models.py:14-20).  Put this directory's parent
(`gnn-ops-benchmark_b200/`) on PYTHONPATH and the reference scripts run
unchanged.  CUDA tensors only: there is no CPU fallback.
"""
from typing import Optional, Tuple

import torch

from gno_b200 import autograd as _ag
from gno_b200 import ops as _ops
from gno_b200 import torch_ops as _torch_ops

__version__ = "2.0.9+gno.b200"

_torch_ops.register()  # torch.ops.torch_scatter.* for TorchScript callers


def _scatter(src, index, dim, out, dim_size, reduce, return_arg=False):
    if out is None and src.requires_grad and torch.is_grad_enabled():
        r = _ag.scatter(src, index, dim, dim_size, reduce)
        if reduce in ("min", "max") and not return_arg:
            return r[0]
        return r
    return _ops.scatter(src, index, dim, out, dim_size, reduce, return_arg=return_arg)


def scatter_sum(src: torch.Tensor, index: torch.Tensor, dim: int = -1,
                out: Optional[torch.Tensor] = None,
                dim_size: Optional[int] = None) -> torch.Tensor:
    return _scatter(src, index, dim, out, dim_size, "sum")


def scatter_add(src: torch.Tensor, index: torch.Tensor, dim: int = -1,
                out: Optional[torch.Tensor] = None,
                dim_size: Optional[int] = None) -> torch.Tensor:
    return _scatter(src, index, dim, out, dim_size, "sum")


def scatter_mul(src: torch.Tensor, index: torch.Tensor, dim: int = -1,
                out: Optional[torch.Tensor] = None,
                dim_size: Optional[int] = None) -> torch.Tensor:
    return _scatter(src, index, dim, out, dim_size, "mul")


def scatter_mean(src: torch.Tensor, index: torch.Tensor, dim: int = -1,
                 out: Optional[torch.Tensor] = None,
                 dim_size: Optional[int] = None) -> torch.Tensor:
    return _scatter(src, index, dim, out, dim_size, "mean")


def scatter_min(src: torch.Tensor, index: torch.Tensor, dim: int = -1,
                out: Optional[torch.Tensor] = None,
                dim_size: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    return _scatter(src, index, dim, out, dim_size, "min", return_arg=True)


def scatter_max(src: torch.Tensor, index: torch.Tensor, dim: int = -1,
                out: Optional[torch.Tensor] = None,
                dim_size: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    return _scatter(src, index, dim, out, dim_size, "max", return_arg=True)


def scatter(src: torch.Tensor, index: torch.Tensor, dim: int = -1,
            out: Optional[torch.Tensor] = None, dim_size: Optional[int] = None,
            reduce: str = "sum") -> torch.Tensor:
    if reduce in ("sum", "add"):
        return scatter_sum(src, index, dim, out, dim_size)
    if reduce == "mul":
        return scatter_mul(src, index, dim, out, dim_size)
    if reduce == "mean":
        return scatter_mean(src, index, dim, out, dim_size)
    if reduce == "min":
        return _scatter(src, index, dim, out, dim_size, "min")
    if reduce == "max":
        return _scatter(src, index, dim, out, dim_size, "max")
    raise ValueError


def segment_csr(src: torch.Tensor, indptr: torch.Tensor, out: Optional[torch.Tensor] = None,
                reduce: str = "sum") -> torch.Tensor:
    if reduce not in ("sum", "add", "mean", "min", "max"):
        raise ValueError
    if out is None and src.requires_grad and torch.is_grad_enabled():
        r = _ag.segment_csr(src, indptr, "sum" if reduce == "add" else reduce)
        return r[0] if reduce in ("min", "max") else r
    return _ops.segment_csr(src, indptr, out, reduce)


def segment_sum_csr(src, indptr, out=None):
    return segment_csr(src, indptr, out, "sum")


def segment_add_csr(src, indptr, out=None):
    return segment_csr(src, indptr, out, "sum")


def segment_mean_csr(src, indptr, out=None):
    return segment_csr(src, indptr, out, "mean")


def segment_min_csr(src, indptr, out=None):
    if src.requires_grad and torch.is_grad_enabled():
        return _ag.segment_csr(src, indptr, "min")
    return _ops.segment_csr(src, indptr, None, "min", return_arg=True)


def segment_max_csr(src, indptr, out=None):
    if src.requires_grad and torch.is_grad_enabled():
        return _ag.segment_csr(src, indptr, "max")
    return _ops.segment_csr(src, indptr, None, "max", return_arg=True)


def gather_csr(src: torch.Tensor, indptr: torch.Tensor,
               out: Optional[torch.Tensor] = None) -> torch.Tensor:
    if out is not None:
        raise NotImplementedError("gno_b200 gather_csr: out= is not supported")
    return _ops.gather_csr(src, indptr)


def segment_coo(src: torch.Tensor, index: torch.Tensor, out: Optional[torch.Tensor] = None,
                dim_size: Optional[int] = None, reduce: str = "sum") -> torch.Tensor:
    if reduce not in ("sum", "add", "mean", "min", "max"):
        raise ValueError
    return _ops.segment_coo(src, index, out, dim_size, reduce)


def segment_sum_coo(src, index, out=None, dim_size=None):
    return segment_coo(src, index, out, dim_size, "sum")


def segment_add_coo(src, index, out=None, dim_size=None):
    return segment_coo(src, index, out, dim_size, "sum")


def segment_mean_coo(src, index, out=None, dim_size=None):
    return segment_coo(src, index, out, dim_size, "mean")


def segment_min_coo(src, index, out=None, dim_size=None):
    return _ops.segment_coo(src, index, out, dim_size, "min", return_arg=True)


def segment_max_coo(src, index, out=None, dim_size=None):
    return _ops.segment_coo(src, index, out, dim_size, "max", return_arg=True)


def gather_coo(src: torch.Tensor, index: torch.Tensor,
               out: Optional[torch.Tensor] = None) -> torch.Tensor:
    if out is not None:
        raise NotImplementedError("gno_b200 gather_coo: out= is not supported")
    return _ops.gather_coo(src, index)


from .composite import (scatter_log_softmax, scatter_logsumexp, scatter_softmax,  # noqa: E402
                        scatter_std)

__all__ = [
    "scatter_std", "scatter_logsumexp", "scatter_softmax", "scatter_log_softmax",
    "scatter_sum", "scatter_add", "scatter_mul", "scatter_mean", "scatter_min", "scatter_max",
    "scatter", "segment_csr", "segment_sum_csr", "segment_add_csr", "segment_mean_csr",
    "segment_min_csr", "segment_max_csr", "gather_csr", "segment_coo", "segment_sum_coo",
    "segment_add_coo", "segment_mean_coo", "segment_min_coo", "segment_max_coo", "gather_coo",
]
