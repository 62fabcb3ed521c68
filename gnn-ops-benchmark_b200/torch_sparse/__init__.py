"""Drop-in `torch_sparse` functions backed by the B200 library (gno_b200).

Keeps the torch-sparse 0.6.12 functional signatures: `coalesce` as imported by
the reference (op_bm_scripts/benchmark_sparse_coalesce.py:7,35-37), plus
`transpose` and `spmm`, whose kernels the reference's sparse_transpose /
sparse_spmm data rows time.  CUDA tensors only: there is no CPU fallback.
"""
from typing import Optional, Tuple

import torch

from gno_b200 import ops as _ops
from gno_b200 import torch_ops as _torch_ops

_torch_ops.register()  # torch.ops.torch_sparse.* for TorchScript callers

__version__ = "0.6.12+gno.b200"


def coalesce(index: torch.Tensor, value: Optional[torch.Tensor], m: int, n: int,
             op: str = "add") -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    return _ops.coalesce(index, value, m, n, op)


def transpose(index: torch.Tensor, value: Optional[torch.Tensor], m: int, n: int,
              coalesced: bool = True) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    return _ops.transpose(index, value, m, n, coalesced)


def spmm(index: torch.Tensor, value: torch.Tensor, m: int, n: int,
         matrix: torch.Tensor) -> torch.Tensor:
    if torch.is_grad_enabled() and (matrix.requires_grad or (value is not None and value.requires_grad)):
        # differentiable form: the CSR of the row-sorted entries through the autograd Function
        # (the value permutation is a differentiable indexing op)
        from gno_b200 import autograd as _ag
        from gno_b200.plan import plan_cache
        plan = plan_cache.get(index[0].contiguous(), m)
        perm = plan.perm.to(torch.int64)
        col_s = index[1][perm]
        val_s = value[perm] if value is not None else None
        squeeze = matrix.dim() == 1
        out = _ag.spmm_csr(plan.rowptr, col_s, val_s, matrix.unsqueeze(-1) if squeeze else matrix, "sum")
        return out.squeeze(-1) if squeeze else out
    return _ops.spmm(index, value, m, n, matrix)


from .tensor import SparseTensor, matmul  # noqa: E402

__all__ = ["coalesce", "transpose", "spmm", "SparseTensor", "matmul"]
