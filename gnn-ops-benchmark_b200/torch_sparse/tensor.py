"""A minimal torch_sparse.SparseTensor over the B200 kernels: the surface
PyG's `message_and_aggregate` path reaches — construction from (row, col, value) or edge_index,
cached CSR (`rowptr`, `col`, value in row order), cached transpose (`t()`), `matmul(dense,
reduce)`, row sums.  Caches play the role of upstream's `rowptr` / `csr2csc` storage caches; the
aggregation plan of the CSR is cached by the operator layer on the rowptr tensor.
"""
from typing import Optional, Tuple

import torch

from gno_b200 import autograd as _ag
from gno_b200 import ops as _ops


class SparseTensor:
    def __init__(self, row: Optional[torch.Tensor] = None, rowptr: Optional[torch.Tensor] = None,
                 col: Optional[torch.Tensor] = None, value: Optional[torch.Tensor] = None,
                 sparse_sizes: Optional[Tuple[int, int]] = None, is_sorted: bool = False):
        if col is None or (row is None and rowptr is None):
            raise ValueError("SparseTensor needs col and one of row / rowptr")
        if sparse_sizes is None:
            if row is None:
                m = rowptr.numel() - 1
            else:
                m = int(row.max()) + 1 if row.numel() else 0
            n = int(col.max()) + 1 if col.numel() else 0
            sparse_sizes = (m, n)
        self._sizes = (int(sparse_sizes[0]), int(sparse_sizes[1]))
        if row is not None and not is_sorted and row.numel() > 1:
            # upstream SparseStorage sorts by row*n+col and keeps duplicate entries (a multigraph's
            # parallel edges count separately in matmul / nnz): sort only, never merge
            if value is not None and value.requires_grad and torch.is_grad_enabled():
                # keep the autograd graph: sort the keys with an explicit permutation and reorder the
                # values by (differentiable) indexing
                m_, n_ = self._sizes
                key = row * n_ + col
                iota = torch.arange(key.numel(), dtype=torch.int32, device=key.device)
                skey, perm = _ops.sort_pairs(key, iota, 0, max(1, (max(m_ * n_, 1) - 1).bit_length()))
                row, col = torch.div(skey, n_, rounding_mode="floor"), skey % n_
                value = value[perm.to(torch.int64)]
            else:
                index, value = _ops.sort_coo(torch.stack([row, col]), value, self._sizes[0], self._sizes[1])
                row, col = index[0], index[1]
            rowptr = None
        self._row, self._col, self._value, self._rowptr = row, col, value, rowptr
        self._t = None

    # ---- construction ---------------------------------------------------------------------------
    @classmethod
    def from_edge_index(cls, edge_index: torch.Tensor, edge_attr: Optional[torch.Tensor] = None,
                        sparse_sizes: Optional[Tuple[int, int]] = None, is_sorted: bool = False):
        return cls(row=edge_index[0], col=edge_index[1], value=edge_attr, sparse_sizes=sparse_sizes,
                   is_sorted=is_sorted)

    # ---- accessors ------------------------------------------------------------------------------
    def sparse_sizes(self):
        return self._sizes

    def sparse_size(self, dim):
        return self._sizes[dim]

    def size(self, dim=None):
        return self._sizes if dim is None else self._sizes[dim]

    def nnz(self):
        return self._col.numel()

    @property
    def device(self):
        return self._col.device

    def _get_rowptr(self):
        if self._rowptr is None:
            self._rowptr = torch.ops.torch_sparse.ind2ptr(self._row, self._sizes[0])
        return self._rowptr

    def _get_row(self):
        if self._row is None:
            self._row = torch.ops.torch_sparse.ptr2ind(self._rowptr, self._col.numel())
        return self._row

    def coo(self):
        return self._get_row(), self._col, self._value

    def csr(self):
        return self._get_rowptr(), self._col, self._value

    def set_value(self, value, layout=None):
        out = SparseTensor(row=self._row, rowptr=self._rowptr, col=self._col, value=value,
                           sparse_sizes=self._sizes, is_sorted=True)
        return out

    # ---- ops ------------------------------------------------------------------------------------
    def t(self):
        if self._t is None:
            row, col, value = self.coo()
            index, tv = _ops.transpose(torch.stack([row, col]), value, self._sizes[0], self._sizes[1])
            self._t = SparseTensor(row=index[0], col=index[1], value=tv,
                                   sparse_sizes=(self._sizes[1], self._sizes[0]), is_sorted=True)
            self._t._t = self
        return self._t

    def matmul(self, other: torch.Tensor, reduce: str = "sum") -> torch.Tensor:
        return matmul(self, other, reduce)

    def __matmul__(self, other):
        return matmul(self, other, "sum")

    def sum(self, dim: Optional[int] = None):
        rowptr, col, value = self.csr()
        if value is None:
            value = torch.ones(col.numel(), device=col.device)
        if dim is None:
            return value.sum()
        if dim in (1, -1):
            return _ops.segment_csr(value.view(-1, 1), rowptr, None, "sum").view(-1)
        if dim == 0:
            return self.t().sum(dim=1)
        raise ValueError("dim must be None, 0 or 1")

    def to_dense(self):
        row, col, value = self.coo()
        v = value if value is not None else torch.ones(col.numel(), device=col.device)
        out = torch.zeros(self._sizes, dtype=v.dtype, device=col.device)
        out[row, col] = v
        return out


def matmul(src: SparseTensor, other: torch.Tensor, reduce: str = "sum") -> torch.Tensor:
    """torch_sparse.matmul(SparseTensor, dense, reduce): out[i] = reduce_k value[k] * other[col[k]]."""
    if not isinstance(other, torch.Tensor):
        raise NotImplementedError("gno_b200: SparseTensor @ SparseTensor (spspmm) is out of scope")
    rowptr, col, value = src.csr()
    squeeze = other.dim() == 1
    mat = other.unsqueeze(-1) if squeeze else other
    reduce = "sum" if reduce == "add" else reduce
    needs_grad = torch.is_grad_enabled() and (mat.requires_grad or (value is not None and value.requires_grad))
    if needs_grad:  # backward = the transposed aggregation (plan on col, cached like upstream's csr2csc)
        out = _ag.spmm_csr(rowptr, col, value, mat, reduce)
        out = out[0] if reduce in ("min", "max") else out
    else:
        out = _ops.spmm_csr(rowptr, col, value, mat, reduce)
    return out.squeeze(-1) if squeeze else out
