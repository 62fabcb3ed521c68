"""Dispatcher registrations: torch.ops.torch_scatter.* / torch.ops.torch_sparse.* with the
upstream schemas (torch-scatter 2.0.9 csrc/scatter.cpp, csrc/segment_csr.cpp; torch-sparse
0.6.12 csrc/spmm.cpp, csrc/convert.cpp), so PyG's TorchScript paths, which call the ops
directly, resolve to the B200 library.  Registered from Python with torch.library: the CUDA key
runs the kernels, the Autograd key routes through gno_b200.autograd.

If a real torch_scatter / torch_sparse extension already defined these ops, registration is
skipped (the namespaces cannot be defined twice) and `registered` says so.
"""
import torch

from . import autograd as _ag
from . import ops as _ops

registered = {"torch_scatter": False, "torch_sparse": False}
_libs = []

_SCATTER_SCHEMAS = {
    "scatter_sum": "(Tensor src, Tensor index, int dim, Tensor? optional_out, int? dim_size) -> Tensor",
    "scatter_mul": "(Tensor src, Tensor index, int dim, Tensor? optional_out, int? dim_size) -> Tensor",
    "scatter_mean": "(Tensor src, Tensor index, int dim, Tensor? optional_out, int? dim_size) -> Tensor",
    "scatter_min": "(Tensor src, Tensor index, int dim, Tensor? optional_out, int? dim_size) -> (Tensor, Tensor)",
    "scatter_max": "(Tensor src, Tensor index, int dim, Tensor? optional_out, int? dim_size) -> (Tensor, Tensor)",
    "segment_sum_csr": "(Tensor src, Tensor indptr, Tensor? optional_out) -> Tensor",
    "segment_mean_csr": "(Tensor src, Tensor indptr, Tensor? optional_out) -> Tensor",
    "segment_min_csr": "(Tensor src, Tensor indptr, Tensor? optional_out) -> (Tensor, Tensor)",
    "segment_max_csr": "(Tensor src, Tensor indptr, Tensor? optional_out) -> (Tensor, Tensor)",
    "gather_csr": "(Tensor src, Tensor indptr, Tensor? optional_out) -> Tensor",
}
_SPARSE_SCHEMAS = {
    "spmm_sum": "(Tensor? opt_row, Tensor rowptr, Tensor col, Tensor? opt_value, Tensor? opt_colptr, "
                "Tensor? opt_csr2csc, Tensor mat) -> Tensor",
    "spmm_mean": "(Tensor? opt_row, Tensor rowptr, Tensor col, Tensor? opt_value, Tensor? opt_rowcount, "
                 "Tensor? opt_colptr, Tensor? opt_csr2csc, Tensor mat) -> Tensor",
    "spmm_min": "(Tensor rowptr, Tensor col, Tensor? opt_value, Tensor mat) -> (Tensor, Tensor)",
    "spmm_max": "(Tensor rowptr, Tensor col, Tensor? opt_value, Tensor mat) -> (Tensor, Tensor)",
    "ind2ptr": "(Tensor ind, int M) -> Tensor",
    "ptr2ind": "(Tensor ptr, int E) -> Tensor",
}


def _scatter_impl(reduce, differentiable):
    def fn(src, index, dim, optional_out, dim_size):
        if optional_out is not None or not differentiable:
            return _ops.scatter(src, index, dim, optional_out, dim_size, reduce,
                                return_arg=reduce in ("min", "max"))
        return _ag.scatter(src, index, dim, dim_size, reduce)
    return fn


def _segment_impl(reduce, differentiable=False):
    def fn(src, indptr, optional_out):
        if optional_out is not None:
            raise NotImplementedError("gno_b200 segment_csr: out= is not supported")
        if differentiable and indptr.dim() == 1:
            return _ag.segment_csr(src, indptr, reduce)
        return _ops.segment_csr(src, indptr, None, reduce, return_arg=reduce in ("min", "max"))
    return fn


def _gather_csr(src, indptr, optional_out):
    if optional_out is not None:
        raise NotImplementedError("gno_b200 gather_csr: out= is not supported")
    return _ops.gather_csr(src, indptr)


def _spmm_sum(opt_row, rowptr, col, opt_value, opt_colptr, opt_csr2csc, mat):
    return _ops.spmm_csr(rowptr, col, opt_value, mat, "sum")


def _spmm_sum_ag(opt_row, rowptr, col, opt_value, opt_colptr, opt_csr2csc, mat):
    return _ag.spmm_csr(rowptr, col, opt_value, mat, "sum")


def _spmm_mean(opt_row, rowptr, col, opt_value, opt_rowcount, opt_colptr, opt_csr2csc, mat):
    return _ops.spmm_csr(rowptr, col, opt_value, mat, "mean")


def _spmm_mean_ag(opt_row, rowptr, col, opt_value, opt_rowcount, opt_colptr, opt_csr2csc, mat):
    return _ag.spmm_csr(rowptr, col, opt_value, mat, "mean")


def _spmm_minmax(reduce, differentiable=False):
    def fn(rowptr, col, opt_value, mat):
        if differentiable:
            return _ag.spmm_csr(rowptr, col, opt_value, mat, reduce)
        return _ops.spmm_csr(rowptr, col, opt_value, mat, reduce, return_arg=True)
    return fn


def _ind2ptr(ind, M):
    """rowptr of a sorted index vector (torch-sparse convert.cpp ind2ptr)."""
    counts = torch.bincount(ind, minlength=M)[:M]
    ptr = torch.zeros(M + 1, dtype=torch.int64, device=ind.device)
    torch.cumsum(counts, 0, out=ptr[1:])
    return ptr


def _ptr2ind(ptr, E):
    counts = ptr[1:] - ptr[:-1]
    return torch.repeat_interleave(torch.arange(counts.numel(), device=ptr.device), counts, output_size=E)


def _register(ns, schemas, impls):
    try:
        lib = torch.library.Library(ns, "DEF")
        for name, schema in schemas.items():
            lib.define(name + schema)
    except RuntimeError:
        return False  # namespace/ops already defined by a real extension
    for name, (cuda_fn, autograd_fn) in impls.items():
        lib.impl(name, cuda_fn, "CUDA")
        if autograd_fn is not None:
            lib.impl(name, autograd_fn, "AutogradCUDA")
    _libs.append(lib)
    return True


def register():
    if not registered["torch_scatter"]:
        impls = {}
        for red in ("sum", "mul", "mean", "min", "max"):
            impls["scatter_" + red] = (_scatter_impl(red, False), _scatter_impl(red, True))
        for red in ("sum", "mean", "min", "max"):
            impls[f"segment_{red}_csr"] = (_segment_impl(red), _segment_impl(red, True))
        impls["gather_csr"] = (_gather_csr, None)
        registered["torch_scatter"] = _register("torch_scatter", _SCATTER_SCHEMAS, impls)
    if not registered["torch_sparse"]:
        impls = {"spmm_sum": (_spmm_sum, _spmm_sum_ag), "spmm_mean": (_spmm_mean, _spmm_mean_ag),
                 "spmm_min": (_spmm_minmax("min"), _spmm_minmax("min", True)),
                 "spmm_max": (_spmm_minmax("max"), _spmm_minmax("max", True)),
                 "ind2ptr": (_ind2ptr, None), "ptr2ind": (_ptr2ind, None)}
        registered["torch_sparse"] = _register("torch_sparse", _SPARSE_SCHEMAS, impls)
    return registered
