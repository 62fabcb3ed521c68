"""Differentiable wrappers: OpProfiler profiles a training step
(graph_benchmark/profile/OpProfiler.py:277-292), so the aggregation needs a backward.

Backward formulas are torch-scatter 2.0.9's (csrc/scatter.cpp ScatterSum/ScatterMean/ScatterMin/
ScatterMax/ScatterMul::backward):
    sum   grad_src = grad_out gathered by index
    mean  grad_src = (grad_out / clamp(count, 1)) gathered by index
    min/max  grad_src = scatter of grad_out to the arg positions (no collisions)
    mul   grad_src = (grad_out * out) gathered by index / src
For the fused gather→scatter the backward w.r.t. x is the transposed aggregation — the same
segment-reduce kernel on a plan sorted by the SOURCE ids (cached like the forward plan).
"""
import torch

from . import ops


def _gather_dim(t, dim, index, like):
    """t gathered along dim by a 1-D or full-shape index, shaped like `like`."""
    if index.dim() == 1 and t.dim() == 2 and dim == 0:
        return ops.index_select(t, 0, index)
    if index.dim() == 1:
        return t.index_select(dim, index)
    if index.shape != like.shape:
        index = index.expand_as(like)
    return t.gather(dim, index)


class _Scatter(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src, index, dim, dim_size, reduce):
        if dim < 0:
            dim += src.dim()
        ctx.dim, ctx.reduce, ctx.shape = dim, reduce, src.shape
        if reduce in ("min", "max"):
            out, arg = ops.scatter(src, index, dim, None, dim_size, reduce, return_arg=True)
            ctx.save_for_backward(arg)
            ctx.mark_non_differentiable(arg)
            return out, arg
        out = ops.scatter(src, index, dim, None, dim_size, reduce)
        if reduce == "mul":
            ctx.save_for_backward(index, src, out)
        else:
            ctx.save_for_backward(index)
        return out

    @staticmethod
    def backward(ctx, grad_out, *unused):
        dim, reduce = ctx.dim, ctx.reduce
        grad_out = grad_out.contiguous()
        if reduce in ("min", "max"):
            (arg,) = ctx.saved_tensors
            shape = list(ctx.shape)
            shape[dim] += 1  # sentinel row for empty outputs, trimmed below
            g = grad_out.new_zeros(shape).scatter_(dim, arg, grad_out)
            return g.narrow(dim, 0, shape[dim] - 1), None, None, None, None
        if reduce == "mul":
            index, src, out = ctx.saved_tensors
            g = _gather_dim(grad_out * out, dim, index, src) / src
            return g.masked_fill_(g.isnan(), 0), None, None, None, None
        (index,) = ctx.saved_tensors
        like = grad_out.new_empty(ctx.shape)
        if reduce == "mean":
            idx1 = index if index.dim() == 1 else None
            if idx1 is not None:
                cnt = torch.bincount(idx1, minlength=grad_out.size(dim)).clamp_(min=1).to(grad_out.dtype)
                shape = [1] * grad_out.dim()
                shape[dim] = -1
                grad_out = grad_out / cnt.view(shape)
            else:
                ones = torch.ones_like(like)
                cnt = ops.scatter(ones, index, dim, None, grad_out.size(dim), "sum").clamp_(min=1)
                grad_out = grad_out / cnt
        return _gather_dim(grad_out, dim, index, like), None, None, None, None


class _GatherScatter(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, src_ids, dst_ids, dim_size, reduce):
        ctx.reduce, ctx.n_src, ctx.dim_size = reduce, x.size(0), dim_size
        if reduce in ("min", "max"):
            out, arg = ops.gather_scatter(x, src_ids, dst_ids, dim_size, reduce, return_arg=True)
            ctx.save_for_backward(src_ids, arg)
            ctx.mark_non_differentiable(arg)
            return out, arg
        out = ops.gather_scatter(x, src_ids, dst_ids, dim_size, reduce)
        ctx.save_for_backward(src_ids, dst_ids)
        return out

    @staticmethod
    def backward(ctx, grad_out, *unused):
        reduce = ctx.reduce
        grad_out = grad_out.contiguous()
        if reduce in ("min", "max"):
            src_ids, arg = ctx.saved_tensors
            E = src_ids.numel()
            # row of x that won each output element (sentinel -> dropped by the range check)
            winner = torch.cat([src_ids, src_ids.new_full((1,), -1)])[arg.clamp(max=E)]
            g = ops.scatter(grad_out, winner, 0, None, ctx.n_src, "sum")  # full-shape index, collisions
            return g, None, None, None, None
        src_ids, dst_ids = ctx.saved_tensors
        if reduce == "mean":
            cnt = torch.bincount(dst_ids, minlength=ctx.dim_size).clamp_(min=1).to(grad_out.dtype)
            grad_out = grad_out / cnt.view(-1, 1)
        elif reduce != "sum":
            raise NotImplementedError("gather_scatter backward: sum/mean/min/max")
        # transposed aggregation: grad_x[s] = sum over edges with source s of grad_out[dst]
        return ops.gather_scatter(grad_out, dst_ids, src_ids, ctx.n_src, "sum"), None, None, None, None


class _SegmentCSR(torch.autograd.Function):
    """torch_scatter.segment_csr (1-D indptr) with upstream's backward (csrc/segment_csr.cpp):
    sum -> gather_csr(grad), mean -> gather_csr(grad / count), min/max -> grad scattered to arg."""

    @staticmethod
    def forward(ctx, src, indptr, reduce):
        ctx.reduce, ctx.shape = reduce, src.shape
        if reduce in ("min", "max"):
            out, arg = ops.segment_csr(src, indptr, None, reduce, return_arg=True)
            ctx.save_for_backward(arg)
            ctx.mark_non_differentiable(arg)
            return out, arg
        ctx.save_for_backward(indptr)
        return ops.segment_csr(src, indptr, None, reduce)

    @staticmethod
    def backward(ctx, grad_out, *unused):
        grad_out = grad_out.contiguous()
        E = ctx.shape[0]
        if ctx.reduce in ("min", "max"):
            (arg,) = ctx.saved_tensors
            g = grad_out.new_zeros([E + 1] + list(ctx.shape[1:])).scatter_(0, arg, grad_out)
            return g[:E], None, None
        (indptr,) = ctx.saved_tensors
        if ctx.reduce == "mean":
            cnt = (indptr[1:] - indptr[:-1]).clamp(min=1).to(grad_out.dtype)
            grad_out = grad_out / cnt.view([-1] + [1] * (grad_out.dim() - 1))
        g = ops.gather_csr(grad_out, indptr)            # rows [0, indptr[-1]); zeros before indptr[0]
        if g.size(0) < E:                                # elements past the last pointer get no gradient
            g = torch.cat([g, g.new_zeros([E - g.size(0)] + list(g.shape[1:]))])
        return g, None, None


class _SpmmCSR(torch.autograd.Function):
    """torch_sparse.matmul(SparseTensor, dense, reduce) on a CSR: backward w.r.t. the dense matrix
    is the TRANSPOSED aggregation (same segment-reduce kernel, plan on the column ids, cached);
    w.r.t. the edge values it is one dot product per edge."""

    @staticmethod
    def forward(ctx, rowptr, col, value, mat, reduce):
        ctx.reduce, ctx.n_src = reduce, mat.size(0)
        ctx.has_value = value is not None
        if reduce in ("min", "max"):
            out, arg = ops.spmm_csr(rowptr, col, value, mat, reduce, return_arg=True)
            ctx.save_for_backward(col, arg, value if value is not None else col.new_empty(0), mat)
            ctx.mark_non_differentiable(arg)
            return out, arg
        ctx.save_for_backward(rowptr, col, value if value is not None else col.new_empty(0), mat)
        return ops.spmm_csr(rowptr, col, value, mat, reduce)

    @staticmethod
    def backward(ctx, grad_out, *unused):
        grad_out = grad_out.contiguous()
        need_mat, need_val = ctx.needs_input_grad[3], ctx.needs_input_grad[2] and ctx.has_value
        if ctx.reduce in ("min", "max"):
            col, arg, value, mat = ctx.saved_tensors
            nnz = col.numel()
            safe = arg.clamp(max=max(nnz - 1, 0))
            hit = arg < nnz
            g_mat = g_val = None
            if need_mat:
                w = grad_out if not ctx.has_value else grad_out * value[safe].to(grad_out.dtype)
                winner = torch.where(hit, col[safe], torch.full_like(arg, -1))
                g_mat = ops.scatter(w.contiguous(), winner, 0, None, ctx.n_src, "sum")
            if need_val:
                contrib = torch.where(hit, grad_out * mat[col[safe], torch.arange(mat.size(1), device=mat.device)],
                                      torch.zeros_like(grad_out))
                g_val = torch.zeros(nnz + 1, dtype=grad_out.dtype, device=grad_out.device)
                g_val.index_add_(0, arg.reshape(-1).clamp(max=nnz), contrib.reshape(-1))
                g_val = g_val[:nnz].to(value.dtype)
            return None, None, g_val, g_mat, None
        rowptr, col, value, mat = ctx.saved_tensors
        if ctx.reduce == "mean":
            cnt = (rowptr[1:] - rowptr[:-1]).clamp(min=1).to(grad_out.dtype)
            grad_out = grad_out / cnt.view(-1, 1)
        elif ctx.reduce not in ("sum", "add"):
            raise NotImplementedError("spmm backward: sum / mean / min / max")
        g_mat = ops.spmm_csr_t(rowptr, col, value if ctx.has_value else None, grad_out, ctx.n_src) if need_mat else None
        g_val = None
        if need_val:
            row = ops.csr_rows(rowptr, col.numel())
            g_val = (grad_out.index_select(0, row) * mat.index_select(0, col)).sum(-1).to(value.dtype)
        return None, None, g_val, g_mat, None


def segment_csr(src, indptr, reduce="sum"):
    """Differentiable torch_scatter.segment_csr (1-D indptr). min/max return (out, arg)."""
    if indptr.dim() != 1:
        raise NotImplementedError("gno_b200: segment_csr gradients cover a 1-D indptr")
    return _SegmentCSR.apply(src, indptr, reduce)


def spmm_csr(rowptr, col, value, mat, reduce="sum"):
    """Differentiable CSR spmm (w.r.t. mat and value). min/max return (out, arg)."""
    return _SpmmCSR.apply(rowptr, col, value, mat, reduce)


def scatter(src, index, dim=-1, dim_size=None, reduce="sum"):
    """Differentiable torch_scatter.scatter (no out=). min/max return (out, arg)."""
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() else 0
    return _Scatter.apply(src, index, dim, dim_size, reduce)


def gather_scatter(x, src_ids, dst_ids, dim_size, reduce="sum"):
    """Differentiable fused message passing. min/max return (out, arg)."""
    return _GatherScatter.apply(x, src_ids, dst_ids, dim_size, reduce)
