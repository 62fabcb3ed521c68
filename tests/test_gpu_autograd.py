"""Backward of the aggregation and the dispatcher registrations
(torch.ops.torch_scatter.* / torch.ops.torch_sparse.*), checked against native torch autograd
of the equivalent formulation (index_select + scatter_reduce / index_add_)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _native_scatter(src, index, dim, n, reduce):
    idx = index
    if index.dim() == 1 and src.dim() > 1:
        shape = [1] * src.dim()
        shape[dim] = -1
        idx = index.view(shape).expand_as(src)
    size = list(src.shape)
    size[dim] = n
    red = {"sum": "sum", "mean": "mean", "max": "amax", "min": "amin", "mul": "prod"}[reduce]
    base = torch.ones(size, dtype=src.dtype, device=src.device) if reduce == "mul" else \
        torch.zeros(size, dtype=src.dtype, device=src.device)
    return base.scatter_reduce(dim, idx, src, red, include_self=False)


@pytest.mark.parametrize("reduce", ["sum", "mean", "max", "min", "mul"])
@pytest.mark.parametrize("full_index", [False, True])
def test_scatter_backward(cuda, reduce, full_index):
    import torch_scatter
    g = torch.Generator().manual_seed(3)
    E, F, N = 400, 12, 37
    src = torch.rand(E, F, generator=g) + 0.5  # distinct values: max/min winners are unique
    index = torch.randint(0, N, (E, F) if full_index else (E,), generator=g)
    w = torch.randn(N, F, generator=g)
    a = src.clone().to(cuda).requires_grad_()
    out = torch_scatter.scatter(a, index.to(cuda), dim=0, dim_size=N, reduce=reduce)
    (out * w.to(cuda)).sum().backward()
    b = src.clone().double().requires_grad_()
    ref = _native_scatter(b, index, 0, N, reduce)
    (ref * w.double()).sum().backward()
    assert torch.allclose(out.detach().cpu().double(), ref.detach(), rtol=1e-4, atol=1e-5)
    assert torch.allclose(a.grad.cpu().double(), b.grad, rtol=1e-4, atol=1e-5), reduce


@pytest.mark.parametrize("reduce", ["sum", "mean", "max", "min"])
def test_gather_scatter_backward(cuda, reduce):
    from gno_b200 import autograd as ag
    g = torch.Generator().manual_seed(4)
    n, E, F, N = 90, 2000, 20, 50
    x = torch.rand(n, F, generator=g) + torch.arange(n).view(-1, 1) * 1e-3
    s, d = torch.randint(0, n, (E,), generator=g), torch.randint(0, N, (E,), generator=g)
    w = torch.randn(N, F, generator=g)
    a = x.clone().to(cuda).requires_grad_()
    out = ag.gather_scatter(a, s.to(cuda), d.to(cuda), N, reduce)
    out = out[0] if isinstance(out, tuple) else out
    (out * w.to(cuda)).sum().backward()
    b = x.clone().double().requires_grad_()
    ref = _native_scatter(b.index_select(0, s), d, 0, N, reduce)
    (ref * w.double()).sum().backward()
    assert torch.allclose(out.detach().cpu().double(), ref.detach(), rtol=1e-4, atol=1e-5)
    if reduce in ("sum", "mean"):
        assert torch.allclose(a.grad.cpu().double(), b.grad, rtol=1e-4, atol=1e-4)
    else:
        # ties (the same source row reaching a destination twice) split the gradient differently in
        # torch's amax (evenly) and upstream torch_scatter (first winner): compare the row totals
        assert torch.allclose(a.grad.cpu().double().sum(0), b.grad.sum(0), rtol=1e-4, atol=1e-4)


def test_dispatcher_ops_and_torchscript(cuda):
    import torch_scatter  # noqa: F401  (registers the ops)
    import torch_sparse  # noqa: F401
    import oracle
    g = torch.Generator().manual_seed(6)
    src = torch.randn(300, 8, generator=g)
    idx = torch.randint(0, 20, (300,), generator=g)
    out = torch.ops.torch_scatter.scatter_sum(src.to(cuda), idx.to(cuda), 0, None, 20)
    want, _ = oracle.scatter(src, idx, 0, 20, "sum")
    assert torch.allclose(out.cpu(), want, rtol=1e-5, atol=1e-5)
    v, a = torch.ops.torch_scatter.scatter_max(src.to(cuda), idx.to(cuda), 0, None, 20)
    wv, wa = oracle.scatter(src, idx, 0, 20, "max")
    assert torch.equal(v.cpu(), wv) and torch.equal(a.cpu(), wa)

    @torch.jit.script
    def scripted(s, i):
        return torch.ops.torch_scatter.scatter_sum(s, i, 0, None, 20)

    assert torch.equal(scripted(src.to(cuda), idx.to(cuda)), out)
    # torch_sparse CSR ops
    rowptr = torch.tensor([0, 2, 2, 5, 9])
    col = torch.tensor([1, 3, 0, 2, 3, 0, 1, 2, 3])
    val = torch.rand(9, generator=g)
    mat = torch.randn(4, 6, generator=g)
    got = torch.ops.torch_sparse.spmm_sum(None, rowptr.to(cuda), col.to(cuda), val.to(cuda), None, None, mat.to(cuda))
    want, _ = oracle.spmm_csr(rowptr, col, val, mat, "sum")
    assert torch.allclose(got.cpu(), want, rtol=1e-5, atol=1e-5)
    assert torch.ops.torch_sparse.ind2ptr(torch.tensor([0, 0, 2, 3, 3, 3]).to(cuda), 5).cpu().tolist() == [0, 2, 2, 3, 6, 6]
    assert torch.ops.torch_sparse.ptr2ind(rowptr.to(cuda), 9).cpu().tolist() == [0, 0, 2, 2, 2, 3, 3, 3, 3]


# ---- composites, segment_csr and CSR spmm are differentiable (upstream's are) -------------------
def _ref_softmax(src, index, n):
    mx = _native_scatter(src.detach(), index, 0, n, "max")
    e = (src - mx.index_select(0, index)).exp()
    return e / _native_scatter(e, index, 0, n, "sum").index_select(0, index)


@pytest.mark.parametrize("fn", ["softmax", "log_softmax", "logsumexp", "std"])
def test_composite_backward(cuda, fn):
    import torch_scatter
    g = torch.Generator().manual_seed(21)
    E, F, N = 500, 7, 23
    src = torch.randn(E, F, generator=g)
    index = torch.randint(0, N, (E,), generator=g)
    w = torch.randn(E if fn in ("softmax", "log_softmax") else N, F, generator=g)
    a = src.clone().to(cuda).requires_grad_()
    b = src.clone().double().requires_grad_()
    if fn == "softmax":
        out, ref = torch_scatter.scatter_softmax(a, index.to(cuda), dim=0), _ref_softmax(b, index, N)
    elif fn == "log_softmax":
        out, ref = torch_scatter.scatter_log_softmax(a, index.to(cuda), dim=0), _ref_softmax(b, index, N).log()
    elif fn == "logsumexp":
        out = torch_scatter.scatter_logsumexp(a, index.to(cuda), dim=0, dim_size=N)
        mx = _native_scatter(b.detach(), index, 0, N, "max")
        ref = _native_scatter((b - mx.index_select(0, index)).exp(), index, 0, N, "sum").add(1e-12).log() + mx
    else:
        out = torch_scatter.scatter_std(a, index.to(cuda), dim=0, dim_size=N)
        cnt = torch.bincount(index, minlength=N).double().view(-1, 1)
        mean = _native_scatter(b, index, 0, N, "sum") / cnt.clamp(min=1)
        d = b - mean.index_select(0, index)
        ref = (_native_scatter(d * d, index, 0, N, "sum") / ((cnt - 1).clamp(min=1) + 1e-6)).sqrt()
    assert out.requires_grad and out.grad_fn is not None
    (out * w.to(cuda)).sum().backward()
    (ref * w.double()).sum().backward()
    assert torch.allclose(out.detach().cpu().double(), ref.detach(), rtol=1e-4, atol=1e-5)
    assert torch.allclose(a.grad.cpu().double(), b.grad, rtol=1e-3, atol=1e-4), fn


@pytest.mark.parametrize("reduce", ["sum", "mean", "max", "min"])
def test_segment_csr_backward(cuda, reduce):
    import torch_scatter
    g = torch.Generator().manual_seed(22)
    E, F = 300, 5
    src = torch.rand(E, F, generator=g) + torch.arange(E).view(-1, 1) * 1e-3   # unique extremes
    indptr = torch.tensor([4, 4, 30, 31, 200, 290])                               # sub-range, empty segment
    index = torch.repeat_interleave(torch.arange(5), indptr[1:] - indptr[:-1])
    w = torch.randn(5, F, generator=g)
    a = src.clone().to(cuda).requires_grad_()
    out = torch_scatter.segment_csr(a, indptr.to(cuda), reduce=reduce)
    (out * w.to(cuda)).sum().backward()
    b = src.clone().double().requires_grad_()
    ref = _native_scatter(b[4:290], index, 0, 5, reduce)
    (ref * w.double()).sum().backward()
    assert torch.allclose(out.detach().cpu().double(), ref.detach(), rtol=1e-4, atol=1e-5)
    assert torch.allclose(a.grad.cpu().double(), b.grad, rtol=1e-4, atol=1e-5), reduce


@pytest.mark.parametrize("reduce", ["sum", "mean", "max"])
@pytest.mark.parametrize("with_value", [False, True])
def test_sparse_matmul_backward(cuda, reduce, with_value):
    """adj_t @ x must carry gradients to x (and to the edge values): a GCN layer's backward."""
    import torch_sparse
    g = torch.Generator().manual_seed(23)
    m, n, nnz, F = 40, 30, 400, 6
    row = torch.randint(0, m, (nnz,), generator=g)
    col = torch.randint(0, n, (nnz,), generator=g)
    val = torch.rand(nnz, generator=g) + 0.5
    x = torch.rand(n, F, generator=g) + torch.arange(n).view(-1, 1) * 1e-2
    w = torch.randn(m, F, generator=g)
    va = val.clone().to(cuda).requires_grad_() if with_value else None
    adj = torch_sparse.SparseTensor(row=row.to(cuda), col=col.to(cuda), value=va, sparse_sizes=(m, n))
    xa = x.clone().to(cuda).requires_grad_()
    out = adj.matmul(xa, reduce=reduce)
    (out * w.to(cuda)).sum().backward()
    xb = x.clone().double().requires_grad_()
    vb = val.clone().double().requires_grad_()
    msgs = xb.index_select(0, col) * (vb.view(-1, 1) if with_value else 1.0)
    ref = _native_scatter(msgs, row, 0, m, reduce)
    (ref * w.double()).sum().backward()
    assert torch.allclose(out.detach().cpu().double(), ref.detach(), rtol=1e-4, atol=1e-5)
    if reduce == "max":  # duplicate (row, col) entries tie: compare what does not depend on the split
        assert torch.allclose(xa.grad.cpu().double().sum(0), xb.grad.sum(0), rtol=1e-4, atol=1e-4)
    else:
        assert torch.allclose(xa.grad.cpu().double(), xb.grad, rtol=1e-4, atol=1e-4)
        if with_value:
            assert torch.allclose(va.grad.cpu().double(), vb.grad, rtol=1e-4, atol=1e-4)
    # functional COO spmm
    if reduce == "sum" and with_value:
        xa2 = x.clone().to(cuda).requires_grad_()
        va2 = val.clone().to(cuda).requires_grad_()
        o2 = torch_sparse.spmm(torch.stack([row, col]).to(cuda), va2, m, n, xa2)
        (o2 * w.to(cuda)).sum().backward()
        assert torch.allclose(xa2.grad.cpu().double(), xb.grad, rtol=1e-4, atol=1e-4)
        assert torch.allclose(va2.grad.cpu().double(), vb.grad, rtol=1e-4, atol=1e-4)


def test_sparse_tensor_keeps_duplicate_edges(cuda):
    """Upstream SparseStorage sorts by row*n+col and never merges: parallel edges of a multigraph
    count separately in nnz, matmul('sum') and matmul('mean')."""
    import torch_sparse
    row = torch.tensor([2, 0, 2, 2, 1, 0])
    col = torch.tensor([1, 3, 1, 0, 2, 3])
    adj = torch_sparse.SparseTensor(row=row.to(cuda), col=col.to(cuda), sparse_sizes=(3, 4))
    assert adj.nnz() == 6
    r, c, _ = adj.coo()
    assert r.cpu().tolist() == [0, 0, 1, 2, 2, 2] and c.cpu().tolist() == [3, 3, 2, 0, 1, 1]
    x = torch.tensor([[1.0], [10.0], [100.0], [1000.0]]).to(cuda)
    assert adj.matmul(x, "sum").cpu().view(-1).tolist() == [2000.0, 100.0, 21.0]
    assert adj.matmul(x, "mean").cpu().view(-1).tolist() == [1000.0, 100.0, 7.0]
    val = torch.tensor([1.0, 2.0, 3.0, 4.0, 5.0, 6.0]).to(cuda)
    adj = torch_sparse.SparseTensor(row=row.to(cuda), col=col.to(cuda), value=val, sparse_sizes=(3, 4))
    # stable sort: duplicates keep their input order
    assert adj.coo()[2].cpu().tolist() == [2.0, 6.0, 5.0, 4.0, 1.0, 3.0]
    # weighted extremes (value * x), torch_sparse spmm_max with edge values
    assert adj.matmul(x, "max").cpu().view(-1).tolist() == [6000.0, 500.0, 30.0]
