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
