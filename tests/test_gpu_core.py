"""GPU parity of the core path (sort → plan → segment reduce) against the CPU oracle.
Bit-exact for indices/arg/max/min; rel 1e-5 (fp32) / 1e-2 (bf16, fp16) for sum/mean/mul."""
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu

TOL = {torch.float32: 1e-5, torch.float16: 1e-2, torch.bfloat16: 1e-2}


def close(a, b, dtype, scale=None):
    """|a - b| <= tol * scale, where scale is the magnitude the rounding error is relative to:
    |b| itself, or (for sums, whose terms cancel) the sum of |terms| per output element."""
    a, b = a.float().cpu(), b.float().cpu()
    tol = TOL[dtype]
    assert a.shape == b.shape
    scale = b.abs() if scale is None else torch.maximum(scale.float().cpu(), b.abs())
    err = (a - b).abs()
    bad = err > tol * scale + 1e-30
    assert not bad.any(), f"max err/scale {(err / (scale + 1e-30)).max():.3e} > {tol}"


@pytest.mark.parametrize("n,bits", [(0, 32), (1, 32), (1000, 32), (4096, 32), (4097, 32), (100_000, 17),
                                    (1_000_003, 32), (3_000_000, 21)])
def test_sort_pairs_u32(cuda, n, bits):
    import gno_b200
    g = torch.Generator().manual_seed(n + 1)
    hi = (1 << bits) if bits < 32 else (1 << 31)
    keys = torch.randint(0, hi, (n,), generator=g, dtype=torch.int64).to(torch.int32)
    vals = torch.arange(n, dtype=torch.int32)
    k, v = gno_b200.sort_pairs(keys.to(cuda), vals.to(cuda), 0, bits)
    ref_k, ref_p = torch.sort(keys.to(torch.int64), stable=True)
    assert torch.equal(k.cpu().to(torch.int64), ref_k)
    assert torch.equal(v.cpu().to(torch.int64), ref_p)


@pytest.mark.parametrize("n,bits", [(5000, 64), (1_000_000, 40), (300_000, 64)])
def test_sort_pairs_u64(cuda, n, bits):
    import gno_b200
    g = torch.Generator().manual_seed(n)
    hi = (1 << bits) if bits < 63 else (1 << 62)
    keys = torch.randint(0, hi, (n,), generator=g, dtype=torch.int64)
    if bits == 64:
        keys[::3] = -keys[::3]  # exercise the top bit (unsigned order)
    vals = torch.arange(n, dtype=torch.int32)
    k, v = gno_b200.sort_pairs(keys.to(cuda), vals.to(cuda), 0, bits)
    perm = oracle.argsort_stable(keys)
    assert torch.equal(v.cpu().to(torch.int64), perm)
    assert torch.equal(k.cpu(), keys[perm])


@pytest.mark.parametrize("E,N", [(0, 5), (1, 1), (1000, 10), (100_000, 1000), (200_000, 150_000), (50_000, 3)])
def test_plan_build(cuda, E, N):
    import gno_b200
    g = torch.Generator().manual_seed(E + N)
    idx = torch.randint(0, N, (E,), generator=g)
    C = 64
    plan = gno_b200.build_plan(idx.to(cuda), N, chunk_len=C)
    rowptr, perm = oracle.csr_from_index(idx, N)
    assert torch.equal(plan.rowptr.cpu(), rowptr)
    assert torch.equal(plan.perm.cpu().to(torch.int64), perm)
    assert torch.equal(plan.erow.cpu().to(torch.int64), idx[perm])
    deg = rowptr[1:] - rowptr[:-1]
    assert plan.max_len == (int(deg.max()) if N else 0)
    empty = torch.nonzero(deg == 0).flatten()
    span = torch.nonzero((deg > 0) & (rowptr[:-1] // C != (rowptr[1:] - 1) // C)).flatten()
    assert plan.n_empty == empty.numel() and torch.equal(plan.zrow.cpu().to(torch.int64), empty)
    assert plan.n_span == span.numel() and torch.equal(plan.srow.cpu().to(torch.int64), span)
    assert plan.n_dropped == 0 and plan.E_valid == E


def test_plan_out_of_range(cuda):
    import gno_b200
    idx = torch.tensor([3, -1, 0, 7, 2, 3, 100], dtype=torch.int64)
    plan = gno_b200.build_plan(idx.to(cuda), 5)
    assert plan.n_dropped == 3 and plan.E_valid == 4
    assert plan.rowptr.cpu().tolist() == [0, 1, 1, 2, 4, 4]
    src = torch.arange(7, dtype=torch.float32).view(7, 1) + 1
    out = gno_b200.scatter(src.to(cuda), idx.to(cuda), 0, None, 5, "sum")
    assert out.cpu().view(-1).tolist() == [3.0, 0.0, 5.0, 7.0, 0.0]  # out-of-range entries are dropped


CASES = [  # E, N, F
    (1000, 50, 64), (1000, 50, 100), (5000, 300, 7), (5000, 10, 602), (20000, 2000, 1),
    (3000, 100, 33), (2000, 20, 256), (1000, 2000, 16), (40000, 5, 128),
]


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("reduce", ["sum", "mean", "mul", "min", "max"])
@pytest.mark.parametrize("E,N,F", CASES)
def test_scatter_1d_index(cuda, dtype, reduce, E, N, F):
    import gno_b200
    gno_b200.clear_caches()
    g = torch.Generator().manual_seed(E * 7 + N + F)
    if reduce == "mul":
        # power-of-two factors (and signs): every partial product is exact, so long rows are
        # order-independent and the result must equal the sequential product bit for bit
        p2 = 0.004 if dtype == torch.float16 else 0.02
        u = torch.rand(E, F, generator=g)
        src = torch.where(u < p2, torch.full_like(u, 2.0), torch.where(u < 2 * p2, torch.full_like(u, 0.5),
                          torch.where(u < 0.5, torch.full_like(u, -1.0), torch.ones_like(u)))).to(dtype)
    else:
        src = torch.randn(E, F, generator=g).to(dtype)
    if reduce in ("min", "max"):  # force ties so the arg rule is exercised
        src = (src * 4).round() / 4
        src = src.to(dtype)
    idx = torch.randint(0, N, (E,), generator=g)
    want, want_arg = oracle.scatter(src, idx, 0, N, reduce)
    got = gno_b200.scatter(src.to(cuda), idx.to(cuda), 0, None, N, reduce, return_arg=True)
    if reduce in ("min", "max"):
        out, arg = got
        assert torch.equal(out.cpu(), want), "min/max values must be bit-exact"
        assert torch.equal(arg.cpu(), want_arg), "arg must be bit-exact"
    else:
        if reduce == "mul":
            assert torch.equal(got.cpu(), want), "products of powers of two are exact in any order"
        scale = None
        if reduce in ("sum", "mean"):  # error of a sum is relative to the sum of |terms|
            scale = oracle.scatter(src.float().abs(), idx, 0, N, reduce)[0]
        close(got, want, dtype, scale)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("reduce", ["sum", "mean", "min", "max"])
@pytest.mark.parametrize("E,N,F,split", [(30000, 20, 100, 64), (30000, 20, 602, 128), (20000, 300, 64, 32),
                                         (50000, 3, 16, 1024), (30000, 5000, 100, 256), (70001, 40, 8, 32)])
def test_gather_scatter_split_rows(cuda, dtype, reduce, E, N, F, split):
    """Fused gather→scatter on skewed graphs for several chunk lengths: rows far longer than a
    chunk (many partials combined in order) and many rows per chunk (power-law path)."""
    import gno_b200
    from gno_b200 import plan as planmod
    gno_b200.clear_caches()
    g = torch.Generator().manual_seed(E + F)
    n_src = 500
    x = torch.randn(n_src, F, generator=g)
    if reduce in ("min", "max"):
        x = (x * 2).round() / 2
    x = x.to(dtype)
    # skewed destinations: most edges hit row 0
    dst = (torch.rand(E, generator=g) ** 4 * N).long().clamp_(0, N - 1)
    src_ids = torch.randint(0, n_src, (E,), generator=g)
    want, want_arg = oracle.gather_scatter(x, src_ids, dst, N, reduce)
    plan = planmod.build_plan(dst.to(cuda), N, chunk_len=split)
    assert plan.n_span > 0
    gidx = plan.sorted_ids(src_ids.to(cuda))
    r = gno_b200.segment_reduce(plan, x.to(cuda), reduce, gidx=gidx, eid=plan.perm,
                                want_arg=reduce in ("min", "max"), arg_fill=E)
    if reduce in ("min", "max"):
        assert torch.equal(r[0].cpu(), want)
        assert torch.equal(r[1].cpu(), want_arg)
    else:
        scale = oracle.gather_scatter(x.float().abs(), src_ids, dst, N, reduce)[0]
        close(r, want, dtype, scale)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("reduce", ["sum", "max"])
def test_two_buffer_gather(cuda, dtype, reduce):
    """gno_segment_reduce_two: gather ids below x.size(0) read x, the rest read x2 — bit-identical
    to one launch over the concatenated buffer (same plan, same order of additions)."""
    import gno_b200
    from gno_b200 import plan as planmod
    g = torch.Generator().manual_seed(11)
    N, E, F, R = 700, 90_000, 128, 5000
    dst = (torch.rand(E, generator=g) ** 3 * N).long().clamp_(0, N - 1)
    src = torch.randint(0, R, (E,), generator=g)
    x = ((torch.randn(R, F, generator=g) * 4).round() / 4).to(dtype).to(cuda)
    plan = planmod.build_plan(dst.to(cuda), N)
    gidx = plan.sorted_ids(src.to(cuda))
    arg = reduce == "max"
    whole = gno_b200.segment_reduce(plan, x, reduce, gidx=gidx, eid=plan.perm, want_arg=arg)
    for cut in (1, 1234, R - 1):
        a, b = x[:cut].contiguous(), x[cut:].contiguous()
        two = gno_b200.segment_reduce(plan, a, reduce, gidx=gidx, eid=plan.perm, want_arg=arg, x2=b)
        if arg:
            assert torch.equal(two[0], whole[0]) and torch.equal(two[1], whole[1])
        else:
            assert torch.equal(two, whole)


def test_empty_rows_and_sentinels(cuda):
    import gno_b200
    src = torch.tensor([[1.0, -2.0], [3.0, 0.5]])
    idx = torch.tensor([3, 3])
    out, arg = gno_b200.scatter(src.to(cuda), idx.to(cuda), 0, None, 6, "max", return_arg=True)
    want, want_arg = oracle.scatter(src, idx, 0, 6, "max")
    assert torch.equal(out.cpu(), want) and torch.equal(arg.cpu(), want_arg)
    assert arg.cpu()[0].tolist() == [2, 2]  # sentinel = src.size(dim)
    out = gno_b200.scatter(src.to(cuda), idx.to(cuda), 0, None, 6, "mean")
    assert torch.equal(out.cpu(), oracle.scatter(src, idx, 0, 6, "mean")[0])


def test_nan_inf_and_signed_zero(cuda):
    import gno_b200
    nan, inf = float("nan"), float("inf")
    src = torch.tensor([[nan, -inf, -0.0, 0.0, 5.0], [1.0, -inf, 0.0, -0.0, nan], [nan, -inf, -0.0, 0.0, 5.0]])
    idx = torch.tensor([0, 0, 0])
    for red in ("max", "min"):
        out, arg = gno_b200.scatter(src.to(cuda), idx.to(cuda), 0, None, 1, red, return_arg=True)
        want, want_arg = oracle.scatter(src, idx, 0, 1, red)
        assert torch.equal(out.cpu().view(torch.int32), want.view(torch.int32)), red
        assert torch.equal(arg.cpu(), want_arg), red


def test_determinism(cuda):
    import gno_b200
    g = torch.Generator().manual_seed(5)
    src = torch.randn(200_000, 64, generator=g).to(cuda)
    idx = torch.randint(0, 1000, (200_000,), generator=g).to(cuda)
    a = gno_b200.scatter(src, idx, 0, None, 1000, "sum")
    gno_b200.clear_caches()
    b = gno_b200.scatter(src, idx, 0, None, 1000, "sum")
    assert torch.equal(a, b)


def test_no_cpu_path():
    import gno_b200
    with pytest.raises(gno_b200.GnoError):
        gno_b200.scatter(torch.ones(4, 2), torch.zeros(4, dtype=torch.int64), 0, None, 2, "sum")
