"""GPU parity of the remaining operators against the CPU oracle: full-shape-index scatter
(what the reference scripts pass), last-dim forms, index_add / index_select, spmm,
segment_csr, coalesce / transpose, sort.  Index outputs bit-exact; float sums to tolerance."""
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu

TOL = {torch.float32: 1e-5, torch.float16: 1e-2, torch.bfloat16: 1e-2}


def close(a, b, dtype, scale=None):
    a, b = a.float().cpu(), b.float().cpu()
    tol = TOL[dtype]
    assert a.shape == b.shape
    scale = b.abs() if scale is None else torch.maximum(scale.float().cpu(), b.abs())
    err = (a - b).abs()
    assert not (err > tol * scale + 1e-30).any(), f"max err/scale {(err / (scale + 1e-30)).max():.3e}"


# ---- the reference scripts' call shape: 2-D fp16 src, full-shape int64 index, dim 0/1 -----------
# (op_bm_scripts/benchmark_scatter_add.py:60-84; same in _max/_min/_mean)
@pytest.mark.parametrize("dtype", [torch.float16, torch.float32, torch.bfloat16])
@pytest.mark.parametrize("reduce", ["sum", "mean", "mul", "min", "max"])
@pytest.mark.parametrize("L,rf,dim", [(223, 1, 0), (223, 1, 1), (300, 4, 0), (300, 8, 1), (64, 2, -1)])
def test_scatter_full_shape_index(cuda, dtype, reduce, L, rf, dim):
    import torch_scatter
    g = torch.Generator().manual_seed(L * 10 + rf)
    src = torch.rand(L, L, generator=g)
    if reduce == "mul":
        src = src * 0.5 + 0.75
    if reduce in ("min", "max"):
        src = (src * 16).round() / 16
    src = src.to(dtype)
    idx = torch.randint(0, max(L // rf, 1), (L, L), generator=g)
    fn = getattr(torch_scatter, "scatter_" + {"sum": "add"}.get(reduce, reduce))
    got = fn(src.to(cuda), idx.to(cuda), dim=dim)
    want, want_arg = oracle.scatter(src, idx, dim, None, reduce)
    if reduce in ("min", "max"):
        assert isinstance(got, tuple)
        assert torch.equal(got[0].cpu(), want)
        assert torch.equal(got[1].cpu(), want_arg)
    else:
        scale = oracle.scatter(src.float().abs(), idx, dim, None, reduce)[0] if reduce != "mul" else None
        close(got, want, dtype, scale)


def test_scatter_full_shape_3d(cuda):
    import gno_b200
    g = torch.Generator().manual_seed(3)
    src = torch.randn(5, 40, 7, generator=g)
    idx = torch.randint(0, 9, (5, 40, 7), generator=g)
    for red in ("sum", "max", "min", "mean"):
        got = gno_b200.scatter(src.to(cuda), idx.to(cuda), 1, None, 9, red, return_arg=True)
        want, warg = oracle.scatter(src, idx, 1, 9, red)
        if red in ("max", "min"):
            assert torch.equal(got[0].cpu(), want) and torch.equal(got[1].cpu(), warg)
        else:
            close(got, want, torch.float32, oracle.scatter(src.abs(), idx, 1, 9, red)[0])


# ---- 1-D index along other dims ----------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.float16, torch.float32])
@pytest.mark.parametrize("reduce", ["sum", "mean", "max", "min", "mul"])
def test_scatter_1d_index_lastdim(cuda, dtype, reduce):
    import gno_b200
    g = torch.Generator().manual_seed(11)
    B, E, N = 37, 500, 41
    src = torch.rand(B, E, generator=g)
    if reduce in ("min", "max"):
        src = (src * 8).round() / 8
    if reduce == "mul":
        src = src * 0.5 + 0.75
    src = src.to(dtype)
    idx = torch.randint(0, N, (E,), generator=g)
    got = gno_b200.scatter(src.to(cuda), idx.to(cuda), 1, None, N, reduce, return_arg=True)
    want, warg = oracle.scatter(src, idx, 1, N, reduce)
    if reduce in ("min", "max"):
        assert torch.equal(got[0].cpu(), want) and torch.equal(got[1].cpu(), warg)
    else:
        close(got, want, dtype, oracle.scatter(src.float().abs(), idx, 1, N, reduce)[0] if reduce != "mul" else None)


def test_scatter_1d_index_middle_dim(cuda):
    import gno_b200
    g = torch.Generator().manual_seed(12)
    src = torch.randn(3, 200, 6, generator=g)
    idx = torch.randint(0, 17, (200,), generator=g)
    for red in ("sum", "max"):
        got = gno_b200.scatter(src.to(cuda), idx.to(cuda), 1, None, 17, red, return_arg=True)
        want, warg = oracle.scatter(src, idx, 1, 17, red)
        if red == "max":
            assert torch.equal(got[0].cpu(), want) and torch.equal(got[1].cpu(), warg)
        else:
            close(got, want, torch.float32, oracle.scatter(src.abs(), idx, 1, 17, red)[0])


def test_scatter_expanded_index_and_dim_size_none(cuda):
    """PyG-style call: index expanded with stride 0, dim_size=None → index.max()+1."""
    import torch_scatter
    g = torch.Generator().manual_seed(13)
    src = torch.randn(300, 12, generator=g)
    idx = torch.randint(0, 20, (300,), generator=g)
    got = torch_scatter.scatter(src.to(cuda), idx.to(cuda).view(-1, 1).expand(-1, 12), dim=0, reduce="sum")
    want, _ = oracle.scatter(src, idx, 0, None, "sum")
    assert got.shape == want.shape
    close(got, want, torch.float32, oracle.scatter(src.abs(), idx, 0, None, "sum")[0])
    # out= accumulation (index_add_-like)
    base = torch.randn(20, 12, generator=g)
    got = torch_scatter.scatter_add(src.to(cuda), idx.to(cuda), dim=0, out=base.clone().to(cuda))
    close(got, base + want, torch.float32, base.abs() + oracle.scatter(src.abs(), idx, 0, 20, "sum")[0])


# ---- native-torch spellings: index_add_, index_select -----------------------------------------
@pytest.mark.parametrize("dtype", [torch.float16, torch.float32])
@pytest.mark.parametrize("dim", [0, 1])
def test_index_add(cuda, dtype, dim):
    import gno_b200
    g = torch.Generator().manual_seed(21 + dim)
    L = 257
    inp = torch.rand(L, L, generator=g).to(dtype)
    other = torch.rand(L, L, generator=g).to(dtype)
    idx = torch.randint(0, L, (L,), generator=g)
    got = gno_b200.index_add(inp.to(cuda), dim, idx.to(cuda), other.to(cuda))
    want = oracle.index_add(inp, dim, idx, other)
    scale = oracle.index_add(inp.float().abs(), dim, idx, other.float().abs())
    close(got, want, dtype, scale)
    native = torch.index_add(inp.float(), dim, idx, other.float())
    close(got, native, dtype, scale)


@pytest.mark.parametrize("dtype", [torch.float16, torch.float32, torch.bfloat16])
@pytest.mark.parametrize("dim", [0, 1])
@pytest.mark.parametrize("L,rf", [(129, 1), (250, 4), (67, 8)])
def test_index_select(cuda, dtype, dim, L, rf):
    import gno_b200
    g = torch.Generator().manual_seed(L)
    inp = torch.rand(L, L, generator=g).to(dtype)
    idx = torch.randint(0, L, (max(L // rf, 1),), generator=g)
    got = gno_b200.index_select(inp.to(cuda), dim, idx.to(cuda))
    assert torch.equal(got.cpu(), torch.index_select(inp, dim, idx))


# ---- spmm / segment_csr ------------------------------------------------------------------------
def _random_csr(M, Ncols, nnz, g):
    row = torch.randint(0, M, (nnz,), generator=g).sort().values
    col = torch.randint(0, Ncols, (nnz,), generator=g)
    rowptr = torch.zeros(M + 1, dtype=torch.int64)
    rowptr[1:] = torch.cumsum(torch.bincount(row, minlength=M), 0)
    return row, col, rowptr


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("reduce", ["sum", "mean", "min", "max"])
@pytest.mark.parametrize("M,F,nnz", [(100, 256, 5000), (1000, 33, 20000), (10, 64, 40000)])
def test_spmm_csr(cuda, dtype, reduce, M, F, nnz):
    import gno_b200
    gno_b200.clear_caches()
    g = torch.Generator().manual_seed(M + F)
    row, col, rowptr = _random_csr(M, 77, nnz, g)
    mat = torch.randn(77, F, generator=g)
    if reduce in ("min", "max"):
        mat = (mat * 2).round() / 2
    mat = mat.to(dtype)
    value = torch.rand(nnz, generator=g).to(dtype) if reduce in ("sum", "mean") else None
    want, warg = oracle.spmm_csr(rowptr, col, value, mat, reduce)
    got = gno_b200.spmm_csr(rowptr.to(cuda), col.to(cuda), value.to(cuda) if value is not None else None,
                            mat.to(cuda), reduce, return_arg=True)
    if reduce in ("min", "max"):
        assert torch.equal(got[0].cpu(), want) and torch.equal(got[1].cpu(), warg)
    else:
        v = value.float().abs() if value is not None else None
        scale = oracle.spmm_csr(rowptr, col, v, mat.float().abs(), reduce)[0]
        close(got, want, dtype, scale)


def test_spmm_coo_matches_reference_formulation(cuda):
    """torch_sparse.spmm(index, value, m, n, matrix) vs the oracle and vs torch.sparse.mm
    (the call benchmark_sparse_spmm.py:12-14 times)."""
    import torch_sparse
    g = torch.Generator().manual_seed(5)
    m, n, F, nnz = 300, 200, 48, 6000
    index = torch.stack([torch.randint(0, m, (nnz,), generator=g), torch.randint(0, n, (nnz,), generator=g)])
    value = torch.rand(nnz, generator=g)
    mat = torch.randn(n, F, generator=g)
    got = torch_sparse.spmm(index.to(cuda), value.to(cuda), m, n, mat.to(cuda))
    want = oracle.spmm(index, value, m, n, mat)
    native = torch.sparse.mm(torch.sparse_coo_tensor(index, value, (m, n)), mat)
    scale = oracle.spmm(index, value, m, n, mat.abs())
    close(got, want, torch.float32, scale)
    close(got, native, torch.float32, scale)


@pytest.mark.parametrize("reduce", ["sum", "mean", "min", "max"])
def test_segment_csr(cuda, reduce):
    import torch_scatter
    g = torch.Generator().manual_seed(8)
    src = (torch.randn(1000, 5, generator=g) * 4).round() / 4
    indptr = torch.tensor([0, 0, 10, 10, 500, 999, 1000])
    want, warg = oracle.spmm_csr(indptr, torch.arange(1000), None, src, reduce)
    got = torch_scatter.segment_csr(src.to(cuda), indptr.to(cuda), reduce=reduce)
    if reduce in ("min", "max"):
        assert torch.equal(got.cpu(), want)
        v, a = torch_scatter.segment_max_csr(src.to(cuda), indptr.to(cuda)) if reduce == "max" else \
            torch_scatter.segment_min_csr(src.to(cuda), indptr.to(cuda))
        assert torch.equal(v.cpu(), want) and torch.equal(a.cpu(), warg)
    else:
        close(got, want, torch.float32, oracle.spmm_csr(indptr, torch.arange(1000), None, src.abs(), reduce)[0])


# ---- coalesce / transpose ----------------------------------------------------------------------
def _dup_coo(m, n, nnz, rf, g):
    """COO with duplicates built like benchmark_sparse_coalesce.py:129-159 (index permuted, value not)."""
    index_ = torch.stack([torch.randint(0, m, (nnz,), generator=g), torch.randint(0, n, (nnz,), generator=g)])
    value_ = torch.rand(nnz, generator=g)
    index = torch.cat([index_] * rf, dim=1)
    value = torch.cat([value_] * rf)
    index = index.index_select(1, torch.randperm(index.shape[1], generator=g))
    return index, value


@pytest.mark.parametrize("m,n,nnz,rf", [(100, 100, 2000, 2), (5000, 3, 20000, 4), (1 << 20, 1 << 20, 50000, 8),
                                        (7, 100000, 30000, 1)])
@pytest.mark.parametrize("op", ["add", "mean", "max"])
def test_coalesce(cuda, m, n, nnz, rf, op):
    import torch_sparse
    g = torch.Generator().manual_seed(nnz + rf)
    index, value = _dup_coo(m, n, nnz, rf, g)
    gi, gv = torch_sparse.coalesce(index.to(cuda), value.to(cuda), m, n, op)
    wi, wv = oracle.coalesce(index, value, m, n, op)
    assert torch.equal(gi.cpu(), wi), "coalesced indices must be bit-exact"
    close(gv, wv, torch.float32, oracle.coalesce(index, value.abs(), m, n, "add" if op == "add" else op)[1])
    if op == "add":  # independent formulation: Tensor.coalesce()
        nat = torch.sparse_coo_tensor(index, value, (m, n)).coalesce()
        assert torch.equal(gi.cpu(), nat.indices())
        close(gv, nat.values(), torch.float32)


def test_coalesce_edge_cases(cuda):
    import torch_sparse
    # value=None, K>1 values, already-sorted early exit returns the same tensors, empty input
    g = torch.Generator().manual_seed(1)
    index, _ = _dup_coo(50, 60, 500, 2, g)
    gi, gv = torch_sparse.coalesce(index.to(cuda), None, 50, 60)
    assert gv is None and torch.equal(gi.cpu(), oracle.coalesce(index, None, 50, 60)[0])
    val = torch.randn(index.shape[1], 3, generator=g)
    gi, gv = torch_sparse.coalesce(index.to(cuda), val.to(cuda), 50, 60)
    wi, wv = oracle.coalesce(index, val, 50, 60)
    assert torch.equal(gi.cpu(), wi)
    close(gv, wv, torch.float32, oracle.coalesce(index, val.abs(), 50, 60)[1])
    si, sv = gi, gv
    ri, rv = torch_sparse.coalesce(si, sv, 50, 60)
    assert ri.data_ptr() == si.data_ptr() and rv.data_ptr() == sv.data_ptr()
    ei, ev = torch_sparse.coalesce(torch.zeros(2, 0, dtype=torch.int64, device=cuda),
                                   torch.zeros(0, device=cuda), 5, 5)
    assert ei.shape == (2, 0) and ev.shape == (0,)


@pytest.mark.parametrize("m,n,nnz", [(300, 200, 5000), (5, 100000, 20000), (100000, 5, 20000)])
def test_transpose(cuda, m, n, nnz):
    import torch_sparse
    g = torch.Generator().manual_seed(nnz)
    index, value = _dup_coo(m, n, nnz, 1, g)
    # generic path: unsorted input with duplicates
    gi, gv = torch_sparse.transpose(index.to(cuda), value.to(cuda), m, n)
    wi, wv = oracle.transpose(index, value, m, n)
    assert torch.equal(gi.cpu(), wi)
    close(gv, wv, torch.float32)
    # fast path: coalesced input → stable sort on the new major key only; exact value permutation
    ci, cv = oracle.coalesce(index, value, m, n)
    gi, gv = torch_sparse.transpose(ci.to(cuda), cv.to(cuda), m, n)
    wi, wv = oracle.transpose(ci, cv, m, n)
    assert torch.equal(gi.cpu(), wi)
    assert torch.equal(gv.cpu(), wv), "transposing a coalesced matrix only permutes values"
    # involution
    bi, bv = torch_sparse.transpose(gi, gv, n, m)
    assert torch.equal(bi.cpu(), ci) and torch.equal(bv.cpu(), cv)
    ui, uv = torch_sparse.transpose(index.to(cuda), value.to(cuda), m, n, coalesced=False)
    assert torch.equal(ui.cpu(), index.flip(0)) and torch.equal(uv.cpu(), value)


# ---- sort --------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape,dim", [((100_003,), 0), ((300, 500), 1), ((300, 500), 0), ((20, 30, 40), 1),
                                       ((20, 30, 40), -1), ((1,), 0), ((5, 1), 0)])
@pytest.mark.parametrize("descending", [False, True])
def test_sort_f32(cuda, shape, dim, descending):
    import gno_b200
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(*shape, generator=g)
    x = torch.where(torch.rand(*shape, generator=g) < 0.3, (x * 2).round() / 2, x)  # ties
    flat = x.view(-1)
    flat[::17] = 0.0
    flat[5::29] = -0.0
    flat[3::41] = float("nan")
    flat[7::97] = float("inf")
    flat[9::101] = -float("inf")
    v, i = gno_b200.sort(x.to(cuda), dim, descending)
    wv, wi = oracle.sort(x, dim, descending)
    assert torch.equal(i.cpu(), wi), "sort indices must be bit-exact"
    assert torch.equal(v.cpu().view(torch.int32), wv.view(torch.int32)), "sorted values must be bit-exact"
    tv, ti = torch.sort(x, dim=dim, descending=descending, stable=True)
    assert torch.equal(i.cpu(), ti)
    assert torch.equal(v.cpu().view(torch.int32), tv.view(torch.int32))


# ---- golden vectors made by the reference's own op functions -----------------------------------
def test_golden_native_ops(cuda):
    """tests/golden/native_ops.npz was produced by executing the op functions defined in the
    reference scripts (see tests/golden/make_golden.py) on seeded CPU inputs."""
    import os
    import numpy as np
    import gno_b200
    path = os.path.join(os.path.dirname(__file__), "golden", "native_ops.npz")
    z = np.load(path)
    t = lambda k: torch.from_numpy(z[k])  # noqa: E731
    # op_native_sort (benchmark_native_sort.py:28-30)
    for tag in ("sort1d", "sort2d_d0", "sort2d_d1"):
        dim = int(z[tag + "_dim"])
        v, i = gno_b200.sort(t(tag + "_in").to(cuda), dim)
        assert torch.equal(i.cpu(), t(tag + "_idx")) and torch.equal(v.cpu(), t(tag + "_val")), tag
    # op_native_index_add_ (benchmark_native_index_add_.py:13-16), fp16 like the script
    got = gno_b200.index_add(t("iadd_in").to(cuda), 1, t("iadd_index").to(cuda), t("iadd_src").to(cuda))
    close(got, t("iadd_out"), torch.float16, t("iadd_scale"))
    # index_select (benchmark_native_index_select.py:12-15)
    for dim in (0, 1):
        got = gno_b200.index_select(t("isel_in").to(cuda), dim, t("isel_index").to(cuda))
        assert torch.equal(got.cpu(), t(f"isel_out_d{dim}"))
    # fused index_select→sum (benchmark_fused_index_select_reduce.py:12-15)
    got = gno_b200.index_select(t("isel_in").to(cuda), 0, t("isel_index").to(cuda)).float().sum()
    assert abs(float(got) - float(z["isel_sum_d0"])) <= 1e-2 * float(z["isel_abs_sum_d0"])
    # op_native_smm (benchmark_sparse_spmm.py:12-14)
    got = gno_b200.spmm(t("smm_index").to(cuda), t("smm_value").to(cuda), int(z["smm_m"]), int(z["smm_n"]),
                        t("smm_B").to(cuda))
    close(got, t("smm_out"), torch.float32, t("smm_scale"))
    # op_native_coalesce (benchmark_sparse_coalesce.py:40-42)
    gi, gv = gno_b200.coalesce(t("coal_index").to(cuda), t("coal_value").to(cuda), int(z["coal_m"]), int(z["coal_n"]))
    assert torch.equal(gi.cpu(), t("coal_out_index"))
    close(gv, t("coal_out_value"), torch.float32)
    # op_native_scatter_add_ / scatter_(reduce=multiply) (benchmark_scatter_add.py:22-25, _multiply.py:42-45)
    got = gno_b200.scatter(t("sadd_src").to(cuda), t("sadd_idx").to(cuda), 0, None, t("sadd_src").shape[0], "sum")
    close(got, t("sadd_out"), torch.float16, t("sadd_scale"))
    got = gno_b200.scatter(t("smul_src").to(cuda), t("smul_idx").to(cuda), -1, None, t("smul_src").shape[-1], "mul")
    close(got, t("smul_out"), torch.float32)


# ---- host-buffer API (the e2e path of bench.py) -------------------------------------------------
def test_host_api_and_pipeline(cuda):
    from gno_b200.host import HostPipeline, gather_scatter_host
    g = torch.Generator().manual_seed(77)
    N, E, F = 4000, 150_000, 100
    x = torch.randn(N, F, generator=g).pin_memory()
    ei = torch.stack([torch.randint(0, N, (E,), generator=g),
                      (torch.rand(E, generator=g) ** 3 * N).long().clamp_(0, N - 1)]).pin_memory()
    want, _ = oracle.gather_scatter(x, ei[0], ei[1], N, "sum")
    scale = oracle.gather_scatter(x.abs(), ei[0], ei[1], N, "sum")[0]
    got = gather_scatter_host(x, ei, N, "sum")
    assert not got.is_cuda
    close(got, want, torch.float32, scale)
    pipe = HostPipeline(N, "sum")
    outs = []
    pipe.submit(x, ei)
    for _ in range(3):
        pipe.submit(x, ei)
        outs.append(pipe.result())
    outs.append(pipe.result())
    for o in outs:
        assert torch.equal(o, got), "pipelined results must be bit-identical (deterministic kernels)"


def test_cuda_graph_capture(cuda):
    """The C-ABI never allocates or synchronises, so a warm-plan aggregation (2 launches) can be
    captured in a CUDA graph and replayed — how launch-bound small calls (C1) should be driven."""
    import gno_b200
    g = torch.Generator().manual_seed(5)
    E, N, F = 50_000, 3000, 64
    src = torch.randn(E, F, generator=g).to(cuda)
    idx = torch.randint(0, N, (E,), generator=g).to(cuda)
    want = gno_b200.scatter(src, idx, 0, None, N, "sum").clone()  # builds and caches the plan
    plan = gno_b200.plan_cache.get(idx, N)
    out = torch.empty(N, F, device=cuda)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        gno_b200.segment_reduce(plan, src, "sum", gidx=plan.perm, out=out)  # warm-up on the side stream
    torch.cuda.current_stream().wait_stream(s)
    graph = torch.cuda.CUDAGraph()
    before = gno_b200.launch_count()
    with torch.cuda.graph(graph):
        gno_b200.segment_reduce(plan, src, "sum", gidx=plan.perm, out=out)
    assert gno_b200.launch_count() - before == 2
    out.zero_()
    src.mul_(2.0)
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, gno_b200.scatter(src, idx, 0, None, N, "sum"))
    assert torch.allclose(out, 2 * want, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("reduce", ["sum", "max", "mean"])
def test_scatter_full_shape_column_blocked(cuda, reduce):
    """Wide dim-0 scatter: the accumulator exceeds the L2 budget, so the kernels traverse the
    elements column block by column block (uneven last block)."""
    import gno_b200
    g = torch.Generator().manual_seed(9)
    E, K, N = 300, 4100, 1600
    src = (torch.randn(E, K, generator=g) * 8).round() / 8
    idx = torch.randint(0, N, (E, K), generator=g)
    got = gno_b200.scatter(src.to(cuda), idx.to(cuda), 0, None, N, reduce, return_arg=True)
    want, warg = oracle.scatter(src, idx, 0, N, reduce)
    if reduce == "max":
        assert torch.equal(got[0].cpu(), want) and torch.equal(got[1].cpu(), warg)
    else:
        close(got, want, torch.float32, oracle.scatter(src.abs(), idx, 0, N, reduce)[0])


# ---- widened surface: composites and SparseTensor ---------------------------------
def test_scatter_composites(cuda):
    import torch_scatter
    g = torch.Generator().manual_seed(31)
    src = torch.randn(500, 6, generator=g)
    idx = torch.randint(0, 23, (500,), generator=g)
    sm = torch_scatter.scatter_softmax(src.to(cuda), idx.to(cuda), dim=0).cpu()
    lsm = torch_scatter.scatter_log_softmax(src.to(cuda), idx.to(cuda), dim=0).cpu()
    lse = torch_scatter.scatter_logsumexp(src.to(cuda), idx.to(cuda), dim=0, dim_size=23).cpu()
    std = torch_scatter.scatter_std(src.to(cuda), idx.to(cuda), dim=0, dim_size=23).cpu()
    for i in range(23):
        rows = (idx == i).nonzero().flatten()
        if rows.numel() == 0:
            continue
        ref = torch.softmax(src[rows], dim=0)
        assert torch.allclose(sm[rows], ref, rtol=1e-4, atol=1e-6)
        assert torch.allclose(lsm[rows], torch.log_softmax(src[rows], dim=0), rtol=1e-4, atol=1e-5)
        assert torch.allclose(lse[i], torch.logsumexp(src[rows], dim=0), rtol=1e-4, atol=1e-5)
        if rows.numel() > 1:
            assert torch.allclose(std[i], src[rows].std(dim=0), rtol=1e-3, atol=1e-4)


def test_sparse_tensor(cuda):
    from torch_sparse import SparseTensor, matmul
    g = torch.Generator().manual_seed(32)
    m, n, nnz, F = 120, 90, 3000, 24
    ei = torch.stack([torch.randint(0, m, (nnz,), generator=g), torch.randint(0, n, (nnz,), generator=g)])
    val = torch.rand(nnz, generator=g)
    x = torch.randn(n, F, generator=g)
    ci, cv = oracle.coalesce(ei, val, m, n)
    # unique entries in random order (SparseTensor sorts, and — like upstream — never merges
    # duplicates: test_sparse_tensor_keeps_duplicate_edges covers those)
    shuffle = torch.randperm(ci.size(1), generator=g)
    A = SparseTensor.from_edge_index(ci[:, shuffle].to(cuda), cv[shuffle].to(cuda), sparse_sizes=(m, n))
    dense = torch.zeros(m, n)
    dense[ci[0], ci[1]] = cv
    assert A.nnz() == ci.size(1)
    assert torch.allclose(A.to_dense().cpu(), dense, rtol=1e-5, atol=1e-6)
    assert torch.allclose((A @ x.to(cuda)).cpu(), dense @ x, rtol=1e-4, atol=1e-4)
    assert torch.allclose(matmul(A, x.to(cuda), "mean").cpu(),
                          (dense @ x) / (dense != 0).sum(1).clamp(min=1).view(-1, 1), rtol=1e-4, atol=1e-4)
    At = A.t()
    assert At.sparse_sizes() == (n, m) and At.t() is A
    y = torch.randn(m, F, generator=g)
    assert torch.allclose((At @ y.to(cuda)).cpu(), dense.t() @ y, rtol=1e-4, atol=1e-4)
    assert torch.allclose(A.sum(dim=1).cpu(), dense.sum(1), rtol=1e-5, atol=1e-5)
    assert torch.allclose(A.sum(dim=0).cpu(), dense.sum(0), rtol=1e-5, atol=1e-5)
    rowptr, col, v = A.csr()
    assert rowptr.numel() == m + 1 and int(rowptr[-1]) == A.nnz()


# ---- segment_coo / gather_coo (torch_scatter's sorted-index forms, reached through PyG) ----------
@pytest.mark.parametrize("reduce", ["sum", "mean", "min", "max"])
@pytest.mark.parametrize("batched", [False, True])
def test_segment_coo(cuda, reduce, batched):
    import torch_scatter
    g = torch.Generator().manual_seed(21)
    E, K, N = 700, 6, 40
    if batched:
        src = (torch.randn(3, E, K, generator=g) * 4).round() / 4
        index = torch.sort(torch.randint(0, N, (3, E), generator=g), dim=1).values
        full = index.unsqueeze(-1).expand_as(src)
        want, warg = oracle.scatter(src, full, 1, N, reduce)
    else:
        src = (torch.randn(E, K, generator=g) * 4).round() / 4
        index = torch.sort(torch.randint(0, N, (E,), generator=g)).values
        want, warg = oracle.scatter(src, index, 0, N, reduce)
    got = torch_scatter.segment_coo(src.to(cuda), index.to(cuda), dim_size=N, reduce=reduce)
    if reduce in ("min", "max"):
        assert torch.equal(got.cpu(), want)
        fn = torch_scatter.segment_max_coo if reduce == "max" else torch_scatter.segment_min_coo
        v, a = fn(src.to(cuda), index.to(cuda), None, N)
        assert torch.equal(v.cpu(), want) and torch.equal(a.cpu(), warg)
    else:
        close(got, want, torch.float32, None if reduce == "mean" else
              oracle.scatter(src.abs(), full if batched else index, 1 if batched else 0, N, "sum")[0])
    # dim_size=None: last segment id + 1 (one host sync, as upstream)
    got2 = torch_scatter.segment_coo(src.to(cuda), index.to(cuda), reduce="sum")
    assert got2.size(index.dim() - 1) == int(index.max()) + 1


def test_gather_coo(cuda):
    import torch_scatter
    g = torch.Generator().manual_seed(22)
    src = torch.randn(50, 7, generator=g)
    index = torch.sort(torch.randint(0, 50, (300,), generator=g)).values
    got = torch_scatter.gather_coo(src.to(cuda), index.to(cuda))
    assert torch.equal(got.cpu(), src.index_select(0, index))
    src3 = torch.randn(4, 50, 3, generator=g)
    index2 = torch.sort(torch.randint(0, 50, (4, 120), generator=g), dim=1).values
    got = torch_scatter.gather_coo(src3.to(cuda), index2.to(cuda))
    want = src3.gather(1, index2.unsqueeze(-1).expand(4, 120, 3))
    assert torch.equal(got.cpu(), want)
    src1 = torch.randn(64, generator=g)
    got = torch_scatter.gather_coo(src1.to(cuda), index[:100].clamp(max=63).to(cuda))
    assert torch.equal(got.cpu(), src1[index[:100].clamp(max=63)])


@pytest.mark.parametrize("outer,rows,cols", [(1, 33, 65), (3, 100, 7), (2, 1, 50), (1, 1000, 1000)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.int64])
def test_transpose_batched(cuda, outer, rows, cols, dtype):
    from gno_b200 import ops
    x = torch.arange(outer * rows * cols).view(outer, rows, cols).to(dtype)
    got = ops._transpose_batched(x.to(cuda), outer, rows, cols).view(outer, cols, rows)
    assert torch.equal(got.cpu(), x.transpose(1, 2).contiguous())


# ---- graph construction (fakeDatasets.py:238-259: remove_self_loops → to_undirected | coalesce) --
def test_graph_construction(cuda):
    from gno_b200 import graph
    g = torch.Generator().manual_seed(31)
    n = 60
    ei = torch.stack([torch.randint(0, n, (900,), generator=g), torch.randint(0, n, (900,), generator=g)])
    no_loops, _ = graph.remove_self_loops(ei.to(cuda))
    assert torch.equal(no_loops.cpu(), oracle.remove_self_loops(ei))
    und = graph.to_undirected(no_loops, num_nodes=n)
    assert torch.equal(und.cpu(), oracle.to_undirected(oracle.remove_self_loops(ei), n))
    co = graph.coalesce(ei.to(cuda), num_nodes=n)
    assert torch.equal(co.cpu(), oracle.coalesce_edges(ei, n))
    # edge attributes ride along: both directions carry the value, duplicates add up
    w = torch.rand(900, generator=g)
    ui, uw = graph.to_undirected(ei.to(cuda), w.to(cuda), num_nodes=n)
    both = torch.cat([ei, ei.flip(0)], dim=1)
    wi, wv = oracle.coalesce(both, torch.cat([w, w]), n, n)
    assert torch.equal(ui.cpu(), wi)
    close(uw, wv, torch.float32)
    # the reference's generator end to end: sorted, unique, symmetric, loop-free
    gd = torch.Generator(device=cuda).manual_seed(5)
    fe = graph.fake_edge_index(1000, 1000, 10, is_undirected=True, remove_loops=True, device=cuda, generator=gd)
    key = fe[0] * 1000 + fe[1]
    assert (key[1:] > key[:-1]).all() and (fe[0] != fe[1]).all()
    rkey, _ = torch.sort(fe[1] * 1000 + fe[0])
    assert torch.equal(rkey, key), "an undirected edge list equals its own transpose"
    # collation: node ids shifted per graph, batch vector for global_mean_pool
    e1 = torch.tensor([[0, 1], [1, 2]], device=cuda)
    e2 = torch.tensor([[0], [1]], device=cuda)
    bi, batch = graph.collate([e1, e2], [3, 2])
    assert bi.tolist() == [[0, 1, 3], [1, 2, 4]] and batch.tolist() == [0, 0, 0, 1, 1]


# ---- known-answer vectors published by the upstream packages (READMEs) ---------------------------
def test_upstream_published_examples(cuda):
    """scatter_max (pytorch_scatter README) and coalesce / transpose / spmm (pytorch_sparse README)
    through the shim packages on the GPU; tests/golden/upstream_published.py."""
    import importlib.util
    import os
    import torch_scatter
    import torch_sparse
    spec = importlib.util.spec_from_file_location(
        "upstream_published", os.path.join(os.path.dirname(__file__), "golden", "upstream_published.py"))
    up = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(up)
    v = up.SCATTER_MAX
    out, arg = torch_scatter.scatter_max(v["src"].to(cuda), v["index"].to(cuda), dim=-1)
    assert torch.equal(out.cpu(), v["out"]) and torch.equal(arg.cpu(), v["arg"])
    v = up.COALESCE
    i, val = torch_sparse.coalesce(v["index"].to(cuda), v["value"].to(cuda), m=v["m"], n=v["n"])
    assert torch.equal(i.cpu(), v["out_index"]) and torch.equal(val.cpu(), v["out_value"])
    v = up.TRANSPOSE
    i, val = torch_sparse.transpose(v["index"].to(cuda), v["value"].to(cuda), v["m"], v["n"])
    assert torch.equal(i.cpu(), v["out_index"]) and torch.equal(val.cpu(), v["out_value"])
    v = up.SPMM
    out = torch_sparse.spmm(v["index"].to(cuda), v["value"].to(cuda), v["m"], v["n"], v["matrix"].to(cuda))
    assert torch.equal(out.cpu(), v["out"])


def test_upstream_testsuite_tables(cuda):
    """The `tests = [...]` tables of pytorch_scatter's own test suite (tests/golden/upstream_testsuite.py)
    through the shim package: scatter_{sum,mul,mean,min,max} for every index shape class (1-D, 1-D
    over rows, full-shape, fewer dims than src), segment_coo and segment_csr (shared and per-row
    pointers).  Called twice so the cached-plan path of a full-shape index is covered too."""
    import importlib.util
    import os
    import torch_scatter
    spec = importlib.util.spec_from_file_location(
        "upstream_testsuite", os.path.join(os.path.dirname(__file__), "golden", "upstream_testsuite.py"))
    ut = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ut)
    for v in ut.SCATTER:
        src, index = v["src"].to(cuda), v["index"].to(cuda)
        for _ in range(2):
            assert torch.equal(torch_scatter.scatter_sum(src, index, v["dim"]).cpu(), v["sum"])
            assert torch.equal(torch_scatter.scatter_add(src, index, v["dim"]).cpu(), v["sum"])
            assert torch.equal(torch_scatter.scatter_mul(src, index, v["dim"]).cpu(), v["mul"])
            assert torch.equal(torch_scatter.scatter_mean(src, index, v["dim"]).cpu(), v["mean"])
            for red, fn in (("min", torch_scatter.scatter_min), ("max", torch_scatter.scatter_max)):
                out, arg = fn(src, index, v["dim"])
                assert torch.equal(out.cpu(), v[red]) and torch.equal(arg.cpu(), v["arg_" + red]), red
                assert torch.equal(torch_scatter.scatter(src, index, v["dim"], reduce=red).cpu(), v[red])
    for v in ut.SEGMENT:
        src, index, indptr = v["src"].to(cuda), v["index"].to(cuda), v["indptr"].to(cuda)
        for red in ("sum", "mean"):
            assert torch.equal(torch_scatter.segment_coo(src, index, reduce=red).cpu(), v[red]), red
            assert torch.equal(torch_scatter.segment_csr(src, indptr, reduce=red).cpu(), v[red]), red
        for red in ("min", "max"):
            out, arg = getattr(torch_scatter, f"segment_{red}_coo")(src, index)
            assert torch.equal(out.cpu(), v[red]) and torch.equal(arg.cpu(), v["arg_" + red]), red
            out, arg = getattr(torch_scatter, f"segment_{red}_csr")(src, indptr)
            assert torch.equal(out.cpu(), v[red]) and torch.equal(arg.cpu(), v["arg_" + red]), red
    for v in ut.GATHER:
        src, index, indptr = v["src"].to(cuda), v["index"].to(cuda), v["indptr"].to(cuda)
        assert torch.equal(torch_scatter.gather_coo(src, index).cpu(), v["expected"])
        assert torch.equal(torch_scatter.gather_csr(src, indptr).cpu(), v["expected"])


# ---- full-shape index: the one-launch shared-memory path and its edges --------------------------
def _ref_minmax_with_out(m, a, out0, E, red):
    """torch_scatter's out= form for min/max from the oracle's fresh result: out0 is the starting
    value and survives (arg = E) unless an element beats it strictly."""
    win = (a != E) & ((m > out0) if red == "max" else (m < out0))
    return torch.where(win, m, out0), torch.where(win, a, torch.full_like(a, E))


@pytest.mark.parametrize("dim", [0, 1])
@pytest.mark.parametrize("dtype", [torch.float16, torch.float32, torch.bfloat16])
def test_scatter_full_shape_special_values(cuda, dim, dtype):
    """NaN / inf / signed zeros / the dtype's finite extremes through the on-chip path: fp16 sums
    leave the exact fixed-point bins for fp32 ones, extremes never win a max/min, ties keep the
    first position and the winner keeps its own sign bit."""
    import gno_b200
    g = torch.Generator().manual_seed(17)
    L = 97
    src = ((torch.rand(L, L, generator=g) * 8).round() / 8 - 0.5).to(dtype)
    fin = torch.finfo(dtype)
    specials = torch.tensor([float("nan"), float("inf"), float("-inf"), 0.0, -0.0, fin.max, fin.min], dtype=dtype)
    pos = torch.randint(0, L * L, (60,), generator=g)
    src.view(-1)[pos] = specials[torch.randint(0, specials.numel(), (60,), generator=g)]
    idx = torch.randint(0, 23, (L, L), generator=g)
    for red in ("sum", "mean", "max", "min"):
        got = gno_b200.scatter(src.to(cuda), idx.to(cuda), dim, None, 23, red, return_arg=True)
        want, warg = oracle.scatter(src, idx, dim, 23, red)
        if red in ("max", "min"):
            assert torch.equal(got[0].cpu().view(torch.int16 if dtype != torch.float32 else torch.int32),
                               want.view(torch.int16 if dtype != torch.float32 else torch.int32)), red
            assert torch.equal(got[1].cpu(), warg), red
        else:
            g_, w_ = got.float().cpu(), want.float()
            assert torch.equal(torch.isnan(g_), torch.isnan(w_)) and torch.equal(torch.isinf(g_), torch.isinf(w_))
            ok = torch.isfinite(w_)
            assert torch.equal(torch.sign(g_[~ok & ~torch.isnan(w_)]), torch.sign(w_[~ok & ~torch.isnan(w_)]))
            fin_src = torch.where(torch.isfinite(src.float()), src.float().abs(), torch.zeros(()))
            scale = oracle.scatter(fin_src, idx, dim, 23, red)[0]
            err = (g_ - w_).abs()[ok]
            assert not (err > TOL[dtype] * torch.maximum(scale, w_.abs())[ok] + 1e-30).any(), red


def test_scatter_full_shape_fp16_sum_is_exact(cuda):
    """fp16 sums accumulate in 64-bit fixed point: the result is the correctly rounded exact sum,
    whatever the order (checked against an fp64 sum), and repeated calls are bit-identical."""
    import gno_b200
    g = torch.Generator().manual_seed(23)
    L = 512
    src = (torch.randn(L, L, generator=g) * 30).half()
    idx = torch.randint(0, 7, (L, L), generator=g)       # ~37k terms per bin: fp32 order would matter
    for dim in (0, 1):
        a = gno_b200.scatter(src.to(cuda), idx.to(cuda), dim, None, 7, "sum")
        b = gno_b200.scatter(src.to(cuda), idx.to(cuda), dim, None, 7, "sum")
        assert torch.equal(a.view(torch.int16), b.view(torch.int16))
        exact = torch.zeros(a.shape, dtype=torch.float64).scatter_add_(dim, idx, src.double())
        assert torch.equal(a.cpu().view(torch.int16), exact.half().view(torch.int16))


@pytest.mark.parametrize("shape,dim,N", [((70_000, 3), 0, 11), ((4, 100_000), 1, 60_000), ((3, 500, 20), 1, 40),
                                         ((2, 3, 70_000), 2, 9)])
def test_scatter_full_shape_path_selection(cuda, shape, dim, N):
    """Positions beyond 16 bits (64-bit packed keys), N too large for shared-memory bins (L2-atomic
    fallback), 3-D inputs: all against the oracle."""
    import gno_b200
    g = torch.Generator().manual_seed(sum(shape))
    src = ((torch.rand(*shape, generator=g) * 64).round() / 64).half()
    idx = torch.randint(0, N, shape, generator=g)
    for red in ("sum", "max", "mean", "min", "mul"):
        s = (src * 0.25 + 0.875) if red == "mul" else src
        got = gno_b200.scatter(s.to(cuda), idx.to(cuda), dim, None, N, red, return_arg=True)
        want, warg = oracle.scatter(s, idx, dim, N, red)
        if red in ("max", "min"):
            assert torch.equal(got[0].cpu(), want) and torch.equal(got[1].cpu(), warg), red
        else:
            scale = oracle.scatter(s.float().abs(), idx, dim, N, red)[0] if red != "mul" else None
            close(got, want, torch.float16, scale)


@pytest.mark.parametrize("full_shape", [True, False])
def test_scatter_out_forms(cuda, full_shape):
    """torch_scatter's out= argument for every reduce (upstream scatter.py: sum/mul accumulate,
    mean = (out + sum) / count, min/max start from out and report arg only where src wins)."""
    import torch_scatter
    g = torch.Generator().manual_seed(31)
    E, F, N = 400, 6, 19
    src = (torch.rand(E, F, generator=g) * 16).round() / 16
    idx1 = torch.randint(0, N - 2, (E,), generator=g)           # rows N-2, N-1 stay empty
    idx = idx1.view(-1, 1).expand(E, F).contiguous() if full_shape else idx1
    if full_shape:
        idx = torch.randint(0, N - 2, (E, F), generator=g)
    out0 = (torch.rand(N, F, generator=g) * 16).round() / 16
    ix_full = idx if full_shape else idx1.view(-1, 1).expand(E, F)
    for red in ("sum", "mul", "mean", "max", "min"):
        s = src * 0.5 + 0.75 if red == "mul" else src
        fn = getattr(torch_scatter, "scatter_" + red)
        o = out0.clone().to(cuda)
        got = fn(s.to(cuda), idx.to(cuda), 0, o)
        fresh, farg = oracle.scatter(s, ix_full.contiguous(), 0, N, red if red != "mean" else "sum")
        if red == "sum":
            want = out0 + fresh
        elif red == "mul":
            want = out0 * fresh
        elif red == "mean":
            cnt = torch.zeros(N, F).scatter_add_(0, ix_full, torch.ones(E, F)).clamp_(min=1)
            want = (out0 + fresh) / cnt
        else:
            want, warg = _ref_minmax_with_out(fresh, farg, out0, E, red)
        if red in ("max", "min"):
            assert got[0].data_ptr() == o.data_ptr()
            assert torch.equal(got[0].cpu(), want) and torch.equal(got[1].cpu(), warg), red
        else:
            assert got.data_ptr() == o.data_ptr()
            assert torch.allclose(got.cpu(), want, rtol=1e-5, atol=1e-5), red


# ---- segment_csr / gather_csr with N-D and sub-range indptr --------------------------------------
def _segment_ref(src, ptr, dim, reduce):
    """Python restatement of torch_scatter.segment_csr for one pointer row per leading index:
    out[..., m, :] reduces src[..., ptr[..., m]:ptr[..., m+1], :]; empty segments give 0 / arg = E."""
    lead = list(src.shape[:dim])
    E, M = src.size(dim), ptr.size(-1) - 1
    s3 = src.reshape(-1, E, int(torch.tensor(src.shape[dim + 1:]).prod()) if src.dim() > dim + 1 else 1)
    p2 = ptr.expand(lead + [M + 1]).reshape(-1, M + 1)
    out = torch.zeros(s3.size(0), M, s3.size(2))
    arg = torch.full((s3.size(0), M, s3.size(2)), E, dtype=torch.int64)
    for b in range(s3.size(0)):
        for m in range(M):
            a, z = int(p2[b, m]), int(p2[b, m + 1])
            if z <= a:
                continue
            seg = s3[b, a:z]
            if reduce == "sum":
                out[b, m] = seg.sum(0)
            elif reduce == "mean":
                out[b, m] = seg.mean(0)
            else:
                v, i = (seg.max(0) if reduce == "max" else seg.min(0))
                out[b, m] = v
                # lowest position among ties, like the sequential upstream loop
                hit = seg == v.unsqueeze(0)
                arg[b, m] = hit.float().argmax(0) + a
    shape = lead + [M] + list(src.shape[dim + 1:])
    return out.view(shape), arg.view(shape)


def test_segment_csr_upstream_docstring_example(cuda):
    """torch_scatter's own docstring: src [10, 6, 64], indptr [1, 4] broadcast over dim 0."""
    import torch_scatter
    g = torch.Generator().manual_seed(1)
    src = torch.randn(10, 6, 64, generator=g)
    indptr = torch.tensor([0, 2, 5, 6]).view(1, -1)
    out = torch_scatter.segment_csr(src.to(cuda), indptr.to(cuda), reduce="sum")
    assert list(out.shape) == [10, 3, 64]
    want = torch.stack([src[:, 0:2].sum(1), src[:, 2:5].sum(1), src[:, 5:6].sum(1)], 1)
    assert torch.allclose(out.cpu(), want, rtol=1e-5, atol=1e-5)
    back = torch_scatter.gather_csr(out, indptr.to(cuda))
    assert list(back.shape) == [10, 6, 64]
    assert torch.equal(back[:, 3].cpu(), out[:, 1].cpu()) and torch.equal(back[:, 5].cpu(), out[:, 2].cpu())


@pytest.mark.parametrize("reduce", ["sum", "mean", "min", "max"])
def test_segment_csr_nd_and_subrange_indptr(cuda, reduce):
    import gno_b200
    g = torch.Generator().manual_seed(12)
    src = (torch.randn(4, 3, 50, 7, generator=g) * 4).round() / 4
    cases = {
        # shared pointers with empty segments, covering only [5, 40) of the dim
        "shared_subrange": torch.tensor([5, 5, 9, 20, 20, 40]).view(1, 1, -1),
        # one pointer row per batch, every batch covers its whole dim (flat-rowptr path)
        "per_batch_full": torch.stack([torch.cat([torch.zeros(1, dtype=torch.long),
                                                  torch.sort(torch.randint(0, 51, (5,), generator=g)).values,
                                                  torch.full((1,), 50)]) for _ in range(12)]).view(4, 3, 7),
        # one pointer row per batch with gaps before the first / after the last pointer
        "per_batch_gaps": torch.stack([torch.sort(torch.randint(0, 51, (6,), generator=g)).values
                                       for _ in range(12)]).view(4, 3, 6),
    }
    for name, ptr in cases.items():
        want, warg = _segment_ref(src, ptr, 2, reduce)
        got = gno_b200.segment_csr(src.to(cuda), ptr.to(cuda), None, reduce, return_arg=True)
        if reduce in ("min", "max"):
            assert torch.equal(got[0].cpu(), want), name
            assert torch.equal(got[1].cpu(), warg), name
        else:
            assert torch.allclose(got.cpu(), want, rtol=1e-5, atol=1e-5), name
    # 1-D indptr over a sub-range of dim 0 (elements outside it are ignored)
    s2 = (torch.randn(60, 5, generator=g) * 4).round() / 4
    p1 = torch.tensor([7, 7, 20, 33, 33, 51])
    want, warg = _segment_ref(s2, p1, 0, reduce)
    got = gno_b200.segment_csr(s2.to(cuda), p1.to(cuda), None, reduce, return_arg=True)
    if reduce in ("min", "max"):
        assert torch.equal(got[0].cpu(), want) and torch.equal(got[1].cpu(), warg)
    else:
        assert torch.allclose(got.cpu(), want, rtol=1e-5, atol=1e-5)


def test_gather_csr_nd(cuda):
    import gno_b200
    g = torch.Generator().manual_seed(13)
    src = torch.randn(3, 4, 6, generator=g)
    shared = torch.tensor([0, 2, 2, 5, 9]).view(1, -1)
    got = gno_b200.gather_csr(src.to(cuda), shared.to(cuda)).cpu()
    rows = torch.repeat_interleave(torch.arange(4), shared[0, 1:] - shared[0, :-1])
    assert torch.equal(got, src[:, rows])
    per = torch.tensor([[0, 1, 4, 4, 6], [0, 0, 3, 5, 6], [0, 2, 2, 2, 6]])
    got = gno_b200.gather_csr(src.to(cuda), per.to(cuda)).cpu()
    for b in range(3):
        rows = torch.repeat_interleave(torch.arange(4), per[b, 1:] - per[b, :-1])
        assert torch.equal(got[b], src[b, rows])
    with pytest.raises(ValueError):
        gno_b200.spmm_csr(torch.tensor([0, 2, 3]).to(cuda), torch.zeros(5, dtype=torch.long).to(cuda), None,
                          torch.ones(4, 2).to(cuda))


# ---- integer values (PyG's bookkeeping calls) ----------------------------------------------------
def test_scatter_integer_values(cuda):
    """TopKPooling / to_dense_batch: scatter_add(batch.new_ones(n), batch, dim=0) on int64
    (graph_benchmark/models/ptg_models.py:165-172 -> GraphUNet), plus every integer reduce against a
    sequential restatement (mean = floor division, arg = lowest position among ties)."""
    import torch_scatter
    g = torch.Generator().manual_seed(41)
    batch = torch.sort(torch.randint(0, 13, (5000,), generator=g)).values
    ones = batch.new_ones(batch.numel())
    got = torch_scatter.scatter_add(ones.to(cuda), batch.to(cuda), dim=0)
    assert got.dtype == torch.int64 and torch.equal(got.cpu(), torch.bincount(batch))
    # large magnitudes: exact where an fp32 round trip would not be
    big = torch.randint(-(1 << 40), 1 << 40, (3000,), generator=g)
    idx = torch.randint(0, 7, (3000,), generator=g)
    want = torch.zeros(7, dtype=torch.int64).index_add_(0, idx, big)
    assert torch.equal(torch_scatter.scatter_add(big.to(cuda), idx.to(cuda), dim=0, dim_size=7).cpu(), want)
    for dtype in (torch.int32, torch.int64):
        E, K, N = 700, 5, 11
        src = torch.randint(-9, 10, (E, K), generator=g).to(dtype)
        idx = torch.randint(0, N - 2, (E,), generator=g)
        for red in ("sum", "mean", "min", "max", "mul"):
            s = src.clamp(-2, 2) if red == "mul" else src
            r = getattr(torch_scatter, "scatter_" + red)(s.to(cuda), idx.to(cuda), 0, None, N)
            want = torch.zeros(N, K, dtype=torch.int64)
            warg = torch.full((N, K), E, dtype=torch.int64)
            for i in range(N):
                rows = torch.nonzero(idx == i).flatten()
                if rows.numel() == 0:
                    want[i] = 1 if red == "mul" else 0
                    continue
                v = s[rows].long()
                if red == "sum":
                    want[i] = v.sum(0)
                elif red == "mean":
                    want[i] = torch.div(v.sum(0), rows.numel(), rounding_mode="floor")
                elif red == "mul":
                    want[i] = v.prod(0)
                else:
                    m = v.max(0).values if red == "max" else v.min(0).values
                    want[i] = m
                    warg[i] = rows[(v == m.unsqueeze(0)).float().argmax(0)]
            if red in ("min", "max"):
                assert r[0].dtype == dtype and torch.equal(r[0].cpu().long(), want), (dtype, red)
                assert torch.equal(r[1].cpu(), warg), (dtype, red)
            else:
                assert r.dtype == dtype and torch.equal(r.cpu().long(), want), (dtype, red)
    # broadcast along the last dim (B > 1, K == 1) and bool / uint8 inputs
    src = torch.randint(0, 5, (4, 300), generator=g)
    idx = torch.randint(0, 9, (300,), generator=g)
    want = torch.zeros(4, 9, dtype=torch.int64).index_add_(1, idx, src)
    assert torch.equal(torch_scatter.scatter_add(src.to(cuda), idx.to(cuda), dim=1, dim_size=9).cpu(), want)
    mask = torch.rand(300, generator=g) > 0.5
    assert torch.equal(torch_scatter.scatter_add(mask.to(cuda), idx.to(cuda), dim=0, dim_size=9).cpu(),
                       torch.zeros(9, dtype=torch.int64).index_add_(0, idx, mask.long()))


# ---- full-shape index on the cached plan (second call onwards): atomic-free path ----------------
@pytest.mark.parametrize("dtype", [torch.float16, torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape,dim,N", [((223, 223), 0, 223), ((223, 223), 1, 56), ((300, 40), 0, 37),
                                         ((5, 40, 7), 1, 9), ((3, 1000), 1, 11), ((2000, 3), 0, 2100)])
def test_scatter_full_shape_planned_path(cuda, dtype, shape, dim, N):
    import gno_b200
    from gno_b200 import ops
    g = torch.Generator().manual_seed(sum(shape) + dim)
    src = ((torch.rand(*shape, generator=g) * 32).round() / 32 - 0.25).to(dtype)
    src.view(-1)[::97] = float("nan")
    src.view(-1)[5::131] = 0.0
    src.view(-1)[6::131] = -0.0
    idx = torch.randint(0, N, shape, generator=g)
    idx.view(-1)[::53] = N + 3          # out of range: dropped
    s_d, i_d = src.to(cuda), idx.to(cuda)
    view = torch.int16 if dtype != torch.float32 else torch.int32
    for red in ("sum", "mean", "mul", "min", "max"):
        s_use = s_d if red != "mul" else (s_d.nan_to_num(1.0) * 0.25 + 0.875)
        s_cpu = s_use.cpu()
        want, warg = oracle.scatter(s_cpu, idx, dim, N, red)      # drops out-of-range destinations too
        first = gno_b200.scatter(s_use, i_d, dim, None, N, red, return_arg=True)
        second = gno_b200.scatter(s_use, i_d, dim, None, N, red, return_arg=True)
        assert any(k[0] == "fs_plan" for k in ops._memo_store), "the second call must run on the cached plan"
        for got in (first, second):
            if red in ("min", "max"):
                assert torch.equal(got[0].cpu().view(view), want.view(view)), red
                assert torch.equal(got[1].cpu(), warg), red
            else:
                a, b = got.float().cpu(), want.float()
                assert torch.equal(torch.isnan(a), torch.isnan(b)), red
                ok = ~torch.isnan(b)
                fin = torch.where(torch.isnan(s_cpu.float()), torch.zeros(()), s_cpu.float().abs())
                scale = oracle.scatter(fin, idx, dim, N, "sum" if red != "mul" else "mul")[0]
                assert not ((a - b).abs()[ok] > TOL[dtype] * torch.maximum(scale, b.abs())[ok] + 1e-30).any(), red
    # out= forms on the planned path
    out0 = ((torch.rand(want_shape(shape, dim, N), generator=g) * 8).round() / 8).to(dtype)
    clean = torch.nan_to_num(s_d, nan=0.5)
    for red in ("sum", "max"):
        o = out0.clone().to(cuda)
        got = gno_b200.scatter(clean, i_d, dim, o, None, red, return_arg=True)
        fresh, farg = oracle.scatter(clean.cpu(), idx, dim, N, red)
        if red == "sum":
            mag = oracle.scatter(clean.cpu().float().abs(), idx, dim, N, "sum")[0]
            close(got, (out0.float() + fresh.float()).to(dtype), dtype, out0.float().abs() + mag)
        else:
            w, wa = _ref_minmax_with_out(fresh, farg, out0, shape[dim], "max")
            assert torch.equal(got[0].cpu(), w) and torch.equal(got[1].cpu(), wa)


def want_shape(shape, dim, N):
    s = list(shape)
    s[dim] = N
    return s


# ---- more edges than a plan indexes (>= 2^31): slices along the scatter dim, combined in order --
@pytest.mark.parametrize("dim", [0, 1])
def test_scatter_sliced_beyond_plan_limit(cuda, dim, monkeypatch):
    """benchmark_scatter_multiply.py:52-58 sweeps a 2.4 G-element 1-D tensor; plans index edges with
    32 bits, so longer inputs go through slices.  The slicing logic is exercised with a small limit
    (test_gpu_fullsize has the real 2^31-element case)."""
    import gno_b200
    from gno_b200 import ops
    monkeypatch.setattr(ops, "_MAX_PLAN_EDGES", 700)
    g = torch.Generator().manual_seed(77)
    E, K, N = 2500, 6, 23
    src = (torch.randn(E, K, generator=g) * 4).round() / 4
    idx = torch.randint(0, N - 2, (E,), generator=g)
    idx[::97] = N + 1
    if dim == 1:
        src = src.t().contiguous()
    for red in ("sum", "mean", "mul", "max", "min"):
        s = src if red != "mul" else torch.where(src.abs() > 2, torch.full_like(src, 2.0), torch.full_like(src, -1.0))
        got = gno_b200.scatter(s.to(cuda), idx.to(cuda), dim, None, N, red, return_arg=True)
        want, warg = oracle.scatter(s, idx, dim, N, red)
        if red in ("max", "min"):
            assert torch.equal(got[0].cpu(), want) and torch.equal(got[1].cpu(), warg), red
        elif red == "mul":
            assert torch.equal(got.cpu(), want)
        else:
            assert torch.allclose(got.cpu(), want, rtol=1e-5, atol=1e-4), red
    ones = torch.ones(E, dtype=torch.int64)
    got = gno_b200.scatter(ones.to(cuda), idx.to(cuda), 0, None, N, "sum")
    assert torch.equal(got.cpu(), torch.bincount(idx[idx < N], minlength=N))


def test_scatter_prepared_launch_cache(cuda):
    """Repeated full-shape calls on one index reuse a prepared launch (what the reference scripts'
    timeit loops do): later calls must track a NEW src tensor, a changed index (version bump) and
    different reduces independently."""
    import gno_b200
    from gno_b200 import ops
    gno_b200.clear_caches()
    g = torch.Generator().manual_seed(5)
    L, N = 150, 40
    idx = torch.randint(0, N, (L, L), generator=g)
    i_d = idx.to(cuda)
    for rep in range(4):
        src = (torch.randn(L, L, generator=g) * 8).round() / 8
        for red in ("sum", "max"):
            got = gno_b200.scatter(src.to(cuda), i_d, 0, None, N, red, return_arg=True)
            want, warg = oracle.scatter(src, idx, 0, N, red)
            if red == "max":
                assert torch.equal(got[0].cpu(), want) and torch.equal(got[1].cpu(), warg), rep
            else:
                assert torch.allclose(got.cpu(), want, rtol=1e-5, atol=1e-4), rep
    assert len(ops._fast_calls) >= 2
    i_d[0, 0] = (i_d[0, 0] + 1) % N           # in-place edit: version bump invalidates plan and launch
    idx[0, 0] = (idx[0, 0] + 1) % N
    src = (torch.randn(L, L, generator=g) * 8).round() / 8
    for _ in range(3):
        got = gno_b200.scatter(src.to(cuda), i_d, 0, None, N, "sum")
        assert torch.allclose(got.cpu(), oracle.scatter(src, idx, 0, N, "sum")[0], rtol=1e-5, atol=1e-4)
