"""CPU tests of the oracle itself: (1) against the golden vectors produced by the reference's
own op functions (tests/golden/make_golden.py), (2) against independent torch / scipy
formulations of the upstream semantics (scatter_reduce_, scatter_add_, index_add_,
sparse.mm, Tensor.coalesce, sort), (3) hypothesis property tests for the arg definition."""
import os

import numpy as np
import pytest
import torch
from hypothesis import given, settings, strategies as st

import oracle

GOLD = os.path.join(os.path.dirname(__file__), "golden", "native_ops.npz")


def t(z, k):
    return torch.from_numpy(z[k])


def test_golden_sort():
    z = np.load(GOLD)
    for tag in ("sort1d", "sort2d_d0", "sort2d_d1"):
        v, i = oracle.sort(t(z, tag + "_in"), int(z[tag + "_dim"]))
        assert torch.equal(i, t(z, tag + "_idx")) and torch.equal(v, t(z, tag + "_val"))


def test_golden_index_ops():
    z = np.load(GOLD)
    got = oracle.index_add(t(z, "iadd_in"), 1, t(z, "iadd_index"), t(z, "iadd_src"))
    # fp16 index_add_ on CPU rounds per add; the oracle accumulates in fp32 and rounds once
    assert torch.allclose(got.float(), t(z, "iadd_out").float(), rtol=1e-2, atol=1e-2)
    a, b, idx = t(z, "far_in"), t(z, "far_other"), t(z, "far_index")
    for dim in (0, 1):
        out = oracle.index_add(a, dim, idx, b)
        res = out.index_select(dim, idx).sum(dim)
        assert torch.allclose(res, t(z, f"far_out_d{dim}"), rtol=1e-5, atol=1e-4)


def test_golden_spmm_coalesce_scatter():
    z = np.load(GOLD)
    got = oracle.spmm(t(z, "smm_index"), t(z, "smm_value"), int(z["smm_m"]), int(z["smm_n"]), t(z, "smm_B"))
    assert torch.allclose(got, t(z, "smm_out"), rtol=1e-5, atol=1e-5)
    gi, gv = oracle.coalesce(t(z, "coal_index"), t(z, "coal_value"), int(z["coal_m"]), int(z["coal_n"]))
    assert torch.equal(gi, t(z, "coal_out_index"))
    assert torch.allclose(gv, t(z, "coal_out_value"), rtol=1e-6)
    got, _ = oracle.scatter(t(z, "sadd_src"), t(z, "sadd_idx"), 0, t(z, "sadd_src").shape[0], "sum")
    assert torch.allclose(got.float(), t(z, "sadd_out").float(), rtol=1e-2, atol=1e-2)
    got, _ = oracle.scatter(t(z, "smul_src"), t(z, "smul_idx"), -1, t(z, "smul_src").shape[-1], "mul")
    assert torch.allclose(got, t(z, "smul_out"), rtol=1e-5)


@pytest.mark.parametrize("reduce,native", [("sum", "sum"), ("mean", "mean"), ("mul", "prod"), ("max", "amax"), ("min", "amin")])
@pytest.mark.parametrize("dim", [0, 1])
def test_scatter_vs_scatter_reduce(reduce, native, dim):
    g = torch.Generator().manual_seed(7)
    src = torch.rand(60, 50, generator=g) + 0.5
    idx = torch.randint(0, 20, (60, 50), generator=g)
    shape = [60, 50]
    shape[dim] = 25  # rows 20..24 stay empty
    # untouched rows: 0 for sum/mean/min/max (zeros / masked_fill), 1 for mul (torch_scatter starts from ones)
    base = torch.ones(shape) if reduce == "mul" else torch.zeros(shape)
    want = base.scatter_reduce_(dim, idx, src, native, include_self=False)
    got, arg = oracle.scatter(src, idx, dim, 25, reduce)
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-6)
    if arg is not None:
        # definition: arg = lowest position along dim whose value equals the winner; sentinel = size(dim)
        picked = torch.gather(torch.cat([src, torch.zeros_like(src.narrow(dim, 0, 1))], dim), dim, arg)
        assert torch.equal(picked, got)
        assert int(arg.max()) == src.size(dim)


def test_scatter_1d_vs_index_add_and_expanded():
    g = torch.Generator().manual_seed(8)
    src = torch.randn(500, 13, generator=g)
    idx = torch.randint(0, 40, (500,), generator=g)
    got, _ = oracle.scatter(src, idx, 0, 40, "sum")
    assert torch.allclose(got, torch.zeros(40, 13).index_add_(0, idx, src), rtol=1e-5, atol=1e-5)
    full, _ = oracle.scatter(src, idx.view(-1, 1).expand(-1, 13).contiguous(), 0, 40, "sum")
    assert torch.equal(got, full)


@settings(max_examples=60, deadline=None)
@given(st.integers(1, 40), st.integers(1, 6), st.integers(1, 8), st.integers(0, 2 ** 31 - 1), st.sampled_from(["max", "min"]))
def test_arg_definition(E, K, N, seed, reduce):
    """arg[i,k] = min{e : index[e]=i and src[e,k]=out[i,k]}; empty rows → (0, E)."""
    g = torch.Generator().manual_seed(seed)
    src = torch.randint(-3, 4, (E, K), generator=g).float()
    idx = torch.randint(0, N, (E,), generator=g)
    out, arg = oracle.scatter(src, idx, 0, N, reduce)
    for i in range(N):
        rows = (idx == i).nonzero().flatten()
        for k in range(K):
            if rows.numel() == 0:
                assert out[i, k] == 0 and arg[i, k] == E
            else:
                vals = src[rows, k]
                best = vals.max() if reduce == "max" else vals.min()
                assert out[i, k] == best
                assert arg[i, k] == rows[(vals == best).nonzero()[0, 0]]


def test_gather_scatter_equals_unfused():
    g = torch.Generator().manual_seed(9)
    x = torch.randn(70, 9, generator=g)
    s, d = torch.randint(0, 70, (400,), generator=g), torch.randint(0, 30, (400,), generator=g)
    for red in ("sum", "mean", "max", "min", "mul"):
        a, aa = oracle.gather_scatter(x, s, d, 30, red)
        b, ba = oracle.scatter(x.index_select(0, s), d, 0, 30, red)
        assert torch.equal(a, b)
        assert (aa is None and ba is None) or torch.equal(aa, ba)


def test_spmm_csr_vs_torch_and_scipy():
    import scipy.sparse as sp
    g = torch.Generator().manual_seed(10)
    M, N, F, nnz = 50, 40, 7, 600
    row = torch.randint(0, M, (nnz,), generator=g).sort().values
    col = torch.randint(0, N, (nnz,), generator=g)
    val = torch.rand(nnz, generator=g)
    rowptr = torch.zeros(M + 1, dtype=torch.int64)
    rowptr[1:] = torch.cumsum(torch.bincount(row, minlength=M), 0)
    mat = torch.randn(N, F, generator=g)
    got, _ = oracle.spmm_csr(rowptr, col, val, mat, "sum")
    csr = torch.sparse_csr_tensor(rowptr, col, val, (M, N))
    assert torch.allclose(got, torch.sparse.mm(csr, mat), rtol=1e-5, atol=1e-5)
    s = sp.csr_matrix((val.numpy(), col.numpy(), rowptr.numpy()), shape=(M, N))
    assert np.allclose(got.numpy(), s @ mat.numpy(), rtol=1e-5, atol=1e-5)
    coo = oracle.spmm(torch.stack([row, col]), val, M, N, mat)
    assert torch.allclose(coo, got, rtol=1e-5, atol=1e-5)


def test_coalesce_transpose_vs_native_and_scipy():
    import scipy.sparse as sp
    g = torch.Generator().manual_seed(11)
    m, n, nnz = 30, 45, 900
    index = torch.stack([torch.randint(0, m, (nnz,), generator=g), torch.randint(0, n, (nnz,), generator=g)])
    value = torch.rand(nnz, generator=g)
    gi, gv = oracle.coalesce(index, value, m, n)
    nat = torch.sparse_coo_tensor(index, value, (m, n)).coalesce()
    assert torch.equal(gi, nat.indices()) and torch.allclose(gv, nat.values(), rtol=1e-6)
    ti, tv = oracle.transpose(index, value, m, n)
    s = sp.coo_matrix((value.numpy(), (index[0].numpy(), index[1].numpy())), shape=(m, n)).T.tocsr()
    s.sum_duplicates()
    s.sort_indices()
    c = s.tocoo()
    assert np.array_equal(ti[0].numpy(), c.row) and np.array_equal(ti[1].numpy(), c.col)
    assert np.allclose(tv.numpy(), c.data, rtol=1e-6)


@pytest.mark.parametrize("descending", [False, True])
def test_sort_vs_torch(descending):
    g = torch.Generator().manual_seed(12)
    x = (torch.randn(40, 50, generator=g) * 2).round() / 2
    x.view(-1)[::7] = float("nan")
    x.view(-1)[3::11] = -0.0
    x.view(-1)[5::13] = 0.0
    for dim in (0, 1):
        v, i = oracle.sort(x, dim, descending)
        tv, ti = torch.sort(x, dim=dim, descending=descending, stable=True)
        assert torch.equal(i, ti)
        assert torch.equal(v.view(torch.int32), tv.view(torch.int32))


def test_csr_from_index():
    idx = torch.tensor([3, 0, 3, 1, 0, 3])
    rowptr, perm = oracle.csr_from_index(idx, 5)
    assert rowptr.tolist() == [0, 2, 3, 3, 6, 6]
    assert perm.tolist() == [1, 4, 3, 0, 2, 5]


def test_graph_construction_restatement():
    """oracle.to_undirected / coalesce_edges / remove_self_loops against a key-based torch formulation."""
    g = torch.Generator().manual_seed(3)
    n = 40
    ei = torch.stack([torch.randint(0, n, (500,), generator=g), torch.randint(0, n, (500,), generator=g)])
    nl = oracle.remove_self_loops(ei)
    assert torch.equal(nl, ei[:, ei[0] != ei[1]])
    both = torch.cat([nl, nl.flip(0)], dim=1)
    key = torch.unique(both[0] * n + both[1])
    und = oracle.to_undirected(nl, n)
    assert torch.equal(und[0] * n + und[1], key)
    co = oracle.coalesce_edges(ei, n)
    assert torch.equal(co[0] * n + co[1], torch.unique(ei[0] * n + ei[1]))


def test_oracle_matches_upstream_published_examples():
    """The worked examples of the pytorch_scatter / pytorch_sparse READMEs (tests/golden/upstream_published.py)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "upstream_published", os.path.join(os.path.dirname(__file__), "golden", "upstream_published.py"))
    up = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(up)
    v = up.SCATTER_MAX
    out, arg = oracle.scatter(v["src"], v["index"], -1, None, "max")
    assert torch.equal(out, v["out"]) and torch.equal(arg, v["arg"])
    v = up.COALESCE
    i, val = oracle.coalesce(v["index"], v["value"], v["m"], v["n"])
    assert torch.equal(i, v["out_index"]) and torch.equal(val, v["out_value"])
    v = up.TRANSPOSE
    i, val = oracle.transpose(v["index"], v["value"], v["m"], v["n"])
    assert torch.equal(i, v["out_index"]) and torch.equal(val, v["out_value"])
    v = up.SPMM
    assert torch.equal(oracle.spmm(v["index"], v["value"], v["m"], v["n"], v["matrix"]), v["out"])


def _load_golden(name):
    import importlib.util
    spec = importlib.util.spec_from_file_location(name, os.path.join(os.path.dirname(__file__), "golden", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _upstream_broadcast(index, src):
    """pytorch_scatter utils.broadcast for an index with fewer dims than src (trailing dims)."""
    if 1 < index.dim() < src.dim():
        index = index.reshape(tuple(index.shape) + (1,) * (src.dim() - index.dim())).expand(src.shape).contiguous()
    return index


def test_oracle_matches_upstream_testsuite_tables():
    """The `tests = [...]` tables of upstream's own test suites (tests/golden/upstream_testsuite.py):
    scatter sum / mul / mean / min / max + arg for every index shape class, and the segment tables
    (segment_coo == scatter over the sorted index; segment_csr == the same after expanding indptr)."""
    ut = _load_golden("upstream_testsuite")
    for v in ut.SCATTER:
        for red in ("sum", "mul", "mean", "min", "max"):
            out, arg = oracle.scatter(v["src"], _upstream_broadcast(v["index"], v["src"]), v["dim"], None, red)
            assert torch.equal(out, v[red]), (red, out)
            if red in ("min", "max"):
                assert torch.equal(arg, v["arg_" + red]), (red, arg)
    for v in ut.SEGMENT:
        dim = v["index"].dim() - 1
        n_seg = v["indptr"].size(-1) - 1
        # segment ids from the row pointers must reproduce the table's own index
        ptr = v["indptr"].view(-1, n_seg + 1)
        ids = torch.stack([torch.repeat_interleave(torch.arange(n_seg), p[1:] - p[:-1]) for p in ptr])
        assert torch.equal(ids.view(v["index"].shape), v["index"])
        for red in ("sum", "mean", "min", "max"):
            out, arg = oracle.scatter(v["src"], v["index"], dim, n_seg, red)
            assert torch.equal(out, v[red]), (red, out)
            if red in ("min", "max"):
                assert torch.equal(arg, v["arg_" + red]), (red, arg)
    for v in ut.GATHER:     # gather is pure indexing: the table against torch's own gather / index_select
        idx, src = v["index"], v["src"]
        if idx.dim() == 1:
            got = src.index_select(0, idx)
        else:
            got = src.gather(idx.dim() - 1, idx)
        assert torch.equal(got, v["expected"])
