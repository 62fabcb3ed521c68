"""ctypes/numpy wrapper around oracle.c — the checker used by tests/,
__graft_entry__.smoke() and bench.py's CPU-baseline legs.  NOT product code.

PARITY UNPINNED for the torch_scatter / torch_sparse semantics (the upstream
packages are absent; see oracle.c) beyond the four worked examples their
READMEs publish (tests/golden/upstream_published.py); pinned against
tests/golden/*.npz for the native torch ops the reference scripts call.

All functions take and return torch CPU tensors.  Half / bfloat16 inputs are
widened to float32, reduced in float32 and rounded once on return.
"""
import ctypes
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

RED = {"sum": 0, "add": 0, "mean": 1, "mul": 2, "min": 3, "max": 4}

_FINITE = {
    torch.float32: (-3.402823466e+38, 3.402823466e+38),
    torch.float16: (-65504.0, 65504.0),
    torch.bfloat16: (-3.3895313892515355e+38, 3.3895313892515355e+38),
}


def build(force=False):
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B", "liboracle.so"], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.oracle_coalesce.restype = ctypes.c_int64
    return _lib


def _f32(t):
    return np.ascontiguousarray(t.detach().cpu().to(torch.float32).numpy())


def _i64(t):
    return np.ascontiguousarray(t.detach().cpu().to(torch.int64).numpy())


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else ctypes.c_void_p(0)


def _c64(v):
    return ctypes.c_int64(int(v))


def _cf(v):
    return ctypes.c_float(float(v))


def _bek(shape, dim):
    B = int(np.prod(shape[:dim], dtype=np.int64)) if dim > 0 else 1
    E = int(shape[dim])
    K = int(np.prod(shape[dim + 1:], dtype=np.int64)) if dim + 1 < len(shape) else 1
    return B, E, K


def scatter(src, index, dim=-1, dim_size=None, reduce="sum"):
    """torch_scatter.scatter_* restated (oracle.c: oracle_scatter). Returns (out, arg);
    arg is None unless reduce is min/max."""
    if dim < 0:
        dim += src.dim()
    is_1d = index.dim() == 1 and src.dim() > 1  # 1-D index broadcast along `dim`
    if not is_1d and tuple(index.shape) != tuple(src.shape):
        src = src[tuple(slice(0, s) for s in index.shape)]
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() else 0
    N = int(dim_size)
    B, E, K = _bek(list(src.shape), dim)
    s, ix = _f32(src), _i64(index)
    out_shape = list(src.shape[:dim]) + [N] + list(src.shape[dim + 1:])
    out = np.empty(out_shape, dtype=np.float32)
    want_arg = reduce in ("min", "max")
    arg = np.empty(out_shape, dtype=np.int64) if want_arg else None
    lo, hi = _FINITE[src.dtype]
    lib().oracle_scatter(_p(s), _p(ix), ctypes.c_int(1 if is_1d else 0), _c64(B), _c64(E), _c64(K),
                         _c64(N), ctypes.c_int(RED[reduce]), _cf(lo), _cf(hi), _p(out), _p(arg))
    o = torch.from_numpy(out).to(src.dtype)
    return o, (torch.from_numpy(arg) if want_arg else None)


def gather_scatter(x, src_ids, dst_ids, dim_size, reduce="sum"):
    """scatter(x.index_select(0, src_ids), dst_ids, 0, dim_size, reduce) (oracle_gather_scatter)."""
    N, F = int(dim_size), int(x.shape[1])
    xs, si, di = _f32(x), _i64(src_ids), _i64(dst_ids)
    out = np.empty((N, F), dtype=np.float32)
    want_arg = reduce in ("min", "max")
    arg = np.empty((N, F), dtype=np.int64) if want_arg else None
    lo, hi = _FINITE[x.dtype]
    lib().oracle_gather_scatter(_p(xs), _c64(F), _p(si), _p(di), _c64(si.size), _c64(N),
                                ctypes.c_int(RED[reduce]), _cf(lo), _cf(hi), _p(out), _p(arg))
    return torch.from_numpy(out).to(x.dtype), (torch.from_numpy(arg) if want_arg else None)


def index_add(input, dim, index, source):
    """torch.index_add(input, dim, index, source) (oracle_index_add)."""
    if dim < 0:
        dim += input.dim()
    B, E, K = _bek(list(source.shape), dim)
    N = int(input.shape[dim])
    io = _f32(input).copy()
    s, ix = _f32(source), _i64(index)
    lib().oracle_index_add(_p(io), _p(s), _p(ix), _c64(B), _c64(E), _c64(K), _c64(N))
    return torch.from_numpy(io).to(input.dtype)


def spmm_csr(rowptr, col, value, mat, reduce="sum"):
    """CSR SpMM with reductions (oracle_spmm_csr). Returns (out, arg)."""
    M, F = int(rowptr.numel() - 1), int(mat.shape[1])
    rp, c, m = _i64(rowptr), _i64(col), _f32(mat)
    v = _f32(value) if value is not None else None
    out = np.empty((M, F), dtype=np.float32)
    want_arg = reduce in ("min", "max")
    arg = np.empty((M, F), dtype=np.int64) if want_arg else None
    lo, hi = _FINITE[mat.dtype]
    lib().oracle_spmm_csr(_p(rp), _p(c), _p(v), _p(m), _c64(M), _c64(F), ctypes.c_int(RED[reduce]),
                          _cf(lo), _cf(hi), _p(out), _p(arg))
    return torch.from_numpy(out).to(mat.dtype), (torch.from_numpy(arg) if want_arg else None)


def spmm(index, value, m, n, matrix):
    """torch_sparse.spmm(index, value, m, n, matrix) = index_select * value → scatter_add on row
    (torch-sparse 0.6.12 spmm.py), restated through the sequential COO loop."""
    row, col = index[0], index[1]
    msgs = matrix.to(torch.float32).index_select(0, col) * value.to(torch.float32).unsqueeze(-1)
    out, _ = scatter(msgs, row, 0, m, "sum")
    return out.to(matrix.dtype)


def coalesce(index, value, m, n, op="add"):
    """torch_sparse.coalesce (oracle_coalesce). Returns (index[2, nnz'], value or None)."""
    E = int(index.shape[1])
    row, col = _i64(index[0]), _i64(index[1])
    K = 0
    v = None
    if value is not None:
        v = _f32(value.reshape(E, -1))
        K = v.shape[1]
    out_row = np.empty(max(E, 1), dtype=np.int64)
    out_col = np.empty(max(E, 1), dtype=np.int64)
    out_val = np.empty((max(E, 1), max(K, 1)), dtype=np.float32) if v is not None else None
    cnt = lib().oracle_coalesce(_p(row), _p(col), _p(v), _c64(K), _c64(E), _c64(m), _c64(n),
                                ctypes.c_int(RED[op]), _p(out_row), _p(out_col), _p(out_val))
    idx = torch.from_numpy(np.stack([out_row[:cnt], out_col[:cnt]]))
    if value is None:
        return idx, None
    ov = torch.from_numpy(out_val[:cnt]).to(value.dtype).reshape([cnt] + list(value.shape[1:]))
    return idx, ov


def transpose(index, value, m, n):
    """torch_sparse.transpose(index, value, m, n, coalesced=True)."""
    return coalesce(torch.stack([index[1], index[0]]), value, n, m, "add")


def sort(input, dim=-1, descending=False):
    """torch.sort(input, dim, descending, stable=True) (oracle_sort_f32)."""
    if dim < 0:
        dim += input.dim()
    outer, length, inner = _bek(list(input.shape), dim)
    x = _f32(input)
    vals = np.empty(x.shape, dtype=np.float32)
    idx = np.empty(x.shape, dtype=np.int64)
    lib().oracle_sort_f32(_p(x), _c64(outer), _c64(length), _c64(inner),
                          ctypes.c_int(1 if descending else 0), _p(vals), _p(idx))
    return torch.from_numpy(vals), torch.from_numpy(idx)


def argsort_stable(keys):
    k = np.ascontiguousarray(keys.detach().cpu().numpy().astype(np.uint64))
    perm = np.empty(k.shape, dtype=np.int64)
    lib().oracle_argsort_u64(_p(k), _c64(k.size), _p(perm))
    return torch.from_numpy(perm)


def csr_from_index(index, num_rows):
    """Stable dst-sort → (rowptr, perm): the plan the CUDA path builds."""
    idx = index.detach().cpu().to(torch.int64)
    perm = argsort_stable(idx)
    counts = torch.bincount(idx, minlength=num_rows)[:num_rows]
    rowptr = torch.zeros(num_rows + 1, dtype=torch.int64)
    rowptr[1:] = torch.cumsum(counts, 0)
    return rowptr, perm


# ---- graph construction (torch_geometric.utils restated; fakeDatasets.py:238-259) ---------------
def remove_self_loops(edge_index):
    keep = [e for e in range(edge_index.shape[1]) if int(edge_index[0, e]) != int(edge_index[1, e])]
    return edge_index[:, keep]


def to_undirected(edge_index, num_nodes):
    """Both directions of every edge, sorted by (row, col), duplicates removed."""
    pairs = set()
    for e in range(edge_index.shape[1]):
        a, b = int(edge_index[0, e]), int(edge_index[1, e])
        pairs.add((a, b))
        pairs.add((b, a))
    out = sorted(pairs)
    return torch.tensor(out, dtype=torch.int64).t().reshape(2, -1)


def coalesce_edges(edge_index, num_nodes):
    out = sorted({(int(edge_index[0, e]), int(edge_index[1, e])) for e in range(edge_index.shape[1])})
    return torch.tensor(out, dtype=torch.int64).t().reshape(2, -1)
