"""Host side of the dst-sorted CSR plan (include/gno_b200.h: gno_plan_build).

A plan is the reusable part of an aggregation: the stable argsort of the
destination index (`perm`), the CSR `rowptr`, the destination row of every
sorted edge (`erow`), and the lists of rows the finish pass touches (rows cut
by a chunk boundary, empty rows).  GNN layers reuse one graph many times, so plans are
cached on the identity of the index tensor — the same role the rowptr /
csr2csc caches play inside torch_sparse.SparseTensor upstream.
"""
import collections
import ctypes

import torch

from . import _lib
from ._lib import check, gno_csr, lib

def default_chunk_len(num_edges):
    """Edges per worker chunk: large enough that chunk-boundary partials are ~1 % of the
    traffic, small enough that small inputs still fill 148 SMs.  Thresholds from
    profiles/c1_chunk.py on a B200 (F=64 fp32, L2 flushed): 250 k edges 27.6 / 33.8 / 44.2 us at
    32 / 64 / 128; 1 M edges 78.8 / 72.7 / 68.6 / 85.0 us at 32 / 64 / 128 / 256; 4 M edges
    226 / 220 / 212 us at 64 / 128 / 256."""
    if num_edges >= 3_000_000:
        return 256
    if num_edges >= 750_000:
        return 128
    if num_edges >= 400_000:
        return 64
    return 32


class _NullCtx:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


_NULL = _NullCtx()


def _on_device(device):
    """`with torch.cuda.device(d)` only when d is not already current: the context manager costs
    several microseconds per call (two driver round trips), which is what the reference scripts'
    timeit loops see on small inputs."""
    if device.index is None or device.index == torch.cuda.current_device():
        return _NULL
    return torch.cuda.device(device)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _workspace(nbytes, device):
    return torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=device)


class CSRPlan:
    """dst-sorted edge list of a 1-D index.  All arrays live on the index's device."""

    __slots__ = ("N", "E", "E_valid", "rowptr", "perm", "erow", "chunk_len", "n_span", "n_empty",
                 "max_len", "n_dropped", "srow", "zrow", "device", "_gidx_cache", "_keepalive")

    def __init__(self):
        self._gidx_cache = collections.OrderedDict()
        self._keepalive = None

    def csr(self, gidx, eid):
        """The C struct for one launch. gidx/eid: int32 tensors or None (identity)."""
        return gno_csr(self.N, self.E_valid, self.rowptr.data_ptr(),
                       self.erow.data_ptr() if self.erow is not None else None,
                       gidx.data_ptr() if gidx is not None else None,
                       eid.data_ptr() if eid is not None else None,
                       self.chunk_len, self.n_span,
                       self.srow.data_ptr() if self.n_span > 0 else None,
                       self.n_empty, self.zrow.data_ptr() if self.n_empty > 0 else None)

    def sorted_ids(self, ids):
        """int32 copy of `ids` (int64 [E]) reordered by the plan: ids[perm]. Cached."""
        if ids.dtype != torch.int64 or ids.dim() != 1 or ids.numel() != self.E:
            raise ValueError(f"ids must be a 1-D int64 tensor of {self.E} elements "
                             f"(got {ids.dtype}, shape {tuple(ids.shape)})")
        if ids.device != self.device:
            raise ValueError("ids must live on the plan's device")
        ids = ids.contiguous()  # a strided view (a row of edges.t()) is copied, not misread
        key = (ids.data_ptr(), ids._version, ids.numel())
        hit = self._gidx_cache.get(key)
        if hit is not None:
            return hit[0]
        out = torch.empty(self.E, dtype=torch.int32, device=self.device)
        check(lib.gno_permute_i64_to_i32(_ptr(ids), _ptr(self.perm), _ptr(out), self.E,
                                         _stream(self.device)))
        self._gidx_cache[key] = (out, ids)
        while len(self._gidx_cache) > 4:
            self._gidx_cache.popitem(last=False)
        return out

    def _finish(self, info):
        """Read the device counters (one sync) and build the span / empty row lists."""
        dev = self.device
        self.n_dropped, self.max_len, self.n_span, self.n_empty = (int(v) for v in info.tolist())
        self.E_valid = self.E - self.n_dropped
        self.srow = torch.empty(self.n_span, dtype=torch.int32, device=dev)
        self.zrow = torch.empty(self.n_empty, dtype=torch.int32, device=dev)
        if self.n_span or self.n_empty:
            nbytes = ctypes.c_size_t()
            check(lib.gno_plan_lists_workspace(self.N, ctypes.byref(nbytes)))
            ws = _workspace(nbytes.value, dev)
            with _on_device(dev):
                check(lib.gno_plan_lists(_ptr(self.rowptr), self.N, self.chunk_len,
                                         _ptr(self.srow) if self.n_span else None,
                                         _ptr(self.zrow) if self.n_empty else None,
                                         _ptr(ws), ws.numel(), _stream(dev)))


def build_plan(index, num_rows, chunk_len=None):
    """Sort `index` (1-D int64, CUDA) by destination; build rowptr, per-edge rows, row lists."""
    if not index.is_cuda:
        raise _lib.GnoError("gno_b200 has no CPU path: index must be a CUDA tensor")
    if index.dim() != 1 or index.dtype != torch.int64:
        raise ValueError("plan index must be a 1-D int64 tensor")
    index = index.contiguous()
    dev = index.device
    E, N = index.numel(), int(num_rows)
    p = CSRPlan()
    p.N, p.E, p.device = N, E, dev
    p.chunk_len = int(chunk_len) if chunk_len else default_chunk_len(E)
    p.rowptr = torch.empty(N + 1, dtype=torch.int64, device=dev)
    p.perm = torch.empty(max(E, 1), dtype=torch.int32, device=dev)[:E]
    p.erow = torch.empty(max(E, 1), dtype=torch.int32, device=dev)[:E]
    info = torch.empty(4, dtype=torch.int64, device=dev)
    nbytes = ctypes.c_size_t()
    check(lib.gno_plan_workspace(E, N, ctypes.byref(nbytes)))
    ws = _workspace(nbytes.value, dev)
    with _on_device(dev):
        check(lib.gno_plan_build(_ptr(index), E, N, p.chunk_len, _ptr(p.rowptr), _ptr(p.perm),
                                 _ptr(p.erow), _ptr(info), _ptr(ws), ws.numel(), _stream(dev)))
    p._finish(info)
    return p


def plan_from_rowptr(rowptr, nnz, chunk_len=None):
    """Plan over a caller-supplied CSR rowptr (segment_csr / SparseTensor inputs)."""
    if not rowptr.is_cuda:
        raise _lib.GnoError("gno_b200 has no CPU path: rowptr must be a CUDA tensor")
    rowptr = rowptr.contiguous().to(torch.int64)
    dev = rowptr.device
    N, E = rowptr.numel() - 1, int(nnz)
    p = CSRPlan()
    p.N, p.E, p.device = N, E, dev
    p.chunk_len = int(chunk_len) if chunk_len else default_chunk_len(E)
    p.rowptr, p.perm = rowptr, None
    p.erow = torch.empty(max(E, 1), dtype=torch.int32, device=dev)[:E]
    info = torch.empty(4, dtype=torch.int64, device=dev)
    with _on_device(dev):
        check(lib.gno_plan_from_rowptr(_ptr(rowptr), N, E, p.chunk_len, _ptr(p.erow), _ptr(info),
                                       _stream(dev)))
    p._finish(info)
    return p


class PlanCache:
    """LRU of plans keyed on the identity + version of the index tensor."""

    def __init__(self, capacity=8):
        self.capacity = capacity
        self._d = collections.OrderedDict()
        self.hits = 0
        self.misses = 0

    def get(self, index, num_rows, chunk_len=None):
        if index.dtype != torch.int64 or index.dim() != 1:
            raise ValueError("plan index must be a 1-D int64 tensor")
        index = index.contiguous()  # two views of one buffer (t[:E], t[::2]) must not share a plan
        key = (index.data_ptr(), index._version, index.numel(), int(num_rows), chunk_len,
               str(index.device))
        p = self._d.get(key)
        if p is not None:
            self._d.move_to_end(key)
            self.hits += 1
            return p
        self.misses += 1
        p = build_plan(index, num_rows, chunk_len)
        if p.n_dropped:
            import warnings
            warnings.warn(f"gno_b200: {p.n_dropped} index entries outside [0, {int(num_rows)}) were dropped "
                          "(upstream torch_scatter leaves this undefined on CUDA)", RuntimeWarning, stacklevel=3)
        p._keepalive = index  # the key is a raw pointer: keep the tensor alive
        self._d[key] = p
        while len(self._d) > self.capacity:
            self._d.popitem(last=False)
        return p

    def clear(self):
        self._d.clear()


plan_cache = PlanCache()
