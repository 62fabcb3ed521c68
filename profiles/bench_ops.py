"""Per-op device timings for every BASELINE.json config (C1..C4, C5-per-GPU shape) with the
HBM-roofline fraction of each.  Not the driver's bench (that is /bench.py): this is the
wider table DESIGN.md / profiles/ quote.

    python profiles/bench_ops.py [--only c1,c2,...] [--iters 10] > gpurun_out/ops.jsonl

One JSON line per measurement (appended as they finish, so a timeout keeps partial results).
Timing: CUDA events on the current stream, 3 warm-ups, median of `iters`; every working set
is far larger than the 126 MB L2 except C1, where a 256 MB scratch write flushes L2 between
iterations.
"""
import argparse
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "gnn-ops-benchmark_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402

import bench as B  # noqa: E402
import gno_b200  # noqa: E402
from gno_b200 import plan as planmod  # noqa: E402

PEAK, PEAK_SRC = B.peaks()
DEV = torch.device("cuda:0")
_flush_buf = None


def flush_l2():
    global _flush_buf
    if _flush_buf is None:
        _flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    _flush_buf.fill_(1)


def timeit(fn, iters, flush=False):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush:
            flush_l2()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)


RESULTS = []  # every emitted line, for callers that run this in-process (bench.py per_config)


try:
    with open(os.path.join(ROOT, "profiles", "traffic.json")) as _f:
        _TRAFFIC = json.load(_f).get("detail", {})
except OSError:
    _TRAFFIC = {}


def emit(name, ms, units, unit_name, abytes, dram_key=None, **extra):
    """dram_key: entry of profiles/traffic.json (ncu dram__bytes of the same launch) — frac_dram is
    the share of the measured copy peak the kernel's REAL DRAM traffic reaches in the time measured
    here; frac_of_peak is on algorithmic bytes and exceeds 1 where L2 serves repeated gathers."""
    gbs = abytes / (ms * 1e-3) / 1e9
    line = {"op": name, "ms": round(ms, 4), unit_name + "_per_s": units / (ms * 1e-3),
            "algorithmic_GB": round(abytes / 1e9, 3), "achieved_GBps": round(gbs, 1),
            "frac_of_peak": round(gbs / PEAK, 4), "peak_GBps": PEAK, "peak_source": PEAK_SRC}
    t = _TRAFFIC.get(dram_key) if dram_key else None
    if t:
        line["dram_GB_ncu"] = round(t["dram_bytes"] / 1e9, 3)
        line["frac_dram"] = round(t["dram_bytes"] / (ms * 1e-3) / 1e9 / PEAK, 4)
        line["l2_hit_pct_ncu"] = t["l2_hit_pct"].get("segreduce_staged_kernel")
    line.update(extra)
    RESULTS.append(line)
    print(json.dumps(line), flush=True)


def agg_bytes(N, E, F, es, arg=False, weight=False):
    """E*(F*s + 4 [+s weight]) + N*(F*s [+8F arg] + 4)."""
    return E * (F * es + 4 + (es if weight else 0)) + N * (F * es + (8 * F if arg else 0) + 4)


def run_c1(iters):
    E, N, F = 1_000_000, 100_000, 64
    g = torch.Generator(device=DEV).manual_seed(42)
    src = torch.rand(E, F, device=DEV, generator=g)
    idx = torch.randint(0, N, (E,), device=DEV, generator=g)
    gno_b200.clear_caches()
    ms_cold = timeit(lambda: (gno_b200.clear_caches(), gno_b200.scatter(src, idx, 0, None, N, "sum")), iters, True)
    ms = timeit(lambda: gno_b200.scatter(src, idx, 0, None, N, "sum"), iters, True)
    ab = E * (F * 4 + 8) + N * F * 4  # int64 index as given
    ix = idx.view(-1, 1).expand(E, F)
    nat = timeit(lambda: torch.zeros(N, F, device=DEV).scatter_add_(0, ix, src), iters, True)
    nat_max = timeit(lambda: torch.zeros(N, F, device=DEV).scatter_reduce_(0, ix, src, "amax", include_self=False), iters, True)
    emit("C1 scatter_sum fp32 [1M,64]->100k (plan cached)", ms, E, "edges", ab, dram_key="c1_sum", l2="flushed",
         native_torch_ms={"zeros+scatter_add_ (what torch_scatter.scatter_sum runs)": round(nat, 4),
                          "zeros+scatter_reduce_(amax)": round(nat_max, 4)})
    emit("C1 scatter_sum fp32 [1M,64]->100k (cold: plan build included)", ms_cold, E, "edges", ab, l2="flushed")
    for red in ("mean", "max", "min", "mul"):
        ms = timeit(lambda: gno_b200.scatter(src, idx, 0, None, N, red, return_arg=True), iters, True)
        emit(f"C1 scatter_{red} fp32 (plan cached)", ms, E, "edges",
             ab + (N * F * 8 if red in ("max", "min") else 0), l2="flushed")
    # the script-faithful call: fp16 [L,L] + full-shape int64 index, dims 0/1 (benchmark_scatter_add.py:60-84)
    L = 6708
    s16 = torch.rand(L, L, device=DEV, generator=g).half()
    ifull = torch.randint(0, L, (L, L), device=DEV, generator=g)
    for red in ("sum", "max", "mean"):
        for dim in (0, 1):
            ms = timeit(lambda: gno_b200.scatter(s16, ifull, dim, None, L, red, return_arg=True), iters)
            natred = {"sum": "sum", "max": "amax", "mean": "mean"}[red]
            nat = timeit(lambda: torch.zeros(L, L, device=DEV, dtype=torch.float16).scatter_reduce_(
                dim, ifull, s16, natred, include_self=False), iters)
            emit(f"script-shape scatter_{red} fp16 ({L},{L}) full-shape index dim{dim}", ms, L * L, "elems",
                 L * L * (2 + 8) + L * L * 2, a100_ref_ms={"sum": (6.688, 3.678), "max": (14.531, 6.704),
                                                            "mean": (13.602, 7.616)}[red][dim],
                 native_torch_ms={f"zeros+scatter_reduce_({natred})": round(nat, 4)})


def graph(name):
    n, e, F, dtype, ex, off = B.WORKLOADS[name]
    src, dst = B.make_graph(n, n, e, ex, off, DEV, 42)
    g = torch.Generator(device=DEV).manual_seed(7)
    x = torch.randn(n, F, device=DEV, generator=g)
    return n, e, F, src, dst, x


def run_c2(iters):
    n, e, F, src, dst, x = graph("products")
    t = timeit(lambda: planmod.build_plan(dst, n), 5)
    nat = timeit(lambda: torch.sort(dst, stable=True), 3)
    emit("C2 plan build (dst radix sort + rowptr + lists), 61.9M edges", t, e, "edges", e * (8 + 4 + 4) + n * 8,
         native_torch_ms={"torch.sort(dst, stable=True) alone (CUB, int64 keys)": round(nat, 3)})
    plan = planmod.build_plan(dst, n)
    gidx = plan.sorted_ids(src)
    out = torch.empty(n, F, device=DEV)
    for red in ("sum", "mean", "max"):
        arg = red == "max"
        ms = timeit(lambda: gno_b200.segment_reduce(plan, x, red, gidx=gidx, eid=plan.perm, want_arg=arg,
                                                    out=None if arg else out), iters)
        extra = {}
        if red == "sum":
            def native():
                o = torch.zeros(n, F, device=DEV)
                o.index_add_(0, dst, x.index_select(0, src))   # materialises the [E, F] messages (24.7 GB)
            extra["native_torch_ms"] = {"index_select + zeros + index_add_ (unfused)": round(timeit(native, 3), 3)}
        emit(f"C2 products gather->scatter_{red} fp32 F=100", ms, e, "edges", agg_bytes(n, e, F, 4, arg),
             dram_key="products_sum" if red == "sum" else None, **extra)


def run_c3(iters):
    n, e, F, src, dst, x = graph("reddit")
    plan = planmod.build_plan(dst, n)
    gidx = plan.sorted_ids(src)
    del src, dst
    for dtype in (torch.float32, torch.bfloat16):
        xs = x.to(dtype)
        es = xs.element_size()
        for red in ("max", "min", "mean", "sum"):
            arg = red in ("max", "min")
            ms = timeit(lambda: gno_b200.segment_reduce(plan, xs, red, gidx=gidx, eid=plan.perm, want_arg=arg), iters)
            emit(f"C3 reddit gather->scatter_{red}{'+arg' if arg else ''} {str(dtype)[6:]} F=602", ms, e, "edges",
                 agg_bytes(n, e, F, es, arg),
                 dram_key=("reddit" if dtype == torch.float32 else "reddit_bf16") + "_" + red)
    # un-fused scatter_max(src[E', 602]) at E' = E/8 (full src would be 276 GB)
    e8 = e // 8
    g = torch.Generator(device=DEV).manual_seed(3)
    idx = torch.randint(0, n, (e8,), device=DEV, generator=g)
    s = torch.randn(e8, F, device=DEV, generator=g)
    ms = timeit(lambda: gno_b200.scatter(s, idx, 0, None, n, "max", return_arg=True), iters)
    emit("C3 unfused scatter_max+arg fp32 src[E/8,602] (plan cached)", ms, e8, "edges", agg_bytes(n, e8, F, 4, True))


def run_c4(iters):
    n, e, _, src, dst, _ = graph("reddit")
    F = 256
    g = torch.Generator(device=DEV).manual_seed(11)
    X = torch.randn(n, F, device=DEV, generator=g)
    # CSR of the graph (row = dst): sort once with our plan, then time spmm on the CSR
    plan = planmod.build_plan(dst, n)
    rowptr = plan.rowptr
    col = plan.sorted_ids(src).to(torch.int64)
    val = torch.rand(e, device=DEV, generator=g)
    ms = timeit(lambda: gno_b200.spmm_csr(rowptr, col, val, X, "sum"), iters)
    try:
        csr = torch.sparse_csr_tensor(rowptr, col, val, (n, n))
        nat = round(timeit(lambda: torch.sparse.mm(csr, X), 3), 3)
        del csr
    except Exception as ex:
        nat = repr(ex)[:100]
    emit("C4 spmm CSR (reddit-shaped, F=256 fp32, plan cached)", ms, e, "nnz", agg_bytes(n, e, F, 4, weight=True),
         dram_key="c4spmm_sum",
         native_torch_ms={"torch.sparse.mm(csr, X) (cuSPARSE)": nat})
    # coalesced COO of the graph, then transpose / coalesce
    index = torch.stack([plan.erow.to(torch.int64), col])
    del plan
    gno_b200.clear_caches()
    ci, cv = gno_b200.coalesce(index, val, n, n)
    nnz = ci.size(1)
    del index
    ms = timeit(lambda: gno_b200.transpose(ci, cv, n, n), iters)
    emit(f"C4 transpose of coalesced COO ({nnz} nnz; order check + 32-bit key sort fast path)", ms, nnz, "nnz",
         2 * (16 + 4) * nnz, passes="3 x 8-bit over 18 row bits")
    perm = torch.randperm(2 * nnz, device=DEV, generator=g)
    dup_i = torch.cat([ci, ci], dim=1).index_select(1, perm)
    dup_v = torch.cat([cv, cv])
    del perm
    ms = timeit(lambda: gno_b200.coalesce(dup_i, dup_v, n, n), max(3, iters // 2))
    try:
        nat = round(timeit(lambda: torch.sparse_coo_tensor(dup_i, dup_v, (n, n)).coalesce(), 3), 3)
    except Exception as ex:
        nat = repr(ex)[:100]
    emit(f"C4 coalesce of 2x duplicated, permuted COO ({2 * nnz} entries)", ms, 2 * nnz, "nnz",
         (16 + 4) * 2 * nnz + (16 + 4) * nnz, passes="5 x 8-bit over 36 key bits (64-bit keys)",
         native_torch_ms={"sparse_coo_tensor(...).coalesce()": nat})


def run_sort(iters):
    g = torch.Generator(device=DEV).manual_seed(5)
    n = 1 << 28
    x = torch.rand(n, device=DEV, generator=g)
    ms = timeit(lambda: gno_b200.sort(x), max(3, iters // 2))
    nat = timeit(lambda: torch.sort(x, stable=True), 3)
    emit(f"sort fp32 1-D {n} (values + int64 indices, stable)", ms, n, "keys", n * (4 + 4 + 8), passes="4 x 8-bit",
         native_torch_ms={"torch.sort(stable=True) (CUB)": round(nat, 3)})
    y = torch.rand(20000, 20000, device=DEV, generator=g)
    for dim in (0, 1):
        ms = timeit(lambda: gno_b200.sort(y, dim), max(3, iters // 2))
        nat = timeit(lambda: torch.sort(y, dim=dim, stable=True), 3)
        emit(f"sort fp32 (20000,20000) dim{dim}", ms, y.numel(), "keys", y.numel() * 16, passes="6 x 8-bit (47-bit key)",
             native_torch_ms={"torch.sort(stable=True)": round(nat, 3)})
    k = torch.randint(0, 1 << 31, (n,), device=DEV, generator=g, dtype=torch.int64).to(torch.int32)
    v = torch.arange(n, device=DEV, dtype=torch.int32)
    ms = timeit(lambda: gno_b200.sort_pairs(k, v), max(3, iters // 2))
    nat = timeit(lambda: torch.sort(k, stable=True), 3)
    emit(f"sort_pairs u32 keys + u32 payload, {n}", ms, n, "keys", n * 16, passes="4 x 8-bit",
         native_torch_ms={"torch.sort(int32 keys, stable=True): values + int64 indices (CUB)": round(nat, 3)})


def run_c5(iters):
    """C5 per-GPU slice at P=8: RMAT-26 has 2^30 edges; one of 8 ranks aggregates 2^27 edges into
    its 2^23 destination rows from the full 2^26 x 128 bf16 feature matrix."""
    n_src, n_dst, e, F = 1 << 26, 1 << 23, 1 << 27, 128
    g = torch.Generator(device=DEV).manual_seed(9)
    # R-MAT (a,b,c,d = .57,.19,.19,.05): pick quadrant bits per level
    def rmat(bits, count):
        ids_r = torch.zeros(count, dtype=torch.int64, device=DEV)
        ids_c = torch.zeros(count, dtype=torch.int64, device=DEV)
        for _ in range(bits):
            u = torch.rand(count, device=DEV, generator=g)
            rb = (u >= 0.76).long()                      # c or d -> row bit 1
            cb = (((u >= 0.57) & (u < 0.76)) | (u >= 0.95)).long()  # b or d -> col bit 1
            ids_r = (ids_r << 1) | rb
            ids_c = (ids_c << 1) | cb
        return ids_r, ids_c
    dst, src = rmat(26, e)
    dst = dst >> 3  # this rank's 2^23 rows (edge-balanced ranges are chosen at plan time in dist mode)
    x = torch.randn(n_src, F, device=DEV, generator=g, dtype=torch.float32).to(torch.bfloat16)
    plan = planmod.build_plan(dst, n_dst)
    gidx = plan.sorted_ids(src)
    del dst, src
    ms = timeit(lambda: gno_b200.segment_reduce(plan, x, "sum", gidx=gidx), iters)
    emit("C5 per-GPU slice (P=8): RMAT-26 shard, 2^27 edges -> 2^23 rows, F=128 bf16", ms, e, "edges",
         agg_bytes(n_dst, e, F, 2), dram_key="c5_sum", max_row_len=plan.max_len, empty_rows=plan.n_empty)


RUNS = {"c1": run_c1, "c2": run_c2, "c3": run_c3, "c4": run_c4, "sort": run_sort, "c5": run_c5}

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=",".join(RUNS))
    ap.add_argument("--iters", type=int, default=10)
    a = ap.parse_args()
    for k in a.only.split(","):
        gno_b200.clear_caches()
        torch.cuda.empty_cache()
        try:
            RUNS[k](a.iters)
        except Exception as ex:  # keep going: partial tables are still useful
            print(json.dumps({"op": k, "error": repr(ex)[:500]}), flush=True)
