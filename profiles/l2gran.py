"""cudaLimitMaxL2FetchGranularity experiment: time the gather-reduce at 32 / 64 / 128-byte L2 fetch
granularity.  python profiles/l2gran.py [products reddit_bf16 c5 c1]
Result (r1m, profiles/r1m_l2gran.log): the limit reads back as set but changes nothing on B200
(products sum 4.762 / 4.762 / 4.762 ms at 64 / 32 / 128) — a hint the hardware ignores."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gnn-ops-benchmark_b200"))
import torch  # noqa: E402

import bench as B  # noqa: E402
import gno_b200  # noqa: E402
from gno_b200 import plan as planmod  # noqa: E402

DEV = torch.device("cuda:0")
torch.cuda.init()
torch.zeros(1, device=DEV)
import ctypes  # noqa: E402

_rt = ctypes.CDLL("libcudart.so.12")  # already loaded by torch
_LIMIT = 0x05  # cudaLimitMaxL2FetchGranularity


def set_gran(nbytes):
    """Set (nbytes > 0) or query (0) the limit; returns the value before the call."""
    cur = ctypes.c_size_t()
    assert _rt.cudaDeviceGetLimit(ctypes.byref(cur), _LIMIT) == 0
    if nbytes:
        assert _rt.cudaDeviceSetLimit(_LIMIT, ctypes.c_size_t(nbytes)) == 0
    return cur.value


def make(case):
    g = torch.Generator(device=DEV).manual_seed(9)
    if case == "c5":
        n_src, n, e, F = 1 << 26, 1 << 23, 1 << 27, 128
        ids_r = torch.zeros(e, dtype=torch.int64, device=DEV)
        ids_c = torch.zeros(e, dtype=torch.int64, device=DEV)
        for _ in range(26):
            u = torch.rand(e, device=DEV, generator=g)
            ids_r = (ids_r << 1) | (u >= 0.76).long()
            ids_c = (ids_c << 1) | (((u >= 0.57) & (u < 0.76)) | (u >= 0.95)).long()
        dst, src = ids_r >> 3, ids_c
        x = torch.randn(n_src, F, device=DEV, generator=g).to(torch.bfloat16)
    elif case == "c1":
        n, e, F = 100_000, 1_000_000, 64
        dst = torch.randint(0, n, (e,), device=DEV, generator=g)
        src = torch.arange(e, device=DEV)
        x = torch.rand(e, F, device=DEV, generator=g)
    else:
        n, e, F, dtype, ex, off = B.WORKLOADS[case]
        src, dst = B.make_graph(n, n, e, ex, off, DEV, 42)
        x = torch.randn(n, F, device=DEV, generator=g).to(dtype)
    plan = planmod.build_plan(dst, n)
    gidx = plan.sorted_ids(src)
    return plan, gidx, x, e


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


cases = sys.argv[1:] or ["products", "reddit_bf16", "c5"]
print("default granularity:", set_gran(0))
for case in cases:
    plan, gidx, x, e = make(case)
    for red in (("sum", "max") if case != "c5" else ("sum",)):
        arg = red == "max"
        for gran in (64, 32, 128, 64, 32):
            set_gran(gran)
            now = set_gran(0)
            ms = timeit(lambda: gno_b200.segment_reduce(plan, x, red, gidx=gidx, eid=plan.perm, want_arg=arg))
            print(json.dumps({"case": case, "reduce": red, "granularity_set": gran, "granularity_now": now,
                              "ms": round(ms, 4), "gedges_s": round(e / ms / 1e6, 3)}), flush=True)
    del plan, gidx, x
    gno_b200.clear_caches()
    torch.cuda.empty_cache()
