"""Host cost per torch_scatter call on the reference scripts' own shapes (fp16 [L, L], full-shape
int64 index, dims 0 / 1 — op_bm_scripts/benchmark_scatter_add.py:60-84): wall time per call as
`timeit(100)` sees it (call + synchronize, like torch.utils.benchmark.Timer on CUDA) next to the
device time of the same calls (CUDA events around the loop), for the shim and for the native torch
op the scripts time beside it.  One JSON line per case.

    PYTHONPATH=gnn-ops-benchmark_b200 python profiles/host_overhead.py
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gnn-ops-benchmark_b200"))
import torch  # noqa: E402
import torch_scatter  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(42)


def measure(fn, n=200):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / n * 1e6
    return wall, a.elapsed_time(b) / n * 1e3


for L in (223, 707, 1414):
    for dim in (0, 1):
        src = torch.rand(L, L, device=dev, dtype=torch.float16)
        idx = torch.randint(0, L, (L, L), device=dev)
        idx1 = idx[0].contiguous() if dim == 1 else idx[:, 0].contiguous()
        cases = {
            "scatter_add (shim)": lambda: torch_scatter.scatter_add(src, idx, dim=dim),
            "scatter_max (shim)": lambda: torch_scatter.scatter_max(src, idx, dim=dim),
            "scatter_mean (shim)": lambda: torch_scatter.scatter_mean(src, idx, dim=dim),
            "scatter_add, 1-D index (shim, plan cached)": lambda: torch_scatter.scatter_add(src, idx1, dim=dim),
            "native zeros_like + scatter_add_": lambda: torch.zeros_like(src).scatter_add_(dim, idx, src),
            "native zeros + scatter_reduce_(amax)": lambda: torch.zeros_like(src).scatter_reduce_(
                dim, idx, src, "amax", include_self=False),
        }
        for name, fn in cases.items():
            wall, devt = measure(fn)
            print(json.dumps({"shape": [L, L], "dim": dim, "call": name, "wall_us_per_call": round(wall, 1),
                              "device_us_per_call": round(devt, 1),
                              "host_bound": wall > 1.3 * devt}), flush=True)
