"""Run one case of profiles/bench_ops.py a few times (for ncu):  python profiles/prof_case.py c5 sum"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "profiles"))
sys.path.insert(0, os.path.join(ROOT, "gnn-ops-benchmark_b200"))
import torch  # noqa: E402

import bench as B  # noqa: E402
import gno_b200  # noqa: E402
from gno_b200 import plan as planmod  # noqa: E402

DEV = torch.device("cuda:0")
case = sys.argv[1]
red = sys.argv[2] if len(sys.argv) > 2 else "sum"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
g = torch.Generator(device=DEV).manual_seed(9)
if case == "c5":
    n_src, n_dst, e, F = 1 << 26, 1 << 23, 1 << 27, 128
    ids_r = torch.zeros(e, dtype=torch.int64, device=DEV)
    ids_c = torch.zeros(e, dtype=torch.int64, device=DEV)
    for _ in range(26):
        u = torch.rand(e, device=DEV, generator=g)
        ids_r = (ids_r << 1) | (u >= 0.76).long()
        ids_c = (ids_c << 1) | (((u >= 0.57) & (u < 0.76)) | (u >= 0.95)).long()
    dst, src = ids_r >> 3, ids_c
    x = torch.randn(n_src, F, device=DEV, generator=g).to(torch.bfloat16)
    n = n_dst
elif case == "rmat26":   # bench.py's default workload at N=1: the full graph
    n, e, F, dtype = B.workload_shape("rmat26")
    src, dst = B.rmat_edges(26, e, DEV, 42)
    x = B.feature_block(0, 1, n, F, dtype, DEV)
elif case == "c4spmm":   # C4: CSR spmm with edge values on the Reddit-shaped graph, F=256 fp32
    n, e, _, _, ex, off = B.WORKLOADS["reddit"]
    F = 256
    src, dst = B.make_graph(n, n, e, ex, off, DEV, 42)
    x = torch.randn(n, F, device=DEV, generator=g)
elif case in ("reddit_bf16", "reddit", "products"):
    n, e, F, dtype, ex, off = B.WORKLOADS[case]
    src, dst = B.make_graph(n, n, e, ex, off, DEV, 42)
    x = torch.randn(n, F, device=DEV, generator=g).to(dtype)
elif case == "c1":
    n, e, F = 100_000, 1_000_000, 64
    dst = torch.randint(0, n, (e,), device=DEV, generator=g)
    src = torch.arange(e, device=DEV)
    x = torch.rand(e, F, device=DEV, generator=g)
plan = planmod.build_plan(dst, n)
gidx = plan.sorted_ids(src)
del src, dst
arg = red in ("max", "min")
w = torch.rand(plan.E, device=DEV, generator=g) if case == "c4spmm" else None
torch.cuda.synchronize()
for _ in range(reps):
    gno_b200.segment_reduce(plan, x, red, gidx=gidx, eid=plan.perm, want_arg=arg, weights=w)
torch.cuda.synchronize()
print("done", case, red, plan.chunk_len, plan.n_span, plan.n_empty)
