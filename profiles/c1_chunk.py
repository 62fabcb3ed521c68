"""C1-shaped scatter_sum (F=64 fp32, L2 flushed) at three sizes over the chunk length of the plan
(GNO_SEG_PDL=0 disables the programmatic dependent launch of the finish kernel)."""
import os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gnn-ops-benchmark_b200"))
import torch, gno_b200
from gno_b200 import plan as planmod
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(42)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for (E, N, F) in ((1_000_000, 100_000, 64), (4_000_000, 400_000, 64), (250_000, 25_000, 64)):
    src = torch.rand(E, F, device=dev, generator=g)
    idx = torch.randint(0, N, (E,), device=dev, generator=g)
    out = torch.empty(N, F, device=dev)
    for cl in (32, 64, 96, 128, 160, 192, 256):
        plan = planmod.build_plan(idx, N, chunk_len=cl)
        ts = []
        for _ in range(15):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            gno_b200.segment_reduce(plan, src, "sum", gidx=plan.perm, out=out)
            b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
        print(E, "chunk", cl, "median us", round(statistics.median(ts[3:]) * 1e3, 1), "span", plan.n_span)
