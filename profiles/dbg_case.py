import os, sys, random
R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "gnn-ops-benchmark_b200"))
import torch, oracle, gno_b200
from gno_b200 import plan as planmod
cuda = torch.device("cuda:0")
seed = 3
rnd = random.Random(seed); g = torch.Generator().manual_seed(seed)
for case in range(14):
    dtype = rnd.choice([torch.float32, torch.bfloat16, torch.float16])
    F = rnd.choice([1, 2, 3, 4, 5, 8, 12, 16, 31, 32, 33, 64, 100, 127, 128, 130, 256, 301, 602])
    N = rnd.choice([1, 2, 7, 100, 1000, 5000])
    E = rnd.choice([0, 1, 31, 32, 33, 255, 256, 257, 1000, 4096, 20000, 50001])
    n_src = rnd.choice([1, 50, 777]); cl = rnd.choice([32, 64, 128, 256])
    reduce = rnd.choice(["sum", "mean", "max", "min", "mul"]); skew = rnd.choice([1, 3, 6])
    x = torch.randn(n_src, F, generator=g)
    if reduce in ("min", "max"): x = (x * 2).round() / 2
    if reduce == "mul": x = torch.rand(n_src, F, generator=g) * 0.2 + 0.9
    x = x.to(dtype)
    dst = (torch.rand(E, generator=g) ** skew * N).long().clamp_(0, max(N - 1, 0))
    src = torch.randint(0, n_src, (E,), generator=g)
    if case != 3: continue
    print(dtype, F, N, E, n_src, cl, reduce, skew)
    for red in ("sum", "mean"):
        want, _ = oracle.gather_scatter(x, src, dst, N, red)
        w64 = torch.zeros(N, F, dtype=torch.float64).index_add_(0, dst, x.double()[src])
        cnt = torch.bincount(dst, minlength=N).clamp(min=1).double().view(-1, 1)
        if red == "mean": w64 = w64 / cnt
        plan = planmod.build_plan(dst.to(cuda), N, chunk_len=cl)
        gidx = plan.sorted_ids(src.to(cuda))
        got = gno_b200.segment_reduce(plan, x.to(cuda), red, gidx=gidx).cpu()
        e_or = (want.double() - w64).abs(); e_gpu = (got.double() - w64).abs()
        print(red, "oracle-vs-f64 max", float(e_or.max()), "gpu-vs-f64 max", float(e_gpu.max()), "rows deg", torch.bincount(dst, minlength=N)[:4].tolist(), "n_span", plan.n_span)
        bad = (e_gpu > 1e-3).nonzero()
        print("  bad count", bad.shape[0], bad[:5].tolist())
    red = "mean"
    want, _ = oracle.gather_scatter(x, src, dst, N, red)
    scale = oracle.gather_scatter(x.float().abs(), src, dst, N, red)[0]
    exact = torch.zeros(N, F, dtype=torch.float64).index_add_(0, dst, x.double()[src]) / torch.bincount(dst, minlength=N).clamp(min=1).double().view(-1, 1)
    plan = planmod.build_plan(dst.to(cuda), N, chunk_len=cl)
    got = gno_b200.segment_reduce(plan, x.to(cuda), red, gidx=plan.sorted_ids(src.to(cuda))).cpu()
    err = (got - want).abs(); lim = 1e-5 * torch.maximum(scale, want.abs()) + (want.double() - exact).abs().float()
    r = err / lim; i = int(r.argmax()); row, col = i // F, i % F
    print("worst ratio", float(r.max()), "row", row, "col", col, "deg", int(torch.bincount(dst, minlength=N)[row]), "got", float(got[row, col]), "want", float(want[row, col]), "exact", float(exact[row, col]), "x", float(x[0, col]))
