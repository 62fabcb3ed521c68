"""python profiles/prof_fullshape.py [L] — the reference scripts' full-shape call (fp16 [L,L], int64 index of
the same shape, dims 0 and 1; benchmark_scatter_add.py:60-84) a few times per (reduce, dim): the first call on
an index tensor runs scatter_onchip_kernel, later ones scatter_planned_kernel (for ncu -k regex:scatter_)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gnn-ops-benchmark_b200"))
import torch  # noqa: E402

import gno_b200  # noqa: E402

L = int(sys.argv[1]) if len(sys.argv) > 1 else 6708
reds = sys.argv[2].split(",") if len(sys.argv) > 2 else ["sum", "max"]
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(42)
s16 = torch.rand(L, L, device=dev, generator=g).half()
ifull = torch.randint(0, L, (L, L), device=dev, generator=g)
for red in reds:
    for dim in (0, 1):
        for _ in range(3):
            gno_b200.scatter(s16, ifull, dim, None, L, red, return_arg=True)
torch.cuda.synchronize()
print("done")
