"""python profiles/prof_sort.py [n] [bits] — run sort_pairs a few times (for ncu)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gnn-ops-benchmark_b200"))
import torch  # noqa: E402

import gno_b200  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 26
bits = int(sys.argv[2]) if len(sys.argv) > 2 else 32
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
k = torch.randint(0, 1 << min(bits, 31), (n,), device=dev, generator=g, dtype=torch.int64).to(torch.int32)
v = torch.arange(n, device=dev, dtype=torch.int32)
for _ in range(2):
    gno_b200.sort_pairs(k, v, 0, bits)
torch.cuda.synchronize()
print("done")
