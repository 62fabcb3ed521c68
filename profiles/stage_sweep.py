"""Products-shaped weak-scaling step (bench.py's N>1 workload) timed for several exchange set-ups
in ONE launch: NCCL all-gather then reduce (stages=1), the K-stage pipelined all-gather, and the
peer-store all-gather.

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 profiles/stage_sweep.py [K ...]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gnn-ops-benchmark_b200"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench as B  # noqa: E402
import gno_b200  # noqa: E402
from gno_b200.dist import DistAggregator  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local_rank = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local_rank)
dev = torch.device("cuda", local_rank)
hipri = os.environ.get("GNO_NCCL_HIPRI", "1") != "0"
opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=hipri)
dist.init_process_group("nccl", device_id=dev, pg_options=opts)

n_local, e_local, F, dtype, ex, off = B.WORKLOADS["products"]
src, dst = B.make_graph(n_local, n_local * world, e_local, ex, off, dev, 42 + rank)
g = torch.Generator(device=dev).manual_seed(1000 + rank)
x_local = torch.randn(n_local, F, device=dev, generator=g)
bounds = torch.arange(world + 1, dtype=torch.int64) * n_local
out = torch.empty(n_local, F, device=dev)
steps = 10


def timed(step):
    for _ in range(3):
        step()
    dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        step()
    b.record()
    dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / steps], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


ref = None
ks = [a for a in sys.argv[1:] if a != "nopush"] or ["1", "2", "4"]
configs = [("allgather", int(k)) for k in ks] + ([] if "nopush" in sys.argv else [("allgather_push", 1)])
for mode, k in configs:
    gno_b200.clear_caches()
    torch.cuda.empty_cache()
    try:
        agg = DistAggregator(bounds, src, dst, rank=rank, world=world, stages=k, exchange=mode)
        x_full = None
        bufs = None
        if mode == "allgather_push":
            x_full = agg.exchange_allgather_push(x_local)
        else:
            x_full = torch.empty(world * agg.max_rows, F, device=dev)
        agg.plan()
        if k > 1:
            agg.stage_plans()
            bufs = [torch.empty(world * r, F, device=dev) for r in agg.stage_rows]
        ms = timed(lambda: agg.aggregate(x_local, "sum", x_full=x_full, out=out, stage_bufs=bufs))
        chk = float(out.double().abs().sum().item())
        if ref is None:
            ref = chk
        if rank == 0:
            print(json.dumps({"gpus": world, "exchange": mode, "stages": k, "nccl_high_priority": hipri,
                              "ms_per_step": round(ms, 3),
                              "gedges_s": round(e_local * world / ms / 1e6, 2),
                              "checksum_rel_diff": abs(chk - ref) / max(ref, 1e-30)}), flush=True)
        del agg, x_full, bufs
    except Exception as e:  # keep the sweep going
        if rank == 0:
            print(json.dumps({"gpus": world, "exchange": mode, "stages": k, "error": repr(e)[:300]}), flush=True)
dist.destroy_process_group()
