"""Sweep of the K-stage needed-rows exchange on the RMAT workload (bench.py's N>1 path): for each
(exchange, stages, stage shares) build the partitioned aggregator once and time
  exchange-only (all stages back to back), reduce-only (all stage plans), and the pipelined step.
One JSON line per setting on rank 0 (max over ranks, CUDA events, 3 warm-ups).

    torchrun --nproc-per-node N --master-addr 127.0.0.1 profiles/rmat_stage_sweep.py \
        [--workload rmat26] [--steps 10] [--settings "push:1;push:4;push:4:0.1,0.2,0.3,0.4;needed:4"]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "gnn-ops-benchmark_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench as B  # noqa: E402
import gno_b200  # noqa: E402
from gno_b200.dist import DistAggregator, edge_balanced_ranges  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="rmat26")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--row-weight", type=int, default=4)
    ap.add_argument("--settings", default="push:1;push:2::source;push:3:0.1,0.9:source;push:4:0.05,0.25,0.7:source",
                    help="';'-separated exchange:stages[:fracs[:split[:ownership[:push_blocks[:merge_own[:push_chunk]]]]]]")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    N, E, F, dtype = B.workload_shape(args.workload)
    scale = N.bit_length() - 1
    src, dst = B.rmat_edges(scale, E, dev, 42)
    counts = torch.bincount(dst, minlength=N)
    bounds = edge_balanced_ranges(counts + args.row_weight, world).cpu()
    del counts
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    m = (dst >= lo) & (dst < hi)
    src, dst = src[m], dst[m] - lo
    del m
    torch.cuda.empty_cache()
    x_local = B.feature_block(rank, world, N, F, dtype, dev)
    out = torch.empty(hi - lo, F, device=dev, dtype=dtype)

    def timed(fn, steps):
        for _ in range(3):
            fn()
        dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            fn()
        b.record()
        dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / steps], device=dev)
        every = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(every, t)
        timed.per_rank = [round(float(v), 3) for v in every]
        return max(timed.per_rank)

    for setting in args.settings.split(";"):
        parts = setting.split(":")
        mode, K = parts[0], int(parts[1])
        fracs = [float(v) for v in parts[2].split(",")] if len(parts) > 2 and parts[2] else None
        split = parts[3] if len(parts) > 3 and parts[3] else "dest"
        own = parts[4] if len(parts) > 4 and parts[4] else "cyclic"
        pblocks = int(parts[5]) if len(parts) > 5 and parts[5] else 148
        merge = bool(int(parts[6])) if len(parts) > 6 and parts[6] else False
        chunk = int(parts[7]) if len(parts) > 7 and parts[7] else 0
        gno_b200.clear_caches()
        torch.cuda.empty_cache()
        t0 = time.perf_counter()
        agg = DistAggregator(bounds, src, dst, rank=rank, world=world, cyclic_rows=N, exchange=mode, stages=K,
                             stage_fracs=fracs, row_weight=args.row_weight, split=split, ownership=own,
                             push_blocks=pblocks, merge_own=merge, push_chunk=chunk)
        plans = agg.xstage_plans()
        if mode == "push":
            recv = agg._push_buffer(F, dtype, dev)[0][:agg.n_needed]
        else:
            recv = torch.empty(agg.n_needed, F, device=dev, dtype=dtype)
        torch.cuda.synchronize()
        setup_s = time.perf_counter() - t0
        step_ms = timed(lambda: agg.aggregate(x_local, "sum", x_full=recv if mode == "needed" else None, out=out),
                        args.steps)
        # timeline of one staged step on every rank (events on both streams, ms from the step start)
        timeline = None
        if K > 1:
            dist.barrier()
            torch.cuda.synchronize()
            agg.trace = []
            agg.aggregate(x_local, "sum", x_full=recv if mode == "needed" else None, out=out)
            torch.cuda.synchronize()
            t0e = agg.trace[0][1]
            timeline = {lab: round(t0e.elapsed_time(e), 3) for lab, e in agg.trace[1:]}
            agg.trace = None
            if rank != 0:
                print(json.dumps({"rank": rank, "timeline_ms": timeline}), flush=True)
        xonly = timed((lambda: agg.exchange_push(x_local)) if mode == "push" else
                      (lambda: agg.exchange_needed(x_local, recv)), args.steps)

        ronly = timed(lambda: agg.reduce_stages(recv, "sum", out, x_local=x_local), args.steps)
        ronly_ranks = timed.per_rank
        rows_edges = torch.tensor([float(hi - lo), float(src.numel())], device=dev, dtype=torch.float64)
        allre = [torch.empty_like(rows_edges) for _ in range(world)]
        dist.all_gather(allre, rows_edges)
        info = torch.tensor([agg.n_needed, agg.n_needed - agg.recv_splits[rank], src.numel()], device=dev,
                            dtype=torch.float64)
        mx = info.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(json.dumps({"workload": args.workload, "n_gpus": world, "exchange": mode, "stages": K, "split": split, "ownership": own, "push_blocks": pblocks, "merge_own": merge, "push_chunk": chunk,
                              "stage_fracs": fracs, "stage_edges_rank0": [p.E for p, *_ in plans], "step_ms": round(step_ms, 3),
                              "exchange_only_ms": round(xonly, 3), "reduce_only_ms": round(ronly, 3),
                              "reduce_only_ms_per_rank": ronly_ranks,
                              "rows_per_rank": [int(v[0]) for v in allre], "edges_per_rank": [int(v[1]) for v in allre],
                              "edges_per_s": E / (step_ms * 1e-3),
                              "recv_rows_max": int(mx[0]), "remote_rows_max": int(mx[1]),
                              "remote_GB_max": round(float(mx[1]) * F * 2 / 1e9, 3), "edges_max": int(mx[2]),
                              "stage_recv_rows_rank0": [agg.stage_row0[s + 1] - agg.stage_row0[s] for s in range(K)],
                              "setup_s": round(setup_s, 1), "timeline_ms_rank0": timeline}), flush=True)
        del agg, plans, recv
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
