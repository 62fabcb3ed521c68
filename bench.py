"""bench.py — headline benchmark of the aggregation path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload products]

A step is one pass of the hot path over one synthetic graph: the fused
index_select→scatter_add message passing of BASELINE.json configs[1]
(ogbn-products-shaped: 2,449,029 nodes, 61,859,140 edges, F=100 fp32).
`value` = aggregated edges/s with the graph plan and features resident in
HBM; `e2e` = the same through the public host-buffer API (pinned host x and
edge_index → device, plan build, aggregation, result → host).  At N>1 every
rank owns a contiguous destination range with its own products-shaped edge
shard (weak scaling), source features are all-gathered with NCCL, outputs stay
partitioned.  One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "gnn-ops-benchmark_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402

WORKLOADS = {
    # name: (nodes, edges, features, dtype, zipf exponent, zipf offset)
    "products": (2_449_029, 61_859_140, 100, torch.float32, 0.75, 100.0),
    "c1": (100_000, 1_000_000, 64, torch.float32, 0.0, 1.0),
    "reddit": (232_965, 114_615_892, 602, torch.float32, 0.8, 300.0),
    "reddit_bf16": (232_965, 114_615_892, 602, torch.bfloat16, 0.8, 300.0),
}
METRIC = "aggregation edges/s (fused index_select->scatter_add, products-shaped graph)"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def make_graph(n_dst, n_src, n_edges, exponent, offset, device, seed):
    """Synthetic power-law graph: destination in-degrees follow a shifted Zipf law
    (weights (i+offset)^-exponent over a random node permutation), sources uniform."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    if exponent > 0:
        w = (torch.arange(n_dst, device=device, dtype=torch.float64) + offset) ** (-exponent)
        cdf = torch.cumsum(w, 0)
        cdf = (cdf / cdf[-1]).to(torch.float32)
        u = torch.rand(n_edges, device=device, generator=g)
        rank = torch.searchsorted(cdf, u).clamp_(max=n_dst - 1)
        relabel = torch.randperm(n_dst, device=device, generator=g)
        dst = relabel[rank]
        del w, cdf, u, rank, relabel
    else:
        dst = torch.randint(0, n_dst, (n_edges,), device=device, generator=g)
    src = torch.randint(0, n_src, (n_edges,), device=device, generator=g)
    return src, dst


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def algorithmic_bytes(n_rows, n_edges, F, es):
    """E*(F*s + 4) + N*(F*s_out + 4)."""
    return n_edges * (F * es + 4) + n_rows * (F * es + 4)


def cpu_reference_step(x, src, dst, n_rows):
    """What the reference executes on CPU for this op: messages = x.index_select(0, src)
    (benchmark_fused_index_select_reduce.py:12-15 / PyG propagate), then
    torch_scatter.scatter_add == zeros.scatter_add_(0, expanded index, messages)
    (torch-scatter 2.0.9 scatter_sum; call site benchmark_scatter_add.py:18)."""
    msgs = x.index_select(0, src)
    out = torch.zeros(n_rows, x.size(1), dtype=x.dtype)
    out.scatter_add_(0, dst.view(-1, 1).expand(-1, x.size(1)), msgs)
    return out


def time_cpu_baseline(n_nodes, n_edges, F, dtype, exponent, offset, budget_s, steps, warmup):
    """Reference CPU path on a bounded sample of the same workload (same node set,
    first `sample` edges), all host threads. Returns (edges/s, sample, cores, ms/step)."""
    torch.manual_seed(42)
    probe = min(n_edges, 1_000_000)
    src, dst = make_graph(n_nodes, n_nodes, probe, exponent, offset, "cpu", 42)
    x = torch.randn(n_nodes, F, dtype=torch.float32).to(dtype)
    cpu_reference_step(x, src, dst, n_nodes)
    t0 = time.perf_counter()
    cpu_reference_step(x, src, dst, n_nodes)
    t_probe = time.perf_counter() - t0
    per_edge = t_probe / probe
    sample = int(min(n_edges, max(probe, budget_s / max(steps + warmup, 1) / per_edge)))
    if sample != probe:
        src, dst = make_graph(n_nodes, n_nodes, sample, exponent, offset, "cpu", 42)
    for _ in range(warmup):
        cpu_reference_step(x, src, dst, n_nodes)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_reference_step(x, src, dst, n_nodes)
    dt = (time.perf_counter() - t0) / steps
    return sample / dt, sample, torch.get_num_threads(), dt * 1e3


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_nodes, n_edges, F, dtype, exponent, offset = WORKLOADS[args.workload]
    v, sample, cores, ms = time_cpu_baseline(n_nodes, n_edges, F, dtype, exponent, offset,
                                             budget_s=150.0, steps=args.steps, warmup=args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "edges/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32" if dtype == torch.float32 else "bf16",
        "data": "synthetic",
        "config": {"workload": args.workload, "nodes": n_nodes, "edges": n_edges, "features": F},
        "cpu_baseline": {"value": v, "unit": "edges/s", "cores": cores, "kind": "port",
                         "sample": f"first {sample} of {n_edges} edges, same {n_nodes}-node feature "
                                   "matrix; torch CPU index_select + scatter_add_ (what torch_scatter "
                                   "2.0.9 scatter_sum executes)"},
        "e2e": {"value": v, "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit_json(line)


def run_ours(args):
    import gno_b200
    from gno_b200 import plan as planmod

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        # High-priority NCCL stream: the chunk all-gathers of the staged exchange then get SMs as
        # aggregation CTAs retire instead of queueing behind the whole kernel (without it the
        # stages serialise: 10.0 -> 10.3 ms at N=4; with it 8.6 ms — profiles/scaling/stage_sweep_*).
        opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), pg_options=opts)
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if args.stages <= 0:
        # measured (profiles/scaling/stage_sweep_{4,8}gpu_hipri.jsonl): 4 stages 8.63 vs 10.00 ms at
        # N=4, 13.30 vs 15.52 ms at N=8; at N=2 the peer-store all-gather (one stage) is used
        args.stages = 4 if world >= 4 else 1
    n_local, e_local, F, dtype, exponent, offset = WORKLOADS[args.workload]
    es = torch.empty((), dtype=dtype).element_size()
    n_global = n_local * world

    # ---- synthetic inputs, resident in HBM ------------------------------------------------
    src, dst = make_graph(n_local, n_global, e_local, exponent, offset, dev, 42 + rank)
    gx = torch.Generator(device=dev)
    gx.manual_seed(1000 + rank)
    x_local = torch.randn(n_local, F, device=dev, generator=gx, dtype=torch.float32).to(dtype)
    x_full = torch.empty(n_global, F, device=dev, dtype=dtype) if world > 1 else x_local
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    agg = None
    if world > 1:
        from gno_b200.dist import DistAggregator
        bounds = torch.arange(world + 1, dtype=torch.int64) * n_local
        # measured (profiles/scaling): the peer-store all-gather beats NCCL at N=2 (6.64 vs 7.18 ms),
        # ties at N=4 (10.16 vs 10.04) and loses at N=8 (17.4 vs 15.6)
        xmode = args.exchange if args.exchange in ("allgather", "allgather_push") else \
            ("allgather_push" if world <= 2 else "allgather")
        try:
            agg = DistAggregator(bounds, src, dst, rank=rank, world=world, stages=args.stages, exchange=xmode)
            if xmode == "allgather_push":
                x_full = agg.exchange_allgather_push(x_local)  # symmetric-memory rendezvous happens here
        except Exception as ex:  # no peer mapping on this box: NCCL all-gather
            if xmode != "allgather_push":
                raise
            print(f"allgather_push unavailable ({ex!r}); using the NCCL all-gather", file=sys.stderr)
            xmode = "allgather"
            agg = DistAggregator(bounds, src, dst, rank=rank, world=world, stages=args.stages)
        plan, gidx = agg.plan()
        if args.stages > 1:
            agg.stage_plans()
            stage_bufs = [torch.empty(world * r, F, device=dev, dtype=dtype) for r in agg.stage_rows]
    else:
        plan = planmod.build_plan(dst, n_local)
        gidx = plan.sorted_ids(src)
    torch.cuda.synchronize()
    plan_ms = (time.perf_counter() - t0) * 1e3
    out = torch.empty(n_local, F, device=dev, dtype=dtype)

    def step():
        if world > 1:  # NCCL all-gather of the feature shards, then the local gather-reduce
            agg.aggregate(x_local, "sum", x_full=x_full, out=out,
                          stage_bufs=stage_bufs if args.stages > 1 else None)
        else:
            gno_b200.segment_reduce(plan, x_full, "sum", gidx=gidx, out=out)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = gno_b200.launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    ev[0].record()
    for i in range(args.steps):
        step()
        ev[i + 1].record()
    barrier()
    launches = gno_b200.launch_count() - launches0
    clocks = sampler.stop()
    total_ms = ev[0].elapsed_time(ev[-1])
    if world > 1:
        t = torch.tensor([total_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = e_local * world / (ms_per_step * 1e-3)

    # ---- kernel-only duration for the roofline (aggregation kernel alone, this rank) -------
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
           for _ in range(args.steps)]
    for a, b in kev:
        a.record()
        gno_b200.segment_reduce(plan, x_full, "sum", gidx=gidx, out=out)
        b.record()
    torch.cuda.synchronize()
    k_ms = statistics.mean(a.elapsed_time(b) for a, b in kev)
    peak, peak_src = peaks()
    abytes = algorithmic_bytes(n_local, e_local, F, es)
    achieved = abytes / (k_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get(args.workload)

    # ---- e2e: host buffers through the public API ------------------------------------------
    # Every step copies that step's inputs (x and the int64 edge_index) from pinned host memory,
    # builds the plan, aggregates and copies the result back to pinned host memory.  N=1 uses
    # gno_b200.host.HostPipeline (two steps in flight: the D2H of step i overlaps the H2D of
    # step i+1); N>1 runs the steps back to back through DistAggregator.
    e2e = None
    cpu = None
    x_host = x_local.cpu().pin_memory()
    ei_host = torch.stack([src, dst]).cpu().pin_memory()
    out_hosts = [torch.empty(n_local, F, dtype=dtype).pin_memory() for _ in range(2)]
    e2e_steps = max(4, min(args.steps, 10))
    del src, dst
    torch.cuda.empty_cache()
    if world == 1:
        from gno_b200.host import HostPipeline, gather_scatter_host
        gno_b200.clear_caches()
        # single call latency (no cross-step overlap)
        gather_scatter_host(x_host, ei_host, n_local, "sum", out_hosts[0])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        gather_scatter_host(x_host, ei_host, n_local, "sum", out_hosts[0])
        single_ms = (time.perf_counter() - t0) * 1e3
        pipe = HostPipeline(n_local, "sum")
        for i in range(2):  # warm-up
            pipe.submit(x_host, ei_host, out_hosts[i % 2])
            pipe.result()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pipe.submit(x_host, ei_host, out_hosts[0])
        for i in range(1, e2e_steps):
            pipe.submit(x_host, ei_host, out_hosts[i % 2])
            pipe.result()
        pipe.result()
        torch.cuda.synchronize()
        e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
        includes = ("H2D x + int64 edge_index (pinned), plan build (dst radix sort), aggregation, D2H out; "
                    "2 steps in flight (HostPipeline), wall clock over %d steps" % e2e_steps)
    else:
        from gno_b200.dist import DistAggregator
        single_ms = None

        def e2e_step():
            gno_b200.clear_caches()
            xd = x_host.to(dev, non_blocking=True)
            eid = ei_host.to(dev, non_blocking=True)
            a2 = DistAggregator(bounds, eid[0], eid[1], rank=rank, world=world)
            o = a2.aggregate(xd, "sum", x_full=x_full)
            out_hosts[0].copy_(o, non_blocking=True)

        e2e_step()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(e2e_steps):
            e2e_step()
        b.record()
        barrier()
        e_ms = a.elapsed_time(b) / e2e_steps
        t = torch.tensor([e_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e_ms = float(t.item())
        includes = "H2D x + int64 edge_index (pinned), plan build, NCCL all-gather, aggregation, D2H out"
    e2e = {"value": e_local * world / (e_ms * 1e-3), "unit": "edges/s",
           "h2d_bytes_per_step": (x_host.numel() * es + ei_host.numel() * 8) * world,
           "d2h_bytes_per_step": out_hosts[0].numel() * es * world,
           "ms_per_step": e_ms, "single_call_ms": single_ms, "includes": includes}
    del x_host, ei_host
    if rank == 0 and world == 1:
        v, sample, cores, ms = time_cpu_baseline(n_local, e_local, F, dtype, exponent, offset,
                                                 budget_s=20.0, steps=3, warmup=1)
        cpu = {"value": v, "unit": "edges/s", "cores": cores, "kind": "port",
               "sample": f"first {sample} of {e_local} edges, same node set; torch CPU index_select + "
                         "scatter_add_ (the calls torch_scatter 2.0.9 scatter_sum makes), 3 steps"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "edges/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if dtype == torch.float32 else "bf16", "data": "synthetic",
            "config": {"workload": args.workload, "nodes_per_gpu": n_local, "edges_per_gpu": e_local,
                       "features": F, "index": "int64 edge_index -> cached dst-sorted CSR plan (int32)",
                       "l2": "inputs larger than L2 (x + col + out >> 126 MB); no flush needed",
                       "plan_build_ms": plan_ms, "max_row_len": plan.max_len,
                       "chunk_len": plan.chunk_len, "rows_cut_by_chunks": plan.n_span,
                       "empty_rows": plan.n_empty,
                       "parallelism": (f"dst-partitioned x{world}, all-gather of x by "
                                       + ("peer-store kernel over NVLink (gno_push_rows)" if xmode == "allgather_push"
                                          else f"NCCL, {args.stages}-stage exchange")) if world > 1 else "single GPU"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "kernel": "segreduce_staged_kernel (+ segfinish_kernel)", "kernel_ms": k_ms,
                         "algorithmic_bytes": abytes},
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        emit_json(line)
    if world > 1:
        dist.destroy_process_group()


def rmat_edges(scale, n_edges, device, seed, a=0.57, b=0.19, c=0.19):
    """R-MAT edge list (a,b,c,d = .57,.19,.19,.05): one quadrant choice per bit level.
    Returns (src = column ids, dst = row ids), int64, identical on every rank for one seed."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    rows = torch.zeros(n_edges, dtype=torch.int64, device=device)
    cols = torch.zeros(n_edges, dtype=torch.int64, device=device)
    for _ in range(scale):
        u = torch.rand(n_edges, device=device, generator=g)
        rows.mul_(2).add_((u >= a + b).to(torch.int64))                           # quadrants c, d
        cols.mul_(2).add_((((u >= a) & (u < a + b)) | (u >= a + b + c)).to(torch.int64))  # b, d
        del u
    return cols, rows


def run_rmat(args):
    """BASELINE.json configs[4]: dst-partitioned aggregation on an RMAT graph (default scale 26,
    2^30 edges, F=128 bf16), STRONG scaling: the graph is fixed, every rank owns an edge-balanced
    destination range and an equal block of the feature rows; features are all-gathered with NCCL,
    outputs stay partitioned."""
    import gno_b200
    from gno_b200 import plan as planmod
    from gno_b200.dist import DistAggregator, edge_balanced_ranges

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    scale = int(args.workload[4:])
    N, E, F, dtype, es = 1 << scale, 1 << (scale + 4), 128, torch.bfloat16, 2
    src, dst = rmat_edges(scale, E, dev, 42)
    counts = torch.bincount(dst, minlength=N)
    # balance kernel time, not just edges: a row costs its output write plus a row switch in the
    # kernel; measured on 4 B200s a weight of 4 edges per row gives the shortest step
    bounds = edge_balanced_ranges(counts + args.row_weight, world).cpu()
    del counts
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    if world > 1:
        m = (dst >= lo) & (dst < hi)
        src, dst = src[m], dst[m] - lo
        del m
    torch.cuda.empty_cache()
    e_local, n_out = src.numel(), hi - lo
    distinct_src = int(torch.unique(src).numel())
    xb = torch.arange(world + 1, dtype=torch.int64) * (N // world)
    gx = torch.Generator(device=dev)
    gx.manual_seed(1000 + rank)
    x_local = torch.randn(N // world, F, device=dev, generator=gx, dtype=torch.float32).to(dtype)
    x_full = torch.empty(N, F, device=dev, dtype=dtype) if world > 1 else x_local
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if world > 1:
        cyc = N if (args.cyclic and args.exchange != "allgather") else None
        try:
            agg = DistAggregator(bounds, src, dst, rank=rank, world=world, feature_bounds=xb,
                                 exchange=args.exchange, cyclic_rows=cyc)
            if args.exchange == "push":
                agg.exchange_push(x_local)  # symmetric-memory rendezvous happens here
        except Exception as ex:  # no peer mapping on this box: same exchange through NCCL all-to-all
            if args.exchange != "push":
                raise
            print(f"push exchange unavailable ({ex!r}); using the NCCL needed-rows all-to-all", file=sys.stderr)
            args.exchange = "needed"
            agg = DistAggregator(bounds, src, dst, rank=rank, world=world, feature_bounds=xb,
                                 exchange="needed", cyclic_rows=cyc)
        plan, gidx = agg.plan()
        if args.exchange == "needed":
            x_full = torch.empty(agg.n_needed, F, device=dev, dtype=dtype)
        elif args.exchange == "push":
            x_full = agg.exchange_push(x_local)  # the symmetric receive buffer
    else:
        plan = planmod.build_plan(dst, n_out)
        gidx = plan.sorted_ids(src)
    torch.cuda.synchronize()
    plan_ms = (time.perf_counter() - t0) * 1e3
    del src, dst
    torch.cuda.empty_cache()
    out = torch.empty(n_out, F, device=dev, dtype=dtype)

    def step():
        if world > 1:
            agg.aggregate(x_local, "sum", x_full=x_full, out=out)
        else:
            gno_b200.segment_reduce(plan, x_full, "sum", gidx=gidx, out=out)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = gno_b200.launch_count()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.steps):
        step()
    b.record()
    barrier()
    launches = gno_b200.launch_count() - launches0
    clocks = sampler.stop()
    total_ms = a.elapsed_time(b)
    ka, kb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ka.record()
    for _ in range(args.steps):
        gno_b200.segment_reduce(plan, x_full, "sum", gidx=gidx, out=out)
    kb.record()
    torch.cuda.synchronize()
    k_ms = ka.elapsed_time(kb) / args.steps
    stats = torch.tensor([total_ms, k_ms, float(e_local)], device=dev, dtype=torch.float64)
    if world > 1:
        mx = stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        total_ms, k_ms_max = float(mx[0]), float(mx[1])
    else:
        k_ms_max = k_ms
    ms_per_step = total_ms / args.steps
    peak, peak_src = peaks()
    abytes = algorithmic_bytes(n_out, e_local, F, es)
    achieved = abytes / (k_ms * 1e-3) / 1e9
    if rank == 0:
        emit_json({
            "metric": "aggregation edges/s (dst-partitioned gather->scatter_add, RMAT graph)",
            "value": E / (ms_per_step * 1e-3), "unit": "edges/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": args.workload, "nodes": N, "edges": E, "features": F,
                       "edges_rank0": e_local, "rows_rank0": n_out, "plan_build_ms": plan_ms,
                       "distinct_sources_rank0": distinct_src,
                       "max_row_len": plan.max_len, "empty_rows_rank0": plan.n_empty,
                       "exchange": args.exchange if world > 1 else None,
                       "exchange_bytes_in_rank0": ((agg.n_needed - agg.recv_splits[rank]) * F * es
                                                   if (world > 1 and args.exchange in ("needed", "push"))
                                                   else (world - 1) * (N // world) * F * es),
                       "local_kernel_ms_max_over_ranks": k_ms_max,
                       "l2": "inputs larger than L2; no flush needed",
                       "parallelism": f"edge-balanced dst ranges x{world}, equal feature blocks, "
                                      f"{args.exchange} exchange of x"
                                      + (", cyclic feature ownership" if args.cyclic and args.exchange != "allgather" else "")
                                      if world > 1 else "single GPU"},
            "clocks": clocks, "e2e": None, "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                         "kernel": "segreduce_kernel (rank 0 shard)", "kernel_ms": k_ms,
                         "algorithmic_bytes": abytes}})
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit_json(line):
    """The one JSON line goes to the real stdout; everything else (NCCL banners, warnings
    from libraries that print to fd 1) was redirected to stderr in main()."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, data)
    else:
        sys.stdout.write(data.decode())
        sys.stdout.flush()


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="products",
                    choices=sorted(WORKLOADS) + ["rmat26", "rmat24", "rmat22", "rmat20"])
    ap.add_argument("--exchange", default="push", choices=["allgather", "allgather_push", "needed", "push"],
                    help="rmat workloads at N>1: all-gather every feature row; only the rows each rank's "
                         "edges read through an NCCL all-to-all (needed); or the same rows stored directly "
                         "into the peers' buffers by the gather kernel over NVLink (push)")
    ap.add_argument("--cyclic", type=int, default=1,
                    help="rmat workloads: feature row i lives on rank i %% N (balances the serving side "
                         "of the needed-rows exchange); 0 = contiguous equal blocks")
    ap.add_argument("--row-weight", type=int, default=4,
                    help="rmat workloads: cost of one destination row in edge units when balancing ranges")
    ap.add_argument("--stages", type=int, default=0,
                    help="exchange pipeline depth of the products workload at N>1 (default: 4 at N>=4 — "
                         "chunk all-gathers on a high-priority NCCL stream overlap the aggregation of "
                         "the chunks already received, DESIGN.md §5 — and 1 below)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload.startswith("rmat"):
        run_rmat(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
