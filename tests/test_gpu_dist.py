"""Multi-GPU parity of the partitioned aggregation (NCCL all-gather + local segment reduce).
Needs >= 2 CUDA devices (gpurun --gpus 2); skipped otherwise."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from gno_b200.dist import DistAggregator, partition_graph
        g = torch.Generator().manual_seed(5)
        N, E, F = 3001, 200_000, 100
        dst = (torch.rand(E, generator=g) ** 3 * N).long().clamp_(0, N - 1)
        src = torch.randint(0, N, (E,), generator=g)
        x = (torch.randn(N, F, generator=g) * 4).round() / 4
        bounds, shards = partition_graph(src, dst, N, world)
        lo, hi = int(bounds[rank]), int(bounds[rank + 1])
        agg = DistAggregator(bounds, shards[rank][0].to(dev), shards[rank][1].to(dev))
        ok = True
        for red in ("sum", "max"):
            got = agg.aggregate(x[lo:hi].to(dev), red, return_arg=True)
            want, warg = oracle.gather_scatter(x, src, dst, N, red)
            if red == "max":
                ok &= torch.equal(got[0].cpu(), want[lo:hi])
            else:
                ok &= torch.allclose(got.cpu(), want[lo:hi], rtol=1e-5, atol=1e-3)
        # pipelined exchange (3 stages): same sums, NVLink overlapped with the gather-reduce
        agg3 = DistAggregator(bounds, shards[rank][0].to(dev), shards[rank][1].to(dev), stages=3)
        for red in ("sum", "mean"):
            got = agg3.aggregate(x[lo:hi].to(dev), red)
            want, _ = oracle.gather_scatter(x, src, dst, N, red)
            ok &= torch.allclose(got.cpu(), want[lo:hi], rtol=1e-5, atol=1e-3)
        # needed-rows-only exchange with equal feature blocks
        xb = torch.tensor([0, N // 2, N])
        aggn = DistAggregator(bounds, shards[rank][0].to(dev), shards[rank][1].to(dev), feature_bounds=xb,
                              exchange="needed")
        got = aggn.aggregate(x[int(xb[rank]):int(xb[rank + 1])].to(dev), "max", return_arg=True)
        want, _ = oracle.gather_scatter(x, src, dst, N, "max")
        ok &= torch.equal(got[0].cpu(), want[lo:hi])
        # the same exchange done by the fused gather + NVLink peer-store kernel (symmetric memory)
        aggp = DistAggregator(bounds, shards[rank][0].to(dev), shards[rank][1].to(dev), feature_bounds=xb,
                              exchange="push")
        for _ in range(3):  # repeated calls reuse the receive buffer: exercises both barriers
            got = aggp.aggregate(x[int(xb[rank]):int(xb[rank + 1])].to(dev), "max", return_arg=True)
            ok &= torch.equal(got[0].cpu(), want[lo:hi])
        agga = DistAggregator(bounds, shards[rank][0].to(dev), shards[rank][1].to(dev), exchange="allgather_push")
        for _ in range(2):
            got = agga.aggregate(x[lo:hi].to(dev), "max", return_arg=True)
            ok &= torch.equal(got[0].cpu(), want[lo:hi])
        aggc = DistAggregator(bounds, shards[rank][0].to(dev), shards[rank][1].to(dev), exchange="push",
                              cyclic_rows=N)
        got = aggc.aggregate(x[rank::world].contiguous().to(dev), "max", return_arg=True)
        ok &= torch.equal(got[0].cpu(), want[lo:hi])
        # K-stage exchange pipelined over destination sub-ranges, overlapped with the reduction on a
        # second stream: peer-store kernel (push) and NCCL all-to-all (needed)
        wsum, _ = oracle.gather_scatter(x, src, dst, N, "sum")
        l2g = torch.nonzero((dst >= lo) & (dst < hi)).flatten()
        for mode, K, fr in (("push", 3, [0.2, 0.3, 0.5]), ("push", 4, None), ("needed", 2, None)):
            aggs = DistAggregator(bounds, shards[rank][0].to(dev), shards[rank][1].to(dev), exchange=mode,
                                  cyclic_rows=N, stages=K, stage_fracs=fr, row_weight=4)
            xl = x[rank::world].contiguous().to(dev)
            for _ in range(3):
                got, garg = aggs.aggregate(xl, "max", return_arg=True)
                ok &= torch.equal(got.cpu(), want[lo:hi])
                garg = garg.cpu()
                sent = garg == shards[rank][0].numel()
                mapped = torch.where(sent, torch.full_like(garg, E), l2g[garg.clamp(max=max(l2g.numel() - 1, 0))])
                ok &= torch.equal(mapped, warg[lo:hi])
                got = aggs.aggregate(xl, "sum")
                ok &= torch.allclose(got.cpu(), wsum[lo:hi], rtol=1e-5, atol=1e-3)
        # source split (own rows first, then remote rows by decreasing reference count; stages accumulate)
        wmean, _ = oracle.gather_scatter(x, src, dst, N, "mean")
        for mode, K, fr, split in (("push", 2, None, "source"), ("push", 4, [0.1, 0.3, 0.6], "source"),
                                   ("needed", 3, [0.2, 0.8], "source"), ("push", 4, [0.2, 0.3, 0.5], "hybrid"),
                                   ("push", 2, None, "hybrid"), ("needed", 3, None, "hybrid")):
            aggs = DistAggregator(bounds, shards[rank][0].to(dev), shards[rank][1].to(dev), exchange=mode,
                                  cyclic_rows=N, stages=K, stage_fracs=fr, split=split, row_weight=2)
            xl = x[rank::world].contiguous().to(dev)
            for _ in range(3):
                ok &= torch.allclose(aggs.aggregate(xl, "sum").cpu(), wsum[lo:hi], rtol=1e-5, atol=1e-3)
                ok &= torch.allclose(aggs.aggregate(xl, "mean").cpu(), wmean[lo:hi], rtol=1e-5, atol=1e-3)
        # xorfold ownership + a capped push grid (the exchange shares the SMs with the reduction)
        from gno_b200.dist import xorfold_global_ids
        xp = torch.cat([x, torch.zeros(N % 2, F)])
        aggx = DistAggregator(bounds, shards[rank][0].to(dev), shards[rank][1].to(dev), exchange="push",
                              cyclic_rows=xp.size(0), ownership="xorfold", stages=3, stage_fracs=[0.3, 0.7],
                              split="source", push_blocks=8)
        xl = xp[xorfold_global_ids(rank, world, xp.size(0) // world)].contiguous().to(dev)
        for _ in range(3):
            ok &= torch.allclose(aggx.aggregate(xl, "sum").cpu(), wsum[lo:hi], rtol=1e-5, atol=1e-3)
        # merge_own: own-source edges and the first remote group in one two-buffer launch
        for K, fr in ((3, [0.3, 0.7]), (4, [0.2, 0.3, 0.5])):
            aggm = DistAggregator(bounds, shards[rank][0].to(dev), shards[rank][1].to(dev), exchange="push",
                                  cyclic_rows=xp.size(0), ownership="xorfold", stages=K, stage_fracs=fr,
                                  split="source", merge_own=True)
            for _ in range(3):
                ok &= torch.allclose(aggm.aggregate(xl, "sum").cpu(), wsum[lo:hi], rtol=1e-5, atol=1e-3)
                ok &= torch.allclose(aggm.aggregate(xl, "mean").cpu(), wmean[lo:hi], rtol=1e-5, atol=1e-3)
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_partitioned_aggregation_nccl():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 CUDA devices")
    world = 2
    with mp.Manager() as m:
        ret = m.dict()
        mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
        assert all(ret[r] for r in range(world)), dict(ret)


def _worker_world1(rank, world, port, ret):
    """The staged exchange on ONE GPU (a 1-rank NCCL group): every code path of the K-stage push /
    needed pipeline — stage plans, stage-major receive layout, the peer-store kernel writing into
    this rank's own symmetric buffer, the second stream and its events — without a second device."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)
    try:
        from gno_b200.dist import DistAggregator
        g = torch.Generator().manual_seed(9)
        N, E, F = 4001, 300_000, 128
        dst = (torch.rand(E, generator=g) ** 3 * N).long().clamp_(0, N - 1)
        src = (torch.rand(E, generator=g) ** 2 * N).long().clamp_(0, N - 1)
        x = ((torch.randn(N, F, generator=g) * 4).round() / 4).to(torch.bfloat16)
        bounds = torch.tensor([0, N])
        wsum, _ = oracle.gather_scatter(x.float(), src, dst, N, "sum")
        wmax, warg = oracle.gather_scatter(x.float(), src, dst, N, "max")
        ok = True
        for mode, K, fr in (("push", 4, [0.1, 0.2, 0.3, 0.4]), ("push", 2, None), ("needed", 3, None), ("push", 1, None)):
            agg = DistAggregator(bounds, src.to(dev), dst.to(dev), rank=0, world=1, exchange=mode, cyclic_rows=N,
                                 stages=K, stage_fracs=fr, row_weight=4)
            for _ in range(2):
                got = agg.aggregate(x.to(dev), "sum")
                ok &= torch.allclose(got.float().cpu(), wsum, rtol=1e-2, atol=1e-2)
                gm, ga = agg.aggregate(x.to(dev), "max", return_arg=True)
                ok &= torch.equal(gm.float().cpu(), wmax)
                ok &= torch.equal(ga.cpu(), warg)
        # source split on one rank: every row is this rank's own, the remote stages are empty
        for split in ("source", "hybrid"):
            agg = DistAggregator(bounds, src.to(dev), dst.to(dev), rank=0, world=1, exchange="push", cyclic_rows=N,
                                 stages=3, stage_fracs=[0.1, 0.9], split=split)
            for _ in range(2):
                ok &= torch.allclose(agg.aggregate(x.to(dev), "sum").float().cpu(), wsum, rtol=1e-2, atol=1e-2)
        agg = DistAggregator(bounds, src.to(dev), dst.to(dev), rank=0, world=1, exchange="push", cyclic_rows=N,
                             stages=3, stage_fracs=[0.1, 0.9], split="source", merge_own=True)
        for _ in range(2):
            ok &= torch.allclose(agg.aggregate(x.to(dev), "sum").float().cpu(), wsum, rtol=1e-2, atol=1e-2)
        ret[0] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_staged_exchange_single_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    with mp.Manager() as m:
        ret = m.dict()
        mp.spawn(_worker_world1, args=(1, _free_port(), ret), nprocs=1, join=True)
        assert ret[0]
