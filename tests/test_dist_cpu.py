"""world_size-2 gloo tests (CPU) of the partitioned-aggregation host logic: edge-balanced
ranges, source remapping into the padded gather buffer, the all-gather exchange.  The local
reduction is done by the oracle here (no GPU); the CUDA path is covered by test_gpu_dist.py."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _graph(seed, N, E, F):
    g = torch.Generator().manual_seed(seed)
    dst = (torch.rand(E, generator=g) ** 3 * N).long().clamp_(0, N - 1)  # skewed in-degrees
    src = torch.randint(0, N, (E,), generator=g)
    x = (torch.randn(N, F, generator=g) * 4).round() / 4
    return src, dst, x


def _worker(rank, world, port, N, E, F, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from gno_b200.dist import DistAggregator, partition_graph
        src, dst, x = _graph(3, N, E, F)
        bounds, shards = partition_graph(src, dst, N, world)
        s_r, d_r = shards[rank]
        agg = DistAggregator(bounds, s_r, d_r)
        lo, hi = int(bounds[rank]), int(bounds[rank + 1])
        x_full = agg.exchange(x[lo:hi].contiguous())
        ok = True
        for red in ("sum", "max", "mean"):
            got, garg = oracle.gather_scatter(x_full, agg.src_padded, d_r, hi - lo, red)
            want, warg = oracle.gather_scatter(x, src, dst, N, red)
            ok &= torch.equal(got, want[lo:hi]) if red == "max" else torch.allclose(got, want[lo:hi], rtol=1e-5, atol=1e-5)
            if red == "max":
                # arg is a position in the LOCAL shard: map back to the global edge id
                local_to_global = torch.nonzero((dst >= lo) & (dst < hi)).flatten()
                sentinel = garg == s_r.numel()
                mapped = torch.where(sentinel, torch.full_like(garg, E), local_to_global[garg.clamp(max=max(s_r.numel() - 1, 0))])
                ok &= torch.equal(mapped, warg[lo:hi])
        # K-stage pipelined layout: per-stage chunk all-gathers + per-stage edge shards sum to the same
        agg3 = DistAggregator(bounds, s_r, d_r, stages=3)
        total = torch.zeros(hi - lo, F)
        for (buf, work), (ids, d_s) in zip(agg3.exchange_stages(x[lo:hi].contiguous()), agg3.stage_edges):
            if work is not None:
                work.wait()
            total += oracle.gather_scatter(buf, ids, d_s, hi - lo, "sum")[0]
        want, _ = oracle.gather_scatter(x, src, dst, N, "sum")
        ok &= torch.allclose(total, want[lo:hi], rtol=1e-5, atol=1e-4)
        ok &= sum(e[0].numel() for e in agg3.stage_edges) == s_r.numel()
        # needed-rows-only exchange: unequal feature blocks, only referenced rows travel
        xb = torch.tensor([0, 100, N])
        aggn = DistAggregator(bounds, s_r, d_r, feature_bounds=xb, exchange="needed")
        xlo, xhi = int(xb[rank]), int(xb[rank + 1])
        recv = aggn.exchange_needed(x[xlo:xhi].contiguous(), gather_rows=lambda t, r: t.index_select(0, r))
        ok &= recv.size(0) == torch.unique(s_r).numel()
        got, _ = oracle.gather_scatter(recv, aggn.src_needed, d_r, hi - lo, "sum")
        ok &= torch.allclose(got, want[lo:hi], rtol=1e-5, atol=1e-4)
        # cyclic feature ownership (row i on rank i % P): balanced serving of hub rows
        aggc = DistAggregator(bounds, s_r, d_r, exchange="needed", cyclic_rows=N)
        recv = aggc.exchange_needed(x[rank::world].contiguous(), gather_rows=lambda t, r: t.index_select(0, r))
        got, _ = oracle.gather_scatter(recv, aggc.src_needed, d_r, hi - lo, "sum")
        ok &= torch.allclose(got, want[lo:hi], rtol=1e-5, atol=1e-4)
        # K-stage needed-rows exchange pipelined over destination sub-ranges: stage s delivers
        # exactly the rows sub-range s still misses (stage-major receive layout)
        for fr in (None, [0.1, 0.3, 0.6]):
            aggs = DistAggregator(bounds, s_r, d_r, exchange="needed", cyclic_rows=N, stages=3, stage_fracs=fr,
                                  row_weight=2)
            recv = aggs.exchange_needed(x[rank::world].contiguous(), gather_rows=lambda t, r: t.index_select(0, r))
            ok &= recv.size(0) == torch.unique(s_r).numel()
            sb = aggs.sub_bounds
            ok &= int(sb[0]) == 0 and int(sb[-1]) == hi - lo
            got = torch.zeros(hi - lo, F)
            seen = 0
            for s in range(3):
                m = aggs.stage_of_edge == s
                a, b = int(sb[s]), int(sb[s + 1])
                ok &= bool(((d_r[m] >= a) & (d_r[m] < b)).all())
                seen += int(m.sum())
                if m.any():  # every source of stage s has landed once stages 0..s are delivered
                    ok &= int(aggs.src_needed[m].max()) < aggs.stage_row0[s + 1]
                got[a:b] = oracle.gather_scatter(recv, aggs.src_needed[m], d_r[m] - a, b - a, "sum")[0]
            ok &= seen == s_r.numel()
            ok &= torch.allclose(got, want[lo:hi], rtol=1e-5, atol=1e-4)
            # the offsets the owners were told (put_off) are the receivers' own block starts
            mine = torch.tensor(aggs.recv_off)            # [stage][owner] on this rank
            allr = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(allr, mine)
            for q in range(world):
                ok &= [allr[q][s][rank].item() for s in range(3)] == [aggs.put_off[s][q] for s in range(3)]
        # source split: stage 0 = own rows, then remote rows by decreasing reference count; a stage
        # reduces the edges whose source lies in its group, accumulating over all rows
        for K, fr in ((2, None), (4, [0.1, 0.3, 0.6])):
            aggs = DistAggregator(bounds, s_r, d_r, exchange="needed", cyclic_rows=N, stages=K, stage_fracs=fr,
                                  split="source")
            recv = aggs.exchange_needed(x[rank::world].contiguous(), gather_rows=lambda t, r: t.index_select(0, r))
            ok &= recv.size(0) == torch.unique(s_r).numel()
            ok &= aggs.recv_cnt[0][1 - rank] == 0 and aggs.serve_cnt[0][1 - rank] == 0   # stage 0 stays on this rank
            ok &= all(aggs.recv_cnt[s][rank] == 0 for s in range(1, K))                   # remote stages hold no own row
            got = torch.zeros(hi - lo, F)
            refs_prev = None
            for s in range(K):
                m = aggs.stage_of_edge == s
                if m.any():
                    ok &= int(aggs.src_needed[m].min()) >= aggs.stage_row0[s]
                    ok &= int(aggs.src_needed[m].max()) < aggs.stage_row0[s + 1]
                    got += oracle.gather_scatter(recv, aggs.src_needed[m], d_r[m], hi - lo, "sum")[0]
                    if s >= 1:  # groups are ordered by decreasing reference count
                        refs = torch.bincount(aggs.src_needed[m] - aggs.stage_row0[s])
                        refs = refs[refs > 0]
                        if refs_prev is not None and refs.numel():
                            ok &= int(refs.max()) <= refs_prev
                        if refs.numel():
                            refs_prev = int(refs.min())
            ok &= torch.allclose(got, want[lo:hi], rtol=1e-5, atol=1e-4)
        # hybrid split: stage 0 = own-source edges over all rows; the remote-source edges by destination
        # sub-range, each stage after the rows it reads first have arrived
        aggh = DistAggregator(bounds, s_r, d_r, exchange="needed", cyclic_rows=N, stages=4, stage_fracs=[0.2, 0.3, 0.5],
                              split="hybrid", row_weight=2)
        recv = aggh.exchange_needed(x[rank::world].contiguous(), gather_rows=lambda t, r: t.index_select(0, r))
        own_edge = (s_r % world) == rank
        ok &= bool((aggh.stage_of_edge[own_edge] == 0).all()) and bool((aggh.stage_of_edge[~own_edge] >= 1).all())
        ok &= aggh.recv_cnt[0][1 - rank] == 0
        got = torch.zeros(hi - lo, F)
        for s in range(4):
            m = aggh.stage_of_edge == s
            if s >= 1:
                a, b = int(aggh.sub_bounds[s - 1]), int(aggh.sub_bounds[s])
                ok &= bool(((d_r[m] >= a) & (d_r[m] < b)).all())
                if m.any():
                    ok &= int(aggh.src_needed[m].max()) < aggh.stage_row0[s + 1]
                    ok &= int(aggh.src_needed[m].min()) >= aggh.stage_row0[1]
            got += oracle.gather_scatter(recv, aggh.src_needed[m], d_r[m], hi - lo, "sum")[0]
        ok &= torch.allclose(got, want[lo:hi], rtol=1e-5, atol=1e-4)
        # xorfold ownership (balanced on ids with skewed bits): pad the feature rows to an even count
        from gno_b200.dist import xorfold_global_ids
        xp = torch.cat([x, torch.zeros(N % 2, F)])
        aggx = DistAggregator(bounds, s_r, d_r, exchange="needed", cyclic_rows=xp.size(0), ownership="xorfold",
                              stages=3, stage_fracs=[0.5, 0.5], split="source")
        mine = xorfold_global_ids(rank, world, xp.size(0) // world)
        recv = aggx.exchange_needed(xp[mine].contiguous(), gather_rows=lambda t, r: t.index_select(0, r))
        got, _ = oracle.gather_scatter(recv, aggx.src_needed, d_r, hi - lo, "sum")
        ok &= torch.allclose(got, want[lo:hi], rtol=1e-5, atol=1e-4)
        # merge_own (source split): the own-source edges and the first remote group become ONE stage that
        # gathers from two buffers — ids below n_local read x_local, the rest read the receive buffer
        aggm = DistAggregator(bounds, s_r, d_r, exchange="needed", cyclic_rows=xp.size(0), ownership="xorfold",
                              stages=3, stage_fracs=[0.3, 0.7], split="source", merge_own=True)
        ok &= aggm.merge_own
        x_loc = xp[mine].contiguous()
        recv = aggm.exchange_needed(x_loc, gather_rows=lambda t, r: t.index_select(0, r))
        both = torch.cat([x_loc, recv])                                  # what the two-base gather addresses
        first = aggm.stage_of_edge <= 1
        own_e = aggm.stage_of_edge[first] == 0
        ids = torch.where(own_e, aggm.src_local[first], aggm.src_needed[first] + aggm.n_local)
        ok &= bool((ids[own_e] < aggm.n_local).all()) and bool((ids[~own_e] >= aggm.n_local).all())
        got = oracle.gather_scatter(both, ids, d_r[first], hi - lo, "sum")[0]
        rest = ~first
        got += oracle.gather_scatter(recv, aggm.src_needed[rest], d_r[rest], hi - lo, "sum")[0]
        ok &= torch.allclose(got, want[lo:hi], rtol=1e-5, atol=1e-4)
        # mixed<D>: the receiver picks its own structure from its average row degree
        for thr, expect in ((1e9, "hybrid"), (0.0, "source")):
            aggq = DistAggregator(bounds, s_r, d_r, exchange="needed", cyclic_rows=N, stages=3, stage_fracs=[0.4, 0.6],
                                  split=f"mixed{thr:g}")
            ok &= aggq.split == expect
            recv = aggq.exchange_needed(x[rank::world].contiguous(), gather_rows=lambda t, r: t.index_select(0, r))
            got, _ = oracle.gather_scatter(recv, aggq.src_needed, d_r, hi - lo, "sum")
            ok &= torch.allclose(got, want[lo:hi], rtol=1e-5, atol=1e-4)
        ret[rank] = (bool(ok), int(d_r.numel()))
    finally:
        dist.destroy_process_group()


def test_partitioned_aggregation_gloo_world2():
    world, N, E, F = 2, 301, 5000, 6
    with mp.Manager() as m:
        ret = m.dict()
        mp.spawn(_worker, args=(world, _free_port(), N, E, F, ret), nprocs=world, join=True)
        assert all(ret[r][0] for r in range(world)), dict(ret)
        edges = [ret[r][1] for r in range(world)]
        assert sum(edges) == E
        assert max(edges) < 0.65 * E, f"ranges are not edge-balanced: {edges}"


def test_edge_balanced_ranges():
    from gno_b200.dist import edge_balanced_ranges
    counts = torch.tensor([100, 1, 1, 1, 1, 96, 50, 50])
    b = edge_balanced_ranges(counts, 3)
    assert b[0] == 0 and b[-1] == 8 and (b[1:] >= b[:-1]).all()
    per = [int(counts[b[i]:b[i + 1]].sum()) for i in range(3)]
    assert sum(per) == 300 and max(per) <= 200
    assert edge_balanced_ranges(torch.zeros(0, dtype=torch.int64), 4).tolist() == [0, 0, 0, 0, 0]
    assert edge_balanced_ranges(torch.tensor([5]), 2).tolist()[-1] == 1


def test_xorfold_ownership_is_a_balanced_bijection():
    from gno_b200.dist import xorfold_global_ids, xorfold_owner
    for world in (1, 2, 4, 8):
        n = 1 << 10
        ids = torch.arange(n)
        owner, local = xorfold_owner(ids, world)
        assert int(owner.min()) >= 0 and int(owner.max()) < world
        assert torch.bincount(owner, minlength=world).tolist() == [n // world] * world
        assert torch.unique(owner * (n // world) + local).numel() == n          # (owner, local) is a bijection
        for q in range(world):
            g = xorfold_global_ids(q, world, n // world)
            o2, l2 = xorfold_owner(g, world)
            assert bool((o2 == q).all()) and torch.equal(l2, torch.arange(n // world))
    # ids whose bits are 1 with probability 0.24 (R-MAT): cyclic ownership is skewed, xorfold is not
    g = torch.Generator().manual_seed(0)
    bits = (torch.rand(200_000, 20, generator=g) < 0.24).long()
    ids = (bits << torch.arange(20)).sum(1)
    cyc = torch.bincount(ids % 2, minlength=2).double() / ids.numel()
    xf = torch.bincount(xorfold_owner(ids, 2)[0], minlength=2).double() / ids.numel()
    assert cyc.max() > 0.7 and xf.max() < 0.52
