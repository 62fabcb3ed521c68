"""Host-buffer entry points: the end-to-end call a user with CPU-resident data makes.

`gather_scatter_host` takes pinned (or pageable) host tensors, moves them to the GPU, builds
the plan, aggregates and returns the result in host memory.  Copies and compute are overlapped
where the data flow allows it:

    copy stream : H2D edge_index ─┬─ H2D x ───────────────┐
    main stream :                 └─ plan build (sort) ───┴─ aggregate ── D2H out

`HostPipeline` keeps two such calls in flight so the D2H of result i overlaps the H2D of
inputs i+1 (PCIe is full duplex): the throughput form used by bench.py's e2e leg.
"""
import torch

from . import ops
from .plan import build_plan


class _Slot:
    def __init__(self, dev):
        self.copy = torch.cuda.Stream(device=dev)
        self.main = torch.cuda.Stream(device=dev)
        self.ei_ready = torch.cuda.Event()
        self.x_ready = torch.cuda.Event()
        self.done = torch.cuda.Event()
        self.keep = None


def _launch(slot, x_host, edge_index_host, num_rows, reduce, out_host, dev):
    """Enqueue one host→device→host aggregation on the slot's streams; returns out_host."""
    with torch.cuda.stream(slot.copy):
        ei = edge_index_host.to(dev, non_blocking=True)
        slot.ei_ready.record(slot.copy)
        x = x_host.to(dev, non_blocking=True)
        slot.x_ready.record(slot.copy)
    with torch.cuda.stream(slot.main):
        slot.main.wait_event(slot.ei_ready)
        plan = build_plan(ei[1], num_rows)          # overlaps the H2D of x
        gidx = plan.sorted_ids(ei[0])
        slot.main.wait_event(slot.x_ready)
        out = ops.segment_reduce(plan, x, reduce, gidx=gidx)
        if out_host is None:
            out_host = torch.empty(out.shape, dtype=out.dtype, pin_memory=True)
        out_host.copy_(out, non_blocking=True)
        slot.done.record(slot.main)
    # tensors allocated on side streams must outlive the work enqueued on them
    for t in (ei, x, out, gidx, plan.perm, plan.erow, plan.rowptr):
        t.record_stream(slot.main)
    slot.keep = (ei, x, out, plan, gidx)
    return out_host


def gather_scatter_host(x_host, edge_index_host, num_rows, reduce="sum", out_host=None, device=None):
    """scatter(x.index_select(0, edge_index[0]), edge_index[1], 0, num_rows, reduce) for HOST
    tensors; returns the result in (pinned) host memory after synchronising."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    slot = _default_slots.get(dev)
    if slot is None:  # streams are kept so the caching allocator can reuse their blocks
        slot = _default_slots[dev] = _Slot(dev)
    cur = torch.cuda.current_stream(dev)
    slot.copy.wait_stream(cur)
    slot.main.wait_stream(cur)
    out_host = _launch(slot, x_host, edge_index_host, num_rows, reduce, out_host, dev)
    slot.done.synchronize()
    slot.keep = None
    return out_host


_default_slots = {}


class HostPipeline:
    """Two aggregations in flight: submit() returns immediately, result() waits for the oldest."""

    def __init__(self, num_rows, reduce="sum", device=None, depth=2):
        self.dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.num_rows, self.reduce = num_rows, reduce
        self.slots = [_Slot(self.dev) for _ in range(depth)]
        self.pending = []
        self.i = 0

    def submit(self, x_host, edge_index_host, out_host=None):
        if len(self.pending) == len(self.slots):
            raise RuntimeError("pipeline full: call result() first")
        slot = self.slots[self.i % len(self.slots)]
        self.i += 1
        out = _launch(slot, x_host, edge_index_host, self.num_rows, self.reduce, out_host, self.dev)
        self.pending.append((slot, out))

    def result(self):
        slot, out = self.pending.pop(0)
        slot.done.synchronize()
        slot.keep = None
        return out
