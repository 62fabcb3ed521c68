"""2-GPU microbenchmark of the push exchange: torchrun --nproc-per-node 2 profiles/push_micro.py"""
import ctypes, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gnn-ops-benchmark_b200"))
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm
from gno_b200._lib import lib, check
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"])); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
F, es = 128, 2
rows = 10_000_000  # 2.56 GB
x = torch.randn(rows, F, device=dev).to(torch.bfloat16)
buf = symm.empty((rows, F), dtype=torch.bfloat16, device=dev)
hdl = symm.rendezvous(buf, dist.group.WORLD)
peer = hdl.get_buffer((rank + 1) % world, (rows, F), torch.bfloat16)
def timed(fn, n=5):
    fn(); hdl.barrier(channel=0); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
t = timed(lambda: peer.copy_(x))
if rank == 0: print("peer copy_ (memcpy) GB/s", rows * F * es / t / 1e6, flush=True)
for order in ("sequential", "random"):
    serve = torch.arange(rows, device=dev) if order == "sequential" else torch.randperm(rows, device=dev)
    P = world
    n_other = rows
    seg = [0] * (P + 1)
    for q in range(P):
        seg[q + 1] = seg[q] + (n_other if q == (rank + 1) % world else 0)
    ptrs = (ctypes.c_void_p * P)(*[int(hdl.buffer_ptrs[q]) for q in range(P)])
    segc = (ctypes.c_int64 * (P + 1))(*seg); off = (ctypes.c_int64 * P)(*([0] * P))
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    for max_blocks in (0, 1184, 592, 296, 148, 74):   # grid cap: how few CTAs still saturate NVLink?
        def push():
            check(lib.gno_push_rows(ctypes.c_void_p(x.data_ptr()), F * es, F * es, ctypes.c_void_p(serve.data_ptr()), rows, P,
                                    ptrs, segc, off, F * es, 0, max_blocks, st))
        t = timed(push)
        if rank == 0: print("push kernel", order, "max_blocks", max_blocks, "GB/s", rows * F * es / t / 1e6, flush=True)
    def push_bar():
        hdl.barrier(channel=0); push(); hdl.barrier(channel=1)
    t = timed(push_bar)
    if rank == 0: print("barrier+push+barrier", order, "GB/s", rows * F * es / t / 1e6, "ms", t, flush=True)
out = torch.empty(world * rows // 4, F, device=dev, dtype=torch.bfloat16)
t = timed(lambda: dist.all_gather_into_tensor(out, x[:rows // 4]))
if rank == 0: print("nccl all_gather inbound GB/s", (world - 1) * (rows // 4) * F * es / t / 1e6, flush=True)
dist.destroy_process_group()
