"""In-graph latency of gradient all-reduce variants for the HybridFusion gradient arena (13.4 MB fp32):
NCCL vs torch symmetric-memory two-shot / multimem (NVLS).  torchrun --nproc-per-node N scripts/allreduce_bench.py"""
import os, sys, time
import torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
import torch.distributed._symmetric_memory as sm
gname = dist.group.WORLD.group_name
def timeit(fn, label, n):
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier()
    g = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.graph(g, stream=side):
            for _ in range(10): fn()
    except Exception as e:  # noqa
        if rank == 0: print(f"{label:28s} n={n}: capture failed: {str(e)[:100]}", flush=True)
        return
    for _ in range(3): g.replay()
    torch.cuda.synchronize(); dist.barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(20): g.replay()
    e.record(); torch.cuda.synchronize()
    t = torch.tensor([s.elapsed_time(e) * 1000 / 200], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0: print(f"{label:28s} n={n}: {float(t):7.1f} us", flush=True)
for n in (3363357, 1780000, 262144):
    n4 = (n + 3) // 4 * 4
    x = torch.randn(n4, device=dev)
    timeit(lambda: dist.all_reduce(x), "nccl all_reduce", n4)
    try:
        t = sm.empty(n4, dtype=torch.float32, device=dev); t.copy_(x)
        hdl = sm.rendezvous(t, dist.group.WORLD)
        if rank == 0: print("multicast supported:", getattr(hdl, "multicast_ptr", 0) != 0, flush=True)
        timeit(lambda: torch.ops.symm_mem.two_shot_all_reduce_(t, "sum", gname), "symm two_shot", n4)
        timeit(lambda: torch.ops.symm_mem.one_shot_all_reduce(t, "sum", gname), "symm one_shot", n4)
        if getattr(hdl, "multicast_ptr", 0):
            timeit(lambda: torch.ops.symm_mem.multimem_all_reduce_(t, "sum", gname), "symm multimem (NVLS)", n4)
    except Exception as e:  # noqa
        if rank == 0: print("symmetric memory path failed:", str(e)[:300], flush=True)
torch.cuda.synchronize(); dist.barrier(); sys.stdout.flush(); os._exit(0)
