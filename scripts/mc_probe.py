"""Is NVLink multicast (NVLS: multimem.* on a multicast address) available to torch symmetric memory on this box?"""
import os, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
t = symm_mem.empty(1 << 20, dtype=torch.float32, device=dev)
h = symm_mem.rendezvous(t, dist.group.WORLD)
mc = getattr(h, "multicast_ptr", None)
print(f"rank {rank}: multicast_ptr={mc} has_multicast_support={getattr(symm_mem, 'has_multicast_support', lambda *a: 'n/a')('cuda', local) if hasattr(symm_mem, 'has_multicast_support') else 'n/a'} buffer_ptrs={[hex(p) for p in h.buffer_ptrs]}", flush=True)
if mc:
    t.fill_(rank + 1.0)
    torch.cuda.synchronize(); dist.barrier()
    try:
        out = torch.empty_like(t)
        torch.ops.symm_mem.multimem_all_reduce_(t, "sum", dist.group.WORLD.group_name)
        torch.cuda.synchronize()
        print(f"rank {rank}: multimem_all_reduce_ -> {float(t[0])} (expect {world * (world + 1) / 2})", flush=True)
    except Exception as e:
        print(f"rank {rank}: multimem_all_reduce_ failed: {e}", flush=True)
dist.barrier(); os._exit(0)
