"""In-graph time of the data-parallel optimizer step alone: peer-memory kernels vs NCCL all-reduce + optimizer."""
import ctypes, importlib, os, sys, time
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_pkg
from helpers import PAMAP2, seeded_case
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
pkg = load_pkg(); N = pkg.native; lib = pkg.lib()
engine = importlib.import_module(pkg.__name__ + ".engine")
for comm in ("p2p", "nccl"):
    model, feats, mask, labels = seeded_case(PAMAP2, 256, 4, 25, 64, seed=31, device=dev)
    eng = engine.FusionEngine(model, 64, precision="bf16", seed=9, use_graph=False, comm=comm)
    eng.grad.normal_(0, 1e-3)
    def fn():
        st = pkg.ops._stream() if hasattr(pkg, "ops") else importlib.import_module(pkg.__name__ + ".ops")._stream()
        if comm == "p2p":
            N.check(lib.msf_dp_optimizer_step(ctypes.byref(eng.plan.shape), ctypes.byref(eng.dp_comm), eng.arena.data_ptr(),
                    eng.exp_avg.data_ptr(), eng.exp_avg_sq.data_ptr(), eng.state.data_ptr(), 1e-3, 0.9, 0.999, 1e-8, 1e-4, 1.0, 1.0, st))
        else:
            eng._nccl_step(lib, st)
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        for _ in range(10): fn()
    for _ in range(3): g.replay()
    torch.cuda.synchronize(); dist.barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(20): g.replay()
    e.record(); torch.cuda.synchronize()
    t = torch.tensor([s.elapsed_time(e) * 1000 / 200], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0: print(f"{comm}: {float(t):.1f} us per optimizer step (world {world})", flush=True)
    if comm == "p2p":
        ts = eng.sig[48:56].tolist()
        print(f"rank {rank} stamps (us rel): " + " ".join(f"{(x - ts[0]) / 1e3:.1f}" for x in ts), flush=True)
torch.cuda.synchronize(); dist.barrier(); sys.stdout.flush(); os._exit(0)
