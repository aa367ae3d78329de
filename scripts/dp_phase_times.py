"""Where the data-parallel step spends its time: the %globaltimer stamps the exchange kernels leave in the signal
block (dp_signals.cuh: SIG_TIME) after graph-replayed bf16 train steps at the benchmark shape, per rank.

    torchrun --nproc-per-node N scripts/dp_phase_times.py
"""
import importlib, os, sys, time
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_pkg
from helpers import PAMAP2, seeded_case
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
engine = importlib.import_module(load_pkg().__name__ + ".engine")
B = 4096
model, feats, mask, labels = seeded_case(PAMAP2, 256, 4, 25, B, seed=31 + rank, device=dev)
torch.manual_seed(0)
model, *_ = seeded_case(PAMAP2, 256, 4, 25, 8, seed=31, device=dev)
eng = engine.FusionEngine(model, B, precision="bf16", seed=9, use_graph=True, comm=os.environ.get("DP_COMM", "auto"))
eng.load_batch(feats, mask, labels)
for _ in range(30):
    eng.train_step_resident()
torch.cuda.synchronize(); dist.barrier()
start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
start.record()
for _ in range(200):
    eng.train_step_resident()
stop.record(); torch.cuda.synchronize()
us = start.elapsed_time(stop) * 1e3 / 200
line = f"rank {rank}/{world} comm={eng.comm} {us:.1f} us/step"
if hasattr(eng, "sig"):
    t = eng.sig[48:56].tolist()
    r0 = t[0]
    names = ["reduce start", "barrier 1 passed", "reduce kernel end", "update start", "barrier 2 passed", "update end", "phase-2 sums pushed"]
    order = [0, 1, 6, 2, 3, 4, 5]
    line += " | " + "  ".join(f"{names[i]} {((t[i] - r0) / 1e3):.1f}" for i in order)
print(line, flush=True)
dist.barrier(); os._exit(0)
