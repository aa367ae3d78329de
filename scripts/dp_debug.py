"""Phase-by-phase run of the peer-memory data-parallel step (debug aid).  torchrun --nproc-per-node 2 scripts/dp_debug.py"""
import faulthandler, importlib, os, sys, time
faulthandler.dump_traceback_later(45, exit=True)
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_pkg
from helpers import PAMAP2, seeded_case
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
t0 = time.perf_counter()
def mark(s): print(f"[rank {rank} +{time.perf_counter()-t0:5.1f}s] {s}", file=sys.stderr, flush=True)
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
mark("pg ready")
engine = importlib.import_module(load_pkg().__name__ + ".engine")
B = 512
model, feats, mask, labels = seeded_case(PAMAP2, 256, 4, 25, B, seed=31, device=dev)
sl = engine.shard_batch(B, rank, world)
use_graph = os.environ.get("DP_GRAPH", "0") == "1"
eng = engine.FusionEngine(model, B // world, precision="fp32", seed=9, use_graph=use_graph, comm=os.environ.get("DP_COMM", "p2p"))
mark(f"engine built comm={eng.comm}")
eng.p = 0.0
shard = ({k: v[sl] for k, v in feats.items()}, mask[sl], labels[sl])
for i in range(3):
    loss = eng.train_step(*shard)
    torch.cuda.synchronize()
    mark(f"step {i} loss {float(loss):.6f} sig {eng.sig[:20].tolist() if hasattr(eng, 'sig') else None}")
mark(f"arena checksum {float(eng.arena.double().sum()):.9f}")
dist.barrier(); os._exit(0)
