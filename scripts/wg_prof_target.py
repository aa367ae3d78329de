"""Target for profiling the weight-gradient launch alone: a few eager train passes at B windows (config 2 shape)."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_pkg
from helpers import PAMAP2, seeded_case
pkg = load_pkg()
ops = importlib.import_module(pkg.__name__ + ".ops")
N = importlib.import_module(pkg.__name__ + "._native")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
model, feats, mask, labels = seeded_case(PAMAP2, 256, 4, 25, B, seed=7, device="cuda")
plan = model._plan()
own = dict(model.named_parameters())
arena = plan.gather([own[k].detach() for k, _, _ in plan.slots])
a16 = plan.pack_bf16(arena)
xs = [feats[m].contiguous() for m in plan.names]
ws = torch.empty(plan.workspace_bytes(B, N.MSF_PREC_BF16), dtype=torch.uint8, device="cuda")
kw = dict(precision=N.MSF_PREC_BF16, arena_bf16=a16, workspace=ws)
for _ in range(4):
    ops.fusion_train_pass_raw(plan, arena, xs, mask, labels, smoothing=0.05, training=True, p=0.1, seed=1, **kw)
torch.cuda.synchronize()
