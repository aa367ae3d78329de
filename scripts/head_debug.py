import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_pkg, Golden
from helpers import PAMAP2, seeded_case, module_from_golden
pkg = load_pkg()
ops = importlib.import_module(pkg.__name__ + ".ops")
N = importlib.import_module(pkg.__name__ + "._native")
B = 4096
model, feats, mask, labels = seeded_case(PAMAP2, 256, 4, 25, B, seed=7, device="cuda")
plan = model._plan()
own = dict(model.named_parameters())
arena = plan.gather([own[k].detach() for k, _, _ in plan.slots])
a16 = plan.pack_bf16(arena)
xs = [feats[m].contiguous() for m in plan.names]
kw = dict(precision=N.MSF_PREC_BF16, training=True, p=0.1, seed=11, offset=3, arena_bf16=a16)
def fused():
    os.environ.pop("MSF_NO_HEAD", None)
    return ops.fusion_forward_raw(plan, arena, xs, mask, **kw)
def unfused():
    os.environ["MSF_NO_HEAD"] = "1"
    r = ops.fusion_forward_raw(plan, arena, xs, mask, **kw)
    os.environ.pop("MSF_NO_HEAD")
    return r
a, b = fused(), fused()
c, d = unfused(), unfused()
torch.cuda.synchronize()
print("fused vs fused    logits", float((a[0] - b[0]).abs().max()), "fw", float((a[1] - b[1]).abs().max()))
print("unfused vs unfused logits", float((c[0] - d[0]).abs().max()), "fw", float((c[1] - d[1]).abs().max()))
diff = (a[1] - c[1]).abs()
print("fused vs unfused  logits", float((a[0] - c[0]).abs().max()), "fw", float(diff.max()))
rows = torch.nonzero(diff.max(1).values > 2e-6).flatten()
print("rows with fw diff:", rows.numel(), rows[:20].tolist())
for r in rows[:6].tolist():
    print(r, mask[r].tolist(), a[1][r].tolist(), c[1][r].tolist())

# packed optimizer
g = Golden("fusion_pamap_small.npz")
plan2 = module_from_golden(g)._plan()
keys = [k for k, _, _ in plan2.slots]
flat = lambda grp: torch.cat([g.t(f"{grp}/{k}").flatten() for k in keys]).cuda()
pa, gr = flat("sd"), flat("grad")
pb = pa.clone()
ma, va, mb, vb = (torch.zeros_like(pa) for _ in range(4))
sa = torch.tensor([7, 3, 1], dtype=torch.int64, device="cuda"); sb = sa.clone()
b16 = plan2.pack_bf16(pa)
ops.fusion_optimizer_step(plan2, pa, gr, ma, va, sa, lr=1e-3, weight_decay=1e-4, max_norm=1.0)
ref16 = plan2.pack_bf16(pa)
ops.fusion_optimizer_step_packed(plan2, pb, gr, mb, vb, sb, b16, lr=1e-3, weight_decay=1e-4, max_norm=1.0)
torch.cuda.synchronize()
print("opt p", float((pa - pb).abs().max()), "m", float((ma - mb).abs().max()), "v", float((va - vb).abs().max()), "state", sa.tolist(), sb.tolist())
r16, o16 = ref16.view(torch.bfloat16).float(), b16.view(torch.bfloat16).float()
bad = torch.nonzero(r16 != o16).flatten()
print("bf16 arena mismatches", bad.numel(), bad[:10].tolist(), "of", r16.numel())
for key, off, shape in plan2.slots:
    n = int(torch.Size(shape).numel())
    d = float((pa[off:off+n] - pb[off:off+n]).abs().max())
    if d > 0: print("  slot", key, d)
