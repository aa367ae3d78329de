"""Per-tensor error report of the bf16 path vs the fp32 oracle (debug aid, GPU)."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import PAMAP2, seeded_case
from conftest import load_pkg
from oracle import fusion_oracle
ops = importlib.import_module(load_pkg().__name__ + ".ops")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 384
model, feats, mask, labels = seeded_case(PAMAP2, 256, 4, 25, B, seed=5, device="cuda")
model.precision = sys.argv[2] if len(sys.argv) > 2 else "bf16"
model.train()
xs = {k: v.clone().requires_grad_(True) for k, v in feats.items()}
logits, info = model(xs, mask, return_attention=True)
loss, dlogits = ops.cross_entropy(logits.detach(), labels, 0.05)
logits.backward(dlogits)
sd = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in model.state_dict().items()}
xo = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in feats.items()}
ref_logits, ref_info = fusion_oracle.hybrid_fusion_forward(sd, model.modality_names, 4, xo, mask.cpu())
fusion_oracle.cross_entropy_label_smoothing(ref_logits, labels.cpu(), 0.05).backward()
print("logits err", float((logits.cpu() - ref_logits).abs().max()), "scale", float(ref_logits.abs().max()))
rows = []
for key, p in model.named_parameters():
    ref = sd[key].grad
    err = float((p.grad.cpu() - ref).abs().max()); sc = float(ref.abs().max())
    rel_fro = float((p.grad.cpu() - ref).norm() / (ref.norm() + 1e-30))
    rows.append((err / (sc + 1e-30), key, err, sc, rel_fro))
for r in sorted(rows, reverse=True)[:14]:
    print("%-60s err %.3e scale %.3e rel_max %.3f rel_fro %.4f" % (r[1], r[2], r[3], r[0], r[4]))
for k in xs:
    ref = xo[k].grad
    print("dx", k, float((xs[k].grad.cpu() - ref).abs().max()), float(ref.abs().max()))
