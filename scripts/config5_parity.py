"""Parity of the scaled variant of BASELINE configs[4] / SURVEY.md §8d config 5 (8 modalities, output_dim 256,
hidden 512, 8 heads, 11 classes) against the CPU oracle, through the drop-in module in fp32 and bf16.

    python scripts/config5_parity.py [batch]

This shape is outside the fused kernels (M > 4, H > 256), so it runs through the un-fused building blocks
(grouped tc_gemm launches, tail kernels).  NOT YET RUN ON A GPU (written when the round's GPU budget was spent);
once it passes, move its two halves into tests/test_gpu_fusion_bf16.py / tests/test_gpu_fusion.py (a parametrised
case of test_config2_shape_matches_oracle and a golden test over the *_seeded.npz fixtures).
The oracle is the checker only (tests/-style script, not a product path)."""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_pkg  # noqa: E402
from helpers import seeded_case  # noqa: E402
from oracle import fusion_oracle  # noqa: E402

pkg = load_pkg()
ops = importlib.import_module(pkg.__name__ + ".ops")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 384
DIMS = {f"video_{i}": 256 for i in range(2)}
DIMS.update({f"imu_{i}": 256 for i in range(6)})
worst = {}
for precision, tol in (("fp32", 1e-5), ("bf16", 1e-2)):
    model, feats, mask, labels = seeded_case(DIMS, 512, 8, 11, B, seed=5, device="cuda")
    model.precision = precision
    model.train()   # dropout p = 0 (seeded_case): the backward path without random masks
    xs = {k: v.clone().requires_grad_(True) for k, v in feats.items()}
    logits, info = model(xs, mask, return_attention=True)
    loss, dlogits = ops.cross_entropy(logits.detach(), labels, 0.05)
    logits.backward(dlogits)
    sd = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in model.state_dict().items()}
    xo = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in feats.items()}
    ref_logits, ref_info = fusion_oracle.hybrid_fusion_forward(sd, model.modality_names, 8, xo, mask.cpu())
    ref_loss = fusion_oracle.cross_entropy_label_smoothing(ref_logits, labels.cpu(), 0.05)
    ref_loss.backward()
    err = {
        "logits": float((logits.detach().cpu() - ref_logits.detach()).abs().max()),
        "fusion_weights": float((info["fusion_weights"].detach().cpu() - ref_info["fusion_weights"].detach()).abs().max()),
        "loss": abs(float(loss) - float(ref_loss)),
        "param_grads": max(float((p.grad.cpu() - sd[k].grad).abs().max()) for k, p in model.named_parameters()),
        "input_grads": max(float((xs[k].grad.cpu() - xo[k].grad).abs().max()) for k in xs),
    }
    dead = max(float(p.grad.abs().max()) for k, p in model.named_parameters()
               if ".query_proj." in k or ".key_proj." in k)
    maps_exact = all(torch.equal(info["attention_maps"][k].cpu(), v) for k, v in ref_info["attention_maps"].items())
    worst[precision] = max(err.values())
    print(precision, "tolerance", tol, err, "dead q/k grads max", dead, "attention maps exact", maps_exact)
    assert max(err.values()) <= tol and dead == 0.0 and maps_exact, precision
print("config 5 parity vs the oracle OK", worst)

# the same CUDA paths against the fixtures made from the unmodified reference at full width
# (tests/golden/fusion_config{2,5}_seeded.npz; tests/test_oracle_golden.py pins the oracle on them on the CPU)
from conftest import Golden  # noqa: E402
from helpers import module_from_seed  # noqa: E402

for case in ("fusion_config2_seeded.npz", "fusion_config5_seeded.npz"):
    g = Golden(case)
    for precision, tol in (("fp32", 1e-5), ("bf16", 1e-2)):
        model = module_from_seed(g, device="cuda", precision=precision)
        model.eval()
        feats = {k: v.cuda() for k, v in g.group("x").items()}
        with torch.no_grad():
            logits, info = model(feats, g.t("mask").cuda(), return_attention=True)
        e_log = float((logits.cpu() - g.t("eval/logits")).abs().max())
        e_fw = float((info["fusion_weights"].cpu() - g.t("eval/fusion_weights")).abs().max())
        keys = [str(k) for k in g["eval/attn_keys"]]
        stack = torch.stack([info["attention_maps"][k].reshape(logits.shape[0], -1).cpu() for k in keys])
        print(case, precision, "logits", e_log, "fusion weights", e_fw, "maps exact", torch.equal(stack, g.t("eval/attn_stack")))
        assert e_log <= tol and e_fw <= tol and torch.equal(stack, g.t("eval/attn_stack")), (case, precision)
print("full-width golden parity OK")
