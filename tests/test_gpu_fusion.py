"""Parity of the CUDA HybridFusion path (through the C ABI) with the CPU oracle
and with golden vectors from the reference.  fp32 mode tolerance: max-abs
<= 1e-5 on logits / weights / gradients (BASELINE.json north_star); attention
maps, fallbacks and dead q/k gradients exact."""
import importlib

import pytest
import torch

from conftest import Golden, load_pkg
from helpers import PAMAP2, dropin_fusion, module_from_golden, seeded_case
from oracle import fusion_oracle

pytestmark = pytest.mark.gpu

TOL = 1e-5
CASES = ["fusion_tiny.npz", "fusion_pamap_small.npz", "fusion_missing_pair.npz", "fusion_tc_shape.npz"]


def _ops():
    return importlib.import_module(load_pkg().__name__ + ".ops")


def _maxabs(a, b):
    return float((a.detach().cpu().double() - b.detach().cpu().double()).abs().max())


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("where", ["cuda", "cpu_staged"])
def test_eval_matches_reference_golden(case, where):
    g = Golden(case)
    dev = "cuda" if where == "cuda" else "cpu"
    model = module_from_golden(g, device=dev).eval()
    feats = {k: v.to(dev) for k, v in g.group("x").items()}
    mask = g.t("mask").to(dev)
    logits, info = model(feats, mask, return_attention=True)
    assert logits.device.type == dev and logits.dtype == torch.float32
    assert _maxabs(logits, g.t("eval/logits")) <= TOL
    assert _maxabs(info["fusion_weights"], g.t("eval/fusion_weights")) <= TOL
    ref_maps = g.group("eval/attn")
    assert set(info["attention_maps"]) == set(ref_maps)
    for key, ref in ref_maps.items():
        got = info["attention_maps"][key].cpu()
        assert got.shape == ref.shape and torch.equal(got, ref), key  # exactly {0,1}
    assert _maxabs(model(feats), g.t("eval/logits_nomask")) <= TOL  # default mask = ones (fusion.py:357-360)
    # fallbacks are exact: all-missing row -> uniform, single modality -> one-hot
    fw = info["fusion_weights"].cpu()
    M = fw.shape[1]
    assert torch.equal(fw[1], torch.full((M,), 1.0 / M))
    assert torch.equal(fw[2], torch.eye(M)[M - 1])


@pytest.mark.parametrize("case", CASES)
def test_train_grads_match_reference_golden(case):
    g = Golden(case)
    ops = _ops()
    model = module_from_golden(g, device="cuda", dropout=0.0).train()
    feats = {k: v.cuda().requires_grad_(True) for k, v in g.group("x").items()}
    logits = model(feats, g.t("mask").cuda())
    loss, dlogits = ops.cross_entropy(logits.detach(), g.t("labels").cuda(), float(g["smoothing"]))
    logits.backward(dlogits)
    assert _maxabs(logits, g.t("train/logits")) <= TOL
    assert abs(float(loss) - float(g["train/loss"])) <= TOL
    grads = dict(model.named_parameters())
    for key, ref in g.group("grad").items():
        got = grads[key].grad
        assert got is not None, key  # dead q/k projections still get a tensor (SURVEY §8b autograd)
        assert _maxabs(got, ref) <= TOL, key
        if ".query_proj." in key or ".key_proj." in key:
            assert float(got.abs().max()) == 0.0, key
    for key, ref in g.group("gradx").items():
        assert _maxabs(feats[key].grad, ref) <= TOL, key


def _oracle_state(model):
    return {k: v.detach().cpu() for k, v in model.state_dict().items()}


def test_config2_shape_matches_oracle():
    """PAMAP2 shape of BASELINE config 2 (M=4, D=128, H=256, heads=4, C=25), B=384 (not a tile multiple)."""
    ops = _ops()
    model, feats, mask, labels = seeded_case(PAMAP2, 256, 4, 25, 384, seed=5, device="cuda")
    model.train()
    xs = {k: v.clone().requires_grad_(True) for k, v in feats.items()}
    logits, info = model(xs, mask, return_attention=True)
    loss, dlogits = ops.cross_entropy(logits.detach(), labels, 0.05)
    logits.backward(dlogits)

    sd = {k: v.clone().requires_grad_(True) for k, v in _oracle_state(model).items()}
    xo = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in feats.items()}
    ref_logits, ref_info = fusion_oracle.hybrid_fusion_forward(sd, model.modality_names, 4, xo, mask.cpu())
    ref_loss = fusion_oracle.cross_entropy_label_smoothing(ref_logits, labels.cpu(), 0.05)
    ref_loss.backward()
    assert _maxabs(logits, ref_logits) <= TOL
    assert _maxabs(info["fusion_weights"], ref_info["fusion_weights"]) <= TOL
    assert abs(float(loss) - float(ref_loss)) <= TOL
    for key, p in model.named_parameters():
        assert _maxabs(p.grad, sd[key].grad) <= TOL, key
    for key in xs:
        assert _maxabs(xs[key].grad, xo[key].grad) <= TOL, key
    # argmax is bit-exact wherever the top-2 gap exceeds the numerical tolerance
    top2 = ref_logits.detach().topk(2, dim=1).values
    safe = (top2[:, 0] - top2[:, 1]) > 4 * TOL
    assert torch.equal(logits.argmax(1).cpu()[safe], ref_logits.argmax(1)[safe])


def test_dropout_masks_injected_into_oracle():
    """Train mode, p = 0.1: dump the Philox masks the kernels drew, inject them
    into the oracle (whose dropout sites are pinned by fusion_pamap_dropout.npz)."""
    ops = _ops()
    g = Golden("fusion_pamap_small.npz")
    model = module_from_golden(g, device="cuda", dropout=0.1).train()
    plan = model._plan()
    names, B, H, heads, M = g.names, g.t("mask").shape[0], plan.H, plan.heads, plan.M
    torch.manual_seed(77)
    expect_seed = int(torch.randint(0, 2**62, (1,)).item())
    torch.manual_seed(77)
    feats = {k: v.cuda().requires_grad_(True) for k, v in g.group("x").items()}
    logits, info = model(feats, g.t("mask").cuda(), return_attention=True)
    loss, dlogits = ops.cross_entropy(logits.detach(), g.t("labels").cuda(), 0.05)
    logits.backward(dlogits)

    p = 0.1
    drops = {"input": {}, "proj": {}, "attn": {}}
    for m, name in enumerate(names):
        drops["input"][name] = ops.dropout_mask(expect_seed, 0, 0, m, B, plan.dims[m], p).cpu()
        drops["proj"][name] = ops.dropout_mask(expect_seed, 0, 1, m, B, H, p).cpu()
    for q in range(M):
        for k in range(M):
            if q != k:
                d = ops.dropout_mask(expect_seed, 0, 2, q * M + k, B, heads, p).cpu()
                drops["attn"][f"{names[q]}_to_{names[k]}"] = d.reshape(B, heads, 1, 1)
    drops["cls"] = ops.dropout_mask(expect_seed, 0, 3, 0, B, H, p).cpu()
    keep = torch.cat([d.flatten() for d in drops["proj"].values()])
    vals = torch.unique(keep).tolist()
    assert len(vals) == 2 and vals[0] == 0.0 and abs(vals[1] - 1.0 / 0.9) < 1e-6
    assert abs(float((keep > 0).float().mean()) - 0.9) < 0.03  # keep-rate, 1/(1-p) scaling

    sd = {k: v.clone().requires_grad_(True) for k, v in g.group("sd").items()}
    xo = {k: v.clone().requires_grad_(True) for k, v in g.group("x").items()}
    ref_logits, ref_info = fusion_oracle.hybrid_fusion_forward(sd, names, heads, xo, g.t("mask"), drops=drops)
    fusion_oracle.cross_entropy_label_smoothing(ref_logits, g.t("labels"), 0.05).backward()
    assert _maxabs(logits, ref_logits) <= TOL
    for key, ref in ref_info["attention_maps"].items():
        assert torch.equal(info["attention_maps"][key].cpu(), ref), key  # {0, 1/(1-p)} per (row, head)
    for key, prm in model.named_parameters():
        assert _maxabs(prm.grad, sd[key].grad) <= TOL, key
    for key in xo:
        assert _maxabs(feats[key].grad, xo[key].grad) <= TOL, key
    # same seed -> same masks -> same logits; different seed -> different
    torch.manual_seed(77)
    again = model({k: v.detach() for k, v in feats.items()}, g.t("mask").cuda())
    assert torch.equal(again, logits)
    torch.manual_seed(78)
    other = model({k: v.detach() for k, v in feats.items()}, g.t("mask").cuda())
    assert not torch.equal(other, logits)


def test_full_size_properties():
    """BASELINE config 2 size (B=4096): properties that need no oracle run."""
    model, feats, mask, _ = seeded_case(PAMAP2, 256, 4, 25, 4096, seed=9, device="cuda")
    model.eval()
    with torch.no_grad():
        logits, info = model(feats, mask, return_attention=True)
        # windows are independent: any batch split gives bit-identical rows
        half = model({k: v[1000:3000] for k, v in feats.items()}, mask[1000:3000])
        assert torch.equal(half, logits[1000:3000])
        # features of a missing modality cannot influence the result (fusion.py:370-374)
        noisy = {k: v.clone() for k, v in feats.items()}
        gone = mask[:, 2] == 0
        noisy["imu_ankle"][gone] = 1e3
        assert torch.equal(model(noisy, mask), logits)
        fw = info["fusion_weights"]
        assert torch.all(fw[mask == 0] == 0) and float((fw.sum(1) - 1).abs().max()) < 1e-6
        for gate in info["attention_maps"].values():
            assert set(torch.unique(gate).tolist()) <= {0.0, 1.0}
        k_idx = {n: i for i, n in enumerate(model.modality_names)}
        gmap = info["attention_maps"]["imu_hand_to_heart_rate"][:, 0, 0, 0]
        assert torch.equal(gmap, (mask[:, k_idx["heart_rate"]] != 0).float())
        # all-missing windows collapse to classifier(0)
        zero_mask = torch.zeros_like(mask[:8])
        out0 = model({k: v[:8] for k, v in feats.items()}, zero_mask)
        b1 = model.classifier[0].bias
        expect = torch.relu(b1) @ model.classifier[3].weight.t() + model.classifier[3].bias
        assert float((out0 - expect).abs().max()) <= TOL


def test_c_abi_rejects_bad_calls():
    pkg = load_pkg()
    ops = _ops()
    plan = ops.get_plan(["a", "b"], [8, 8], 16, 4, 3, [(0, 1), (1, 0)])
    arena = torch.zeros(plan.total, device="cuda")
    xs = [torch.zeros(4, 8, device="cuda")] * 2
    with pytest.raises(pkg.MsfError, match="workspace too small"):
        ops.fusion_forward_raw(plan, arena, xs, None, workspace=torch.empty(16, dtype=torch.uint8, device="cuda"))
    with pytest.raises(pkg.MsfError, match="divisible by num_heads"):
        ops.FusionPlan(["a", "b"], [8, 8], 10, 4, 3, [])
    with pytest.raises(pkg.MsfError, match="not eligible|params_bf16"):
        ops.fusion_forward_raw(plan, arena, xs, None, precision=pkg.native.MSF_PREC_BF16)


def test_out_of_range_labels_never_index_the_logits():
    """A label outside [0, C) (an ignore_index of -100, a class id >= C) must not read out of bounds: the stand-alone
    CE kernel treats the row as having no target class and counts it; the engine refuses a host batch that holds one."""
    import importlib
    import ctypes
    pkg = load_pkg()
    ops = importlib.import_module(pkg.__name__ + ".ops")
    engine = importlib.import_module(pkg.__name__ + ".engine")
    lib = pkg.lib()
    n = ctypes.c_int64()
    pkg.native.check(lib.msf_bad_label_count(ctypes.byref(n), 1))
    logits = torch.randn(64, 25, device="cuda")
    labels = torch.randint(0, 25, (64,), device="cuda")
    labels[3], labels[40] = -100, 25
    loss, grad = ops.cross_entropy(logits, labels, smoothing=0.0)
    assert torch.isfinite(loss).all() and torch.isfinite(grad).all()
    ok = torch.ones(64, dtype=torch.bool, device="cuda")
    ok[3] = ok[40] = False
    ref = torch.nn.functional.cross_entropy(logits[ok], labels[ok], reduction="sum")
    lse = torch.logsumexp(logits[~ok], dim=1).sum()     # rows without a target class contribute their lse
    assert abs(float(loss) * 64 - float(ref + lse)) <= 1e-3
    pkg.native.check(lib.msf_bad_label_count(ctypes.byref(n), 1))
    assert n.value == 2
    from helpers import PAMAP2, seeded_case
    model, feats, mask, labels = seeded_case(PAMAP2, 256, 4, 25, 64, seed=3, device="cpu")
    eng = engine.FusionEngine(model, 64, precision="fp32", use_graph=False)
    bad = labels.clone()
    bad[5] = 25
    with pytest.raises(IndexError):
        eng.load_batch(feats, mask, bad)
