"""Device-side parity at the widths BASELINE.json names, against the unmodified reference's fixtures and the
CPU oracle (the checker, never the thing under test):

* `fusion_config2_seeded.npz` (configs[1]: M=4, D=128, H=256, 4 heads, 25 classes) and `fusion_config5_seeded.npz`
  (configs[4]: M=8, D=256, H=512, 8 heads, 11 classes) replayed through the drop-in HybridFusion on CUDA in fp32
  (max-abs <= 1e-5) and bf16 (<= 1e-2): eval logits / fusion weights / attention maps, train loss and the
  fixture's gradient views (norm, sum, first 64 elements per parameter, full input gradients);
* the fused bf16 train pass at the benchmarked B = 4096 against the oracle directly (not against another kernel
  sequence);
* `HybridFusion.compute_adaptive_weights` stand-alone against `oracle.adaptive_weights`, including the rows the
  reference's own test pins (`/root/reference/tests/test_fusion.py:50-80`: `[1, 0]` and the uniform fallback);
* `FusionEngine.set_lr`: a learning-rate change reaches a captured graph on replay.
"""
import importlib

import pytest
import torch

from conftest import Golden, load_pkg
from helpers import PAMAP2, module_from_seed, seeded_case
from oracle import fusion_oracle

pytestmark = pytest.mark.gpu

TOLS = {"fp32": 1e-5, "bf16": 1e-2}


def _ops():
    return importlib.import_module(load_pkg().__name__ + ".ops")


def _maxabs(a, b):
    return float((a.detach().cpu().double() - b.detach().cpu().double()).abs().max())


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("case", ["fusion_config2_seeded.npz", "fusion_config5_seeded.npz"])
def test_seeded_full_width_fixture_on_device(case, precision):
    g = Golden(case)
    tol = TOLS[precision]
    heads = int(g["heads"])
    model = module_from_seed(g, device="cuda", precision=precision)
    feats = {k: v.cuda() for k, v in g.group("x").items()}
    mask = g.t("mask").cuda()
    B = mask.shape[0]

    model.eval()
    with torch.no_grad():
        logits, info = model(feats, mask, return_attention=True)
    assert _maxabs(logits, g.t("eval/logits")) <= tol
    assert _maxabs(info["fusion_weights"], g.t("eval/fusion_weights")) <= tol
    keys = [str(k) for k in g["eval/attn_keys"]]
    assert sorted(info["attention_maps"]) == keys
    stack = torch.stack([info["attention_maps"][k].reshape(B, heads).cpu() for k in keys])
    assert torch.equal(stack, g.t("eval/attn_stack"))            # gates are exactly {0, 1}

    # train mode (dropout 0): loss and the reference's gradient views
    model.train()
    xs = {k: v.clone().requires_grad_(True) for k, v in feats.items()}
    out = model(xs, mask)
    loss, dlogits = _ops().cross_entropy(out.detach(), g.t("labels").cuda(), float(g["smoothing"]))
    out.backward(dlogits)
    assert abs(float(loss) - float(g["train/loss"])) <= tol
    for key, p in model.named_parameters():
        grad = p.grad.detach().cpu().reshape(-1)
        ref_norm = float(g["gnorm/" + key])
        dead = ".query_proj." in key or ".key_proj." in key
        if dead:
            assert ref_norm == 0.0 and float(grad.abs().max()) == 0.0, key   # tensors of exact zeros, not None
            continue
        assert _maxabs(grad[:64], g.t("ghead/" + key)) <= tol, key
        # whole-tensor views: the norm within the path's relative accuracy (bf16: ReLU-mask flips of ~0 units, see
        # test_gpu_fusion_bf16.py), the sum within what n elements at max-abs tol can move it
        rel = 1e-4 if precision == "fp32" else 0.10
        if grad.numel() < 16 and precision == "bf16":
            continue     # a scalar gradient of ~3e-5 (gating bias) is all rounding noise at bf16; max-abs checked above
        assert abs(float(grad.double().norm()) - ref_norm) <= rel * ref_norm + 1e-9, key
        assert abs(float(grad.double().sum()) - float(g["gsum/" + key])) <= rel * ref_norm * grad.numel() ** 0.5 + 1e-7, key
    for m, ref in g.group("gradx").items():
        assert _maxabs(xs[m].grad, ref) <= tol, m


def test_fused_bf16_train_pass_at_benchmark_batch_matches_oracle():
    """msf_fusion_train_pass (the benchmarked enqueue: projection, chained pair GEMMs, fused head, weight-gradient
    GEMM) at B = 4096 against the fp32 oracle on the same weights and batch; dropout 0 (the Philox-mask variant
    is test_gpu_fusion_bf16.py::test_dropout_masks_injected_into_oracle)."""
    ops = _ops()
    N = load_pkg().native
    B = 4096
    model, feats, mask, labels = seeded_case(PAMAP2, 256, 4, 25, B, seed=13, device="cuda")
    plan = model._plan()
    own = dict(model.named_parameters())
    arena = plan.gather([own[k].detach() for k, _, _ in plan.slots])
    xs = [feats[m].contiguous() for m in plan.names]
    logits, loss, grad, fw, gates = ops.fusion_train_pass_raw(
        plan, arena, xs, mask, labels, smoothing=0.05, precision=N.MSF_PREC_BF16, training=False, p=0.0,
        arena_bf16=plan.pack_bf16(arena))
    torch.cuda.synchronize()
    sd = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in model.state_dict().items()}
    xo = {k: v.cpu() for k, v in feats.items()}
    ref, info = fusion_oracle.hybrid_fusion_forward(sd, list(PAMAP2), 4, xo, mask.cpu())
    ref_loss = fusion_oracle.cross_entropy_label_smoothing(ref, labels.cpu(), 0.05)
    ref_loss.backward()
    assert _maxabs(logits, ref) <= 1e-2
    assert _maxabs(fw, info["fusion_weights"]) <= 1e-2
    assert abs(float(loss) - float(ref_loss.detach())) <= 1e-2
    names = list(PAMAP2)
    for pi, (q, k) in enumerate((q, k) for q in range(4) for k in range(4) if q != k):
        want = info["attention_maps"][f"{names[q]}_to_{names[k]}"].reshape(B, 4)
        assert torch.equal(gates[pi].cpu(), want), (q, k)
    for key, off, shape in plan.slots:
        n = int(torch.Size(shape).numel())
        got, want = grad[off:off + n].view(shape).cpu(), sd[key].grad
        assert _maxabs(got, want) <= 1e-2, key
        if ".query_proj." in key or ".key_proj." in key:
            assert float(got.abs().max()) == 0.0, key
        elif want.numel() >= 16:
            assert float((got.double() - want.double()).norm()) <= 0.10 * float(want.double().norm()) + 1e-12, key


@pytest.mark.parametrize("shape", [("tiny", {"video": 4, "imu": 4}, 8, 1, 3), ("pamap2", PAMAP2, 256, 4, 25)])
def test_compute_adaptive_weights_standalone_matches_oracle(shape):
    """HybridFusion.compute_adaptive_weights (src/fusion.py:429-479) on its own: gating scores, masked softmax,
    NaN -> 0, re-normalisation, mask / uniform fallbacks.  The first three rows are the reference's own pins."""
    _, dims, hidden, heads, classes = shape
    torch.manual_seed(0)
    model, _, _, _ = seeded_case(dims, hidden, heads, classes, 8, seed=0, device="cuda")
    M = len(dims)
    B = 67
    gen = torch.Generator().manual_seed(3)
    agg = {m: torch.randn(B, hidden, generator=gen) for m in dims}
    mask = (torch.rand(B, M, generator=gen) < 0.6).float()
    mask[0] = 1.0                      # all present
    mask[1] = 0.0
    mask[1, 0] = 1.0                   # exactly one present -> one-hot
    mask[2] = 0.0                      # all missing -> uniform 1/M
    mask[3] = 2.0                      # non-binary availability weights multiply the softmax (fusion.py:467)
    got = model.compute_adaptive_weights({m: a.cuda() for m, a in agg.items()}, mask.cuda())
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    want = fusion_oracle.adaptive_weights(sd, list(dims), agg, mask)
    assert got.shape == mask.shape and got.device.type == "cuda"
    assert _maxabs(got, want) <= 1e-6
    assert abs(float(got[0].sum()) - 1.0) <= 1e-6
    assert torch.equal(got[1].cpu(), torch.eye(M)[0])
    assert torch.equal(got[2].cpu(), torch.full((M,), 1.0 / M))
    some = mask.sum(1) > 0                       # all-missing rows take the uniform fallback instead
    assert torch.all(got.cpu()[some][mask[some] == 0] == 0)
    # CPU tensors in, CPU tensors out (staged through the same kernel)
    got_cpu = model.compute_adaptive_weights(agg, mask)
    assert got_cpu.device.type == "cpu" and _maxabs(got_cpu, want) <= 1e-6
    with pytest.raises(ValueError, match="modality_mask must be provided"):
        model.compute_adaptive_weights(agg, None)
    with pytest.raises(KeyError, match="Missing aggregated features for modality"):
        model.compute_adaptive_weights({k: v for k, v in list(agg.items())[:-1]}, mask)


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_set_lr_reaches_captured_graph(precision):
    """The optimizer launches read the learning rate from the device-side train state, so a captured step follows
    FusionEngine.set_lr on replay: with lr = 0 (and no weight decay) a replay leaves the parameters untouched,
    after set_lr(1e-3) the same graph moves them, and the first Adam step moves every live weight by ~lr."""
    engine = importlib.import_module(load_pkg().__name__ + ".engine")
    model, feats, mask, labels = seeded_case(PAMAP2, 256, 4, 25, 256, seed=3, device="cuda")
    eng = engine.FusionEngine(model, 256, precision=precision, seed=5, use_graph=True, lr=0.0, weight_decay=0.0)
    eng.p = 0.0
    eng.load_batch(feats, mask, labels)
    start = eng.arena.clone()
    eng.train_step_resident()                 # captures the graph
    eng.train_step_resident()
    torch.cuda.synchronize()
    assert torch.equal(eng.arena, start)
    eng.set_lr(1e-3)
    eng.train_step_resident()                 # same graph, new learning rate
    torch.cuda.synchronize()
    moved = (eng.arena - start).abs()
    assert float(moved.max()) > 1e-4
    assert float(moved.max()) <= 1.2e-3       # Adam: |update| <= lr / (1 - beta1^t) * ... ~ lr on the third step
    eng.set_lr(0.0)
    frozen = eng.arena.clone()
    eng.train_step_resident()
    torch.cuda.synchronize()
    assert torch.equal(eng.arena, frozen)
