"""Tensor-core (MSF_PREC_BF16: bf16 operands, fp32 accumulate in TMEM) HybridFusion
path against the fp32 CPU oracle and the reference's golden vectors.
Tolerance (BASELINE.json north_star): max-abs <= 1e-2 on logits, fusion weights
and gradients; attention gates, fallbacks and dead q/k gradients stay exact."""
import importlib
import os

import pytest
import torch

from conftest import Golden, load_pkg
from helpers import PAMAP2, module_from_golden, seeded_case
from oracle import fusion_oracle

pytestmark = pytest.mark.gpu

TOL = 1e-2


def _ops():
    return importlib.import_module(load_pkg().__name__ + ".ops")


def _maxabs(a, b):
    return float((a.detach().cpu().double() - b.detach().cpu().double()).abs().max())


def _relfro(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def test_eval_matches_reference_golden():
    g = Golden("fusion_tc_shape.npz")
    model = module_from_golden(g, device="cuda", precision="bf16").eval()
    feats = {k: v.cuda() for k, v in g.group("x").items()}
    mask = g.t("mask").cuda()
    logits, info = model(feats, mask, return_attention=True)
    assert logits.dtype == torch.float32
    assert _maxabs(logits, g.t("eval/logits")) <= TOL
    assert _maxabs(info["fusion_weights"], g.t("eval/fusion_weights")) <= TOL
    for key, ref in g.group("eval/attn").items():
        assert torch.equal(info["attention_maps"][key].cpu(), ref), key  # gates are exactly {0,1}
    fw = info["fusion_weights"].cpu()
    M = fw.shape[1]
    assert torch.equal(fw[1], torch.full((M,), 1.0 / M))      # all-missing window: uniform fallback, exact
    assert torch.equal(fw[2], torch.eye(M)[M - 1])            # single modality: one-hot, exact
    assert _maxabs(model(feats), g.t("eval/logits_nomask")) <= TOL


def test_train_grads_match_reference_golden():
    g = Golden("fusion_tc_shape.npz")
    ops = _ops()
    model = module_from_golden(g, device="cuda", dropout=0.0, precision="bf16").train()
    feats = {k: v.cuda().requires_grad_(True) for k, v in g.group("x").items()}
    logits = model(feats, g.t("mask").cuda())
    loss, dlogits = ops.cross_entropy(logits.detach(), g.t("labels").cuda(), float(g["smoothing"]))
    logits.backward(dlogits)
    assert _maxabs(logits, g.t("train/logits")) <= TOL
    assert abs(float(loss) - float(g["train/loss"])) <= TOL
    grads = dict(model.named_parameters())
    for key, ref in g.group("grad").items():
        got = grads[key].grad
        assert got is not None, key
        assert _maxabs(got, ref) <= TOL, key
        # relative check too: bf16 must track the gradient, not just stay under an absolute bound.
        # ReLU masks are taken from bf16 activations, so a few near-zero units flip; each flip is a
        # full-size error in that unit's gradient: Frobenius-relative error ~ sqrt(flip rate) ~ 5 %.
        if ref.numel() >= 16:
            assert _relfro(got, ref) <= 0.10, key
        if ".query_proj." in key or ".key_proj." in key:
            assert float(got.abs().max()) == 0.0, key
    for key, ref in g.group("gradx").items():
        assert _maxabs(feats[key].grad, ref) <= TOL, key


@pytest.mark.parametrize("batch", [384, 1000])
def test_config2_shape_matches_oracle(batch):
    """PAMAP2 shape of BASELINE config 2 (M=4, D=128, H=256, heads=4, C=25); B not a tile multiple."""
    ops = _ops()
    model, feats, mask, labels = seeded_case(PAMAP2, 256, 4, 25, batch, seed=5, device="cuda")
    model.precision = "bf16"
    model.train()
    xs = {k: v.clone().requires_grad_(True) for k, v in feats.items()}
    logits, info = model(xs, mask, return_attention=True)
    loss, dlogits = ops.cross_entropy(logits.detach(), labels, 0.05)
    logits.backward(dlogits)

    sd = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in model.state_dict().items()}
    xo = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in feats.items()}
    ref_logits, ref_info = fusion_oracle.hybrid_fusion_forward(sd, model.modality_names, 4, xo, mask.cpu())
    ref_loss = fusion_oracle.cross_entropy_label_smoothing(ref_logits, labels.cpu(), 0.05)
    ref_loss.backward()
    assert _maxabs(logits, ref_logits) <= TOL
    assert _maxabs(info["fusion_weights"], ref_info["fusion_weights"]) <= TOL
    assert abs(float(loss) - float(ref_loss)) <= TOL
    for key, p in model.named_parameters():
        ref = sd[key].grad
        assert _maxabs(p.grad, ref) <= TOL, key
        if ".query_proj." in key or ".key_proj." in key:
            assert float(p.grad.abs().max()) == 0.0, key
        elif ref.numel() >= 16:   # a scalar gradient of ~3e-5 (gating bias) is all rounding noise at bf16
            assert _relfro(p.grad, ref) <= 0.10, key  # see test_train_grads_match_reference_golden
    for key in xs:
        assert _maxabs(xs[key].grad, xo[key].grad) <= TOL, key
    for key, ref in ref_info["attention_maps"].items():
        assert torch.equal(info["attention_maps"][key].cpu(), ref), key
    top2 = ref_logits.detach().topk(2, dim=1).values
    safe = (top2[:, 0] - top2[:, 1]) > 4 * TOL
    assert torch.equal(logits.argmax(1).cpu()[safe], ref_logits.argmax(1)[safe])


def test_dropout_masks_injected_into_oracle():
    """Train mode, p = 0.1 on the tensor-core path: the Philox masks the GEMM epilogues
    drew are dumped, injected into the fp32 oracle, and everything must agree to bf16 tolerance."""
    ops = _ops()
    model, feats, mask, labels = seeded_case(PAMAP2, 256, 4, 25, 300, seed=11, device="cuda")
    model.precision = "bf16"
    model.dropout.p = 0.1
    for mod in model.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.1
    model.train()
    plan = model._plan()
    names, B, H, heads, M = model.modality_names, 300, plan.H, plan.heads, plan.M
    torch.manual_seed(77)
    expect_seed = int(torch.randint(0, 2**62, (1,)).item())
    torch.manual_seed(77)
    xs = {k: v.clone().requires_grad_(True) for k, v in feats.items()}
    logits, info = model(xs, mask, return_attention=True)
    loss, dlogits = ops.cross_entropy(logits.detach(), labels, 0.05)
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

    sd = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in model.state_dict().items()}
    xo = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in feats.items()}
    ref_logits, ref_info = fusion_oracle.hybrid_fusion_forward(sd, names, heads, xo, mask.cpu(), drops=drops)
    fusion_oracle.cross_entropy_label_smoothing(ref_logits, labels.cpu(), 0.05).backward()
    assert _maxabs(logits, ref_logits) <= TOL
    for key, ref in ref_info["attention_maps"].items():
        assert torch.equal(info["attention_maps"][key].cpu(), ref), key  # {0, 1/(1-p)} per (window, head)
    for key, prm in model.named_parameters():
        assert _maxabs(prm.grad, sd[key].grad) <= TOL, key
    for key in xo:
        assert _maxabs(xs[key].grad, xo[key].grad) <= TOL, key


def test_full_size_properties():
    """BASELINE config 2 size (B=4096): size-independent properties on the tensor-core path."""
    model, feats, mask, _ = seeded_case(PAMAP2, 256, 4, 25, 4096, seed=9, device="cuda")
    model.precision = "bf16"
    model.eval()
    with torch.no_grad():
        logits, info = model(feats, mask, return_attention=True)
        # windows are independent: any 128-aligned batch split gives bit-identical rows
        half = model({k: v[1024:3072] for k, v in feats.items()}, mask[1024:3072])
        assert torch.equal(half, logits[1024:3072])
        # features of a missing modality cannot influence the result (fusion.py:370-374)
        noisy = {k: v.clone() for k, v in feats.items()}
        gone = mask[:, 2] == 0
        noisy["imu_ankle"][gone] = 1e3
        assert torch.equal(model(noisy, mask), logits)
        fw = info["fusion_weights"]
        assert torch.all(fw[mask == 0] == 0) and float((fw.sum(1) - 1).abs().max()) < 1e-6
        for gate in info["attention_maps"].values():
            assert set(torch.unique(gate).tolist()) <= {0.0, 1.0}
        # fp32 path on the same weights: the two precisions agree to bf16 tolerance
        model.precision = "fp32"
        assert _maxabs(model(feats, mask), logits) <= TOL


def test_engine_bf16_tracks_fp32_engine():
    """Three optimizer steps of the fused train step (graph-captured) in bf16 follow the fp32 engine."""
    pkg = load_pkg()
    engine = importlib.import_module(pkg.__name__ + ".engine")
    losses = {}
    params = {}
    for prec in ("fp32", "bf16"):
        model, feats, mask, labels = seeded_case(PAMAP2, 256, 4, 25, 512, seed=21, device="cuda")
        eng = engine.FusionEngine(model, 512, precision=prec, seed=5, use_graph=True)
        eng.p = 0.0  # identical arithmetic up to precision
        out = []
        for _ in range(3):
            out.append(float(eng.train_step(feats, mask, labels).item()))
        losses[prec] = out
        params[prec] = eng.arena.clone()
    assert losses["bf16"][-1] < losses["bf16"][0]
    for a, b in zip(losses["fp32"], losses["bf16"]):
        assert abs(a - b) <= 2e-2, (losses,)
    assert float((params["fp32"] - params["bf16"]).abs().max()) <= 6.1e-3  # 3 Adam steps of lr 1e-3: sign flips of ~0 grads


def test_engine_step_clips_with_the_norm_of_its_own_gradients():
    """Inside the captured train step the optimizer takes the gradient square norm from the pass that wrote the
    gradients (msf_fusion_call.grad_sq -> MSF_OPT_NORM_GIVEN): after every step, what it used (sq_norm[0]) is the
    square norm of the gradient arena, on resident batches and across graph replays."""
    pkg = load_pkg()
    engine = importlib.import_module(pkg.__name__ + ".engine")
    model, feats, mask, labels = seeded_case(PAMAP2, 256, 4, 25, 512, seed=33, device="cuda")
    eng = engine.FusionEngine(model, 512, precision="bf16", seed=5, use_graph=True, max_grad_norm=0.05)
    eng.load_batch(feats, mask, labels)
    for _ in range(6):
        eng.train_step_resident()
        torch.cuda.synchronize()
        want = float((eng.grad.double() ** 2).sum())
        assert want > 0.05 ** 2          # the clip is active
        assert abs(float(eng.sq_norm[0]) - want) <= 1e-5 * want, (float(eng.sq_norm[0]), float(eng.sq_norm[1]), want)


@pytest.mark.parametrize("batch", [130, 512, 4096])
def test_gradient_slots_are_cleared_every_step(batch):
    """With lr = 0 and weight decay 0 the parameters never move, so every replay of the captured step must leave
    the same gradients: a slot that is accumulated into (bias column sums, gating layers) or never written (dead
    deleted pair modules) and not cleared by the step's first kernel would grow from step to step.  The live part
    of the arena is poisoned before the first step."""
    pkg = load_pkg()
    engine = importlib.import_module(pkg.__name__ + ".engine")
    model, feats, mask, labels = seeded_case(PAMAP2, 256, 4, 25, batch, seed=52, device="cuda")
    eng = engine.FusionEngine(model, batch, precision="bf16", seed=5, use_graph=True, lr=0.0, weight_decay=0.0)
    eng.p = 0.0
    eng.load_batch(feats, mask, labels)
    eng.grad.fill_(float("nan"))
    for key, off, shape in eng.plan.slots:   # the engine promises these stay zero (MSF_TRAIN_DEAD_SLOTS_ZERO)
        if ".query_proj." in key or ".key_proj." in key:
            eng.grad[off:off + int(torch.Size(shape).numel())].zero_()
    grads, losses = [], []
    for _ in range(4):
        losses.append(float(eng.train_step_resident().item()))
        torch.cuda.synchronize()
        grads.append(eng.grad.clone())
    assert torch.isfinite(grads[0]).all()
    scale = float(grads[0].abs().max())
    for g, l in zip(grads[1:], losses[1:]):
        assert abs(l - losses[0]) <= 1e-6 * abs(losses[0])
        assert float((g - grads[0]).abs().max()) <= 1e-4 * scale   # fp32 atomic column sums: rounding only
    for key, off, shape in eng.plan.slots:
        if ".query_proj." in key or ".key_proj." in key:
            n = int(torch.Size(shape).numel())
            assert float(grads[-1][off:off + n].abs().max()) == 0.0, key


def test_train_stream_pipeline_equals_step_by_step():
    """The 2-slot pipelined host-facing loop (H2D of batch i+1 overlapping step i, loss read one step late)
    follows the blocking load_batch + train_step loop (bias gradients are fp32 atomic column sums, so two
    runs agree to rounding, not bit for bit; Adam turns a sign flip of a ~0 gradient into 2*lr per step)."""
    pkg = load_pkg()
    engine = importlib.import_module(pkg.__name__ + ".engine")
    batches = []
    for i in range(5):
        _, feats, mask, labels = seeded_case(PAMAP2, 256, 4, 25, 512, seed=40 + i, device="cpu")
        batches.append(([f.pin_memory() for f in feats.values()], mask.pin_memory(), labels.pin_memory()))
    out = {}
    for mode in ("blocking", "stream"):
        model, *_ = seeded_case(PAMAP2, 256, 4, 25, 512, seed=21, device="cuda")
        eng = engine.FusionEngine(model, 512, precision="bf16", seed=5, use_graph=True)
        if mode == "blocking":
            losses = [float(eng.train_step(*b).item()) for b in batches]
        else:
            losses = list(eng.train_stream(iter(batches)))
            assert list(eng.train_stream(iter([]))) == []
        out[mode] = (losses, eng.arena.clone())
    assert len(out["stream"][0]) == len(batches)
    for a, b in zip(out["blocking"][0], out["stream"][0]):
        assert abs(a - b) <= 1e-3 * abs(a), out
    diff = (out["blocking"][1] - out["stream"][1]).abs()
    assert float(diff.max()) <= 1.01e-2 and float(diff.mean()) <= 1e-4


def test_resident_slots_and_subset_inference():
    """FusionEngine.add_resident_batch / train_step_slot (batches trained on in place, one captured graph per
    slot) follow the copy-in path; FusionEngine.infer_subset (uniform-mask hint) equals infer with that mask."""
    pkg = load_pkg()
    engine = importlib.import_module(pkg.__name__ + ".engine")
    data = [seeded_case(PAMAP2, 256, 4, 25, 512, seed=60 + i, device="cuda")[1:] for i in range(3)]
    out = {}
    for mode in ("copy", "slots"):
        model, *_ = seeded_case(PAMAP2, 256, 4, 25, 512, seed=21, device="cuda")
        eng = engine.FusionEngine(model, 512, precision="bf16", seed=5, use_graph=True)
        if mode == "copy":
            losses = [float(eng.train_step(f, m, y).item()) for f, m, y in data for _ in range(2)]
        else:
            slots = [eng.add_resident_batch({k: v.contiguous() for k, v in f.items()}, m, y) for f, m, y in data]
            losses = [float(eng.train_step_slot(s).item()) for s in slots for _ in range(2)]
        out[mode] = (losses, eng.arena.clone(), eng)
    for a, b in zip(out["copy"][0], out["slots"][0]):
        assert abs(a - b) <= 1e-3 * abs(a), out
    assert float((out["copy"][1] - out["slots"][1]).abs().max()) <= 1.3e-2   # Adam sign flips of ~0 gradients
    eng = out["slots"][2]
    with pytest.raises(ValueError):
        eng.add_resident_batch(data[0][0], data[0][1][:100], data[0][2])
    feats = data[0][0]
    for present in ([0], [1, 3], [0, 1, 2, 3]):
        mask = torch.zeros(512, 4, device="cuda")
        mask[:, present] = 1.0
        l_ref, c_ref, p_ref = (t.clone() for t in eng.infer(feats, mask))
        l_sub, c_sub, p_sub = (t.clone() for t in eng.infer_subset(feats, present))
        assert float((l_sub - l_ref).abs().max()) <= 5e-4   # bf16 rounding flips where the skipped tokens change the summation order
        assert torch.equal(p_sub, p_ref)
        l_again = eng.infer_subset(None, present)[0]          # features kept in the static buffers
        assert torch.equal(l_again, l_sub)
    with pytest.raises(ValueError):
        eng.infer_subset(feats, [])


def test_train_slots_equals_step_by_step():
    """FusionEngine.train_slots (one graph launch holding one optimizer step per listed slot, steps chained by
    programmatic dependent launch as in eager stream order) follows the same steps taken one graph launch each;
    every step's loss is reported, and the call can be repeated."""
    pkg = load_pkg()
    engine = importlib.import_module(pkg.__name__ + ".engine")
    data = [seeded_case(PAMAP2, 256, 4, 25, 512, seed=70 + i, device="cuda")[1:] for i in range(4)]
    out = {}
    for mode in ("single", "group", "eager"):
        model, *_ = seeded_case(PAMAP2, 256, 4, 25, 512, seed=21, device="cuda")
        eng = engine.FusionEngine(model, 512, precision="bf16", seed=5, use_graph=mode != "eager")
        slots = [eng.add_resident_batch({k: v.contiguous() for k, v in f.items()}, m, y) for f, m, y in data]
        losses = []
        for _ in range(2):
            if mode == "single":
                losses += [float(eng.train_step_slot(s).item()) for s in slots]
            else:
                got = eng.train_slots(slots)
                assert tuple(got.shape) == (len(slots),)
                losses += [float(v) for v in got.tolist()]
        torch.cuda.synchronize()
        out[mode] = (losses, eng.arena.clone(), int(eng.state[2].item()))
    assert out["single"][2] == out["group"][2] == out["eager"][2]   # 8 optimizer steps either way
    for mode in ("group", "eager"):
        for a, b in zip(out["single"][0], out[mode][0]):
            assert abs(a - b) <= 1e-3 * abs(a), out
        diff = (out["single"][1] - out[mode][1]).abs()
        assert float(diff.max()) <= 1.7e-2 and float(diff.mean()) <= 2e-4   # Adam sign flips of ~0 gradients
    assert eng.train_slots([]).numel() == 0


def test_pinned_batch_crosses_as_one_transfer_and_trains_the_same():
    """FusionEngine.pinned_batch(): host batches as views of one pinned allocation in the engine's copy order; the
    train_stream input slots are laid out the same way, so msf_memcpy_batch merges the six copies of a batch into
    one.  Same losses and parameters as separately pinned tensors."""
    pkg = load_pkg()
    engine = importlib.import_module(pkg.__name__ + ".engine")
    raw = [seeded_case(PAMAP2, 256, 4, 25, 512, seed=80 + i, device="cpu")[1:] for i in range(4)]
    out = {}
    for mode in ("separate", "packed"):
        model, *_ = seeded_case(PAMAP2, 256, 4, 25, 512, seed=21, device="cuda")
        eng = engine.FusionEngine(model, 512, precision="bf16", seed=5, use_graph=True)
        batches = []
        for feats, mask, labels in raw:
            if mode == "separate":
                batches.append(([f.pin_memory() for f in feats.values()], mask.pin_memory(), labels.pin_memory()))
            else:
                pf, pm, py = eng.pinned_batch()
                assert all(t.is_pinned() for t in pf + [pm, py])
                assert pf[1].data_ptr() == pf[0].data_ptr() + pf[0].numel() * 4   # adjacent, in copy order
                assert py.data_ptr() == pm.data_ptr() + pm.numel() * 4
                for dst, src in zip(pf + [pm, py], list(feats.values()) + [mask, labels]):
                    dst.copy_(src)
                batches.append((pf, pm, py))
        out[mode] = (list(eng.train_stream(iter(batches))), eng.arena.clone())
    for a, b in zip(out["separate"][0], out["packed"][0]):
        assert abs(a - b) <= 1e-3 * abs(a), out
    diff = (out["separate"][1] - out["packed"][1]).abs()
    assert float(diff.max()) <= 1.01e-2 and float(diff.mean()) <= 1e-4   # as test_train_stream_pipeline_equals_step_by_step


@pytest.mark.parametrize("batch", [300, 4096])
def test_bf16_features_equal_fp32_features_of_the_same_values(batch):
    """msf_fusion_call.x_bf16: features stored as bf16 rows give bit-identical logits and weight gradients to fp32
    rows holding the same (bf16-representable) values — the projection kernel only changes how it reads them —
    and stay within the bf16 tolerance of the oracle on the un-rounded features."""
    pkg, ops = load_pkg(), _ops()
    N = pkg.native
    model, feats, mask, labels = seeded_case(PAMAP2, 256, 4, 25, batch, seed=91, device="cuda")
    plan = model._plan()
    own = dict(model.named_parameters())
    arena = plan.gather([own[k].detach() for k, _, _ in plan.slots])
    a16 = plan.pack_bf16(arena)
    x16 = [feats[m].to(torch.bfloat16).contiguous() for m in plan.names]
    x32 = [x.float().contiguous() for x in x16]
    kw = dict(smoothing=0.05, precision=N.MSF_PREC_BF16, training=True, p=0.1, seed=11, arena_bf16=a16)
    got = {}
    for name, xs in (("bf16", x16), ("fp32", x32)):
        logits, loss, grad, fw, gates = ops.fusion_train_pass_raw(plan, arena, xs, mask, labels, **kw)
        got[name] = (logits.clone(), float(loss), grad.clone(), fw.clone())
    assert torch.equal(got["bf16"][0], got["fp32"][0])
    assert torch.equal(got["bf16"][3], got["fp32"][3])
    # weight-gradient GEMMs and their split halves are bit-reproducible; bias column sums are fp32 atomics
    assert _maxabs(got["bf16"][2], got["fp32"][2]) <= 1e-5
    # eval mode against the oracle on the ORIGINAL fp32 features: the extra input rounding stays inside 1e-2
    logits_eval, _, _, _ = ops.fusion_forward_raw(plan, arena, x16, mask, precision=N.MSF_PREC_BF16, arena_bf16=a16)
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    ref, _ = fusion_oracle.hybrid_fusion_forward(sd, list(PAMAP2), 4, {k: v.cpu() for k, v in feats.items()}, mask.cpu())
    assert _maxabs(logits_eval, ref) <= TOL


def test_bf16_features_are_refused_off_the_fused_projection_kernel():
    pkg, ops = load_pkg(), _ops()
    N = pkg.native
    model, feats, mask, labels = seeded_case(PAMAP2, 256, 4, 25, 64, seed=92, device="cuda")
    plan = model._plan()
    own = dict(model.named_parameters())
    arena = plan.gather([own[k].detach() for k, _, _ in plan.slots])
    x16 = [feats[m].to(torch.bfloat16).contiguous() for m in plan.names]
    with pytest.raises(pkg.MsfError):     # fp32 parity path reads fp32 features
        ops.fusion_forward_raw(plan, arena, x16, mask, precision=N.MSF_PREC_F32)
    mixed = [x16[0]] + [feats[m].contiguous() for m in plan.names[1:]]
    with pytest.raises(pkg.MsfError):     # one dtype for all modalities
        ops.fusion_forward_raw(plan, arena, mixed, mask, precision=N.MSF_PREC_BF16, arena_bf16=plan.pack_bf16(arena))


def test_engine_streams_bf16_host_batches():
    """FusionEngine(feature_dtype=bfloat16): pinned bf16 host batches (half the PCIe bytes) through train_stream follow
    an fp32-feature engine fed the same rounded values; resident fp32 batches still train on the same engine."""
    pkg = load_pkg()
    engine = importlib.import_module(pkg.__name__ + ".engine")
    raw = [seeded_case(PAMAP2, 256, 4, 25, 512, seed=60 + i, device="cpu")[1:] for i in range(5)]
    out = {}
    for mode in ("fp32", "bf16", "bf16-packed"):
        model, *_ = seeded_case(PAMAP2, 256, 4, 25, 512, seed=21, device="cuda")
        fd = torch.float32 if mode == "fp32" else torch.bfloat16
        eng = engine.FusionEngine(model, 512, precision="bf16", seed=5, use_graph=True, feature_dtype=fd)
        batches, moved = [], 0
        for feats, mask, labels in raw:
            f16 = [f.to(torch.bfloat16) for f in feats.values()]
            if mode == "bf16-packed":
                pf, pm, py = eng.pinned_batch()
                assert pf[0].dtype == torch.bfloat16 and pf[1].data_ptr() == pf[0].data_ptr() + pf[0].numel() * 2
                for dst, src in zip(pf + [pm, py], f16 + [mask, labels]):
                    dst.copy_(src)
                batches.append((pf, pm, py))
            else:
                fs = [(f if mode == "bf16" else f.float()).pin_memory() for f in f16]
                batches.append((fs, mask.pin_memory(), labels.pin_memory()))
            moved = sum(t.numel() * t.element_size() for t in batches[-1][0])
        out[mode] = (list(eng.train_stream(iter(batches))), eng.arena.clone(), moved)
        if mode == "bf16":   # a resident fp32 batch on the same engine
            slot = eng.add_resident_batch([f.cuda() for f in raw[0][0].values()], raw[0][1].cuda(), raw[0][2].cuda())
            assert torch.isfinite(eng.train_step_slot(slot)).all()
    assert out["bf16"][2] * 2 == out["fp32"][2]
    for mode in ("bf16", "bf16-packed"):
        for a, b in zip(out["fp32"][0], out[mode][0]):
            assert abs(a - b) <= 1e-3 * abs(a), out
        diff = (out["fp32"][1] - out[mode][1]).abs()
        assert float(diff.max()) <= 1.01e-2 and float(diff.mean()) <= 1e-4   # as test_train_stream_pipeline_equals_step_by_step


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_accumulated_step_equals_one_large_batch(precision):
    """FusionEngine.train_step_accumulated (accumulate_grad_batches, config/base.yaml:75): two micro-batches of 256
    windows give the gradient — and, dropout off, the optimizer step — of the 512-window batch they were cut from."""
    pkg = load_pkg()
    engine = importlib.import_module(pkg.__name__ + ".engine")
    out = {}
    for mode in ("whole", "micro"):
        model, feats, mask, labels = seeded_case(PAMAP2, 256, 4, 25, 512, seed=33, device="cuda")
        B = 512 if mode == "whole" else 256
        eng = engine.FusionEngine(model, B, precision=precision, seed=5, use_graph=False)
        eng.p = 0.0
        losses = []
        for _ in range(3):
            if mode == "whole":
                losses.append(float(eng.train_step(feats, mask, labels).item()))
            else:
                halves = [({k: v[s] for k, v in feats.items()}, mask[s], labels[s]) for s in (slice(0, 256), slice(256, 512))]
                losses.append(float(eng.train_step_accumulated(halves).item()))
        torch.cuda.synchronize()
        out[mode] = (losses, eng.arena.clone(), eng.grad.clone(), int(eng.state[2].item()))
    assert out["whole"][3] == out["micro"][3] == 4          # three optimizer steps either way (1-based counter)
    for a, b in zip(out["whole"][0], out["micro"][0]):
        assert abs(a - b) <= (1e-5 if precision == "fp32" else 1e-3) * abs(a), out
    gdiff = (out["whole"][2] - out["micro"][2]).abs()
    assert float(gdiff.max()) <= (2e-6 if precision == "fp32" else 2e-4)     # last step's (clipped-scale) gradient
    diff = (out["whole"][1] - out["micro"][1]).abs()
    assert float(diff.max()) <= 1.01e-2 and float(diff.mean()) <= 1e-4   # Adam sign flips of ~0 gradients
