"""Per-modality LayerNorm between encoders and fusion (src/train.py:170-171,267-268):
the stand-alone msf_layer_norm_* kernels against the reference's golden vector and torch's CPU op, the LayerNorm
fused into the projection kernel against the un-fused sequence and the oracle, and the encoder -> LN -> fusion
glue (pipeline.EncodeFuse) against the same wiring done by hand with the oracle."""
import importlib

import pytest
import torch

from conftest import Golden, load_pkg
from helpers import PAMAP2, seeded_case
from oracle import encoder_oracle, fusion_oracle

pytestmark = pytest.mark.gpu


def _ops():
    return importlib.import_module(load_pkg().__name__ + ".ops")


def _maxabs(a, b):
    return float((a.detach().cpu().double() - b.detach().cpu().double()).abs().max())


def test_layer_norm_kernel_matches_reference_golden():
    g = Golden("encoders_small.npz")
    ops = _ops()
    x, w, b = g.t("lstm/out").cuda(), g.t("ln/weight").cuda(), g.t("ln/bias").cuda()
    assert _maxabs(ops.layer_norm(x, w, b, 1e-5), g.t("ln/out")) <= 1e-5
    assert _maxabs(ops.layer_norm(x, w, b, 1e-5), encoder_oracle.layer_norm(x.cpu(), w.cpu(), b.cpu())) <= 1e-5


@pytest.mark.parametrize("rows,dim", [(1, 16), (37, 128), (4096, 128), (300, 256), (5, 100)])
def test_layer_norm_forward_backward_match_torch(rows, dim):
    ops = _ops()
    gen = torch.Generator().manual_seed(rows + dim)
    x = (torch.randn(rows, dim, generator=gen) * 3 + 1.5).requires_grad_(True)
    w = (torch.rand(dim, generator=gen) + 0.5).requires_grad_(True)
    b = (torch.rand(dim, generator=gen) - 0.5).requires_grad_(True)
    gy = torch.randn(rows, dim, generator=gen)
    ref = torch.nn.functional.layer_norm(x, (dim,), w, b, 1e-5)
    ref.backward(gy)
    xc, wc, bc = (t.detach().cuda().requires_grad_(True) for t in (x, w, b))
    out = ops.layer_norm(xc, wc, bc, 1e-5)
    out.backward(gy.cuda())
    assert _maxabs(out, ref) <= 1e-5
    assert _maxabs(xc.grad, x.grad) <= 2e-5
    scale = max(1.0, float(w.grad.abs().max()), float(b.grad.abs().max()))
    assert _maxabs(wc.grad, w.grad) <= 1e-5 * scale * rows ** 0.5     # atomic column sums over the rows
    assert _maxabs(bc.grad, b.grad) <= 1e-5 * scale * rows ** 0.5
    # without affine parameters
    out2 = ops.layer_norm(xc.detach(), None, None, 1e-5)
    assert _maxabs(out2, torch.nn.functional.layer_norm(x.detach(), (dim,), None, None, 1e-5)) <= 1e-5


@pytest.mark.parametrize("precision,tol", [("bf16", 1e-2), ("fp32", 1e-5)])
@pytest.mark.parametrize("batch", [130, 1000])
def test_fused_input_layer_norm_matches_oracle(precision, tol, batch):
    """HybridFusion.forward(input_norms=...) — LayerNorm inside the projection kernel on the tensor-core path, the
    stand-alone kernel on the fp32 path — against torch LayerNorm + the fusion oracle on the CPU: logits, gradients
    of the raw encoder outputs, of the LayerNorm parameters and of the fusion parameters (dropout 0)."""
    ops = _ops()
    model, feats, mask, labels = seeded_case(PAMAP2, 256, 4, 25, batch, seed=17, device="cuda")
    model.precision = precision
    model.train()
    gen = torch.Generator().manual_seed(5)
    norms = torch.nn.ModuleDict({m: torch.nn.LayerNorm(d) for m, d in PAMAP2.items()})
    for n in norms.values():
        with torch.no_grad():
            n.weight.copy_(torch.rand(n.weight.shape, generator=gen) + 0.5)
            n.bias.copy_(torch.rand(n.bias.shape, generator=gen) * 0.4 - 0.2)
    norms_dev = norms.cuda()
    del norms_dev["imu_chest"]        # a modality without LayerNorm (train.py:267: `modality in self.layer_norms`)
    raw = {k: (v * 2.5 + 0.7).clone().requires_grad_(True) for k, v in feats.items()}   # un-normalised statistics
    logits = model(raw, mask, input_norms=norms_dev)
    loss, dlogits = ops.cross_entropy(logits.detach(), labels, 0.05)
    logits.backward(dlogits)

    sd = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in model.state_dict().items()}
    raw_cpu = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in raw.items()}
    ref_norms = {m: torch.nn.LayerNorm(PAMAP2[m]) for m in norms_dev}
    for m, n in ref_norms.items():
        n.load_state_dict({k: v.cpu() for k, v in norms_dev[m].state_dict().items()})
    normed = {m: (ref_norms[m](x) if m in ref_norms else x) for m, x in raw_cpu.items()}
    ref_logits, _ = fusion_oracle.hybrid_fusion_forward(sd, list(PAMAP2), 4, normed, mask.cpu())
    fusion_oracle.cross_entropy_label_smoothing(ref_logits, labels.cpu(), 0.05).backward()
    assert _maxabs(logits, ref_logits) <= tol
    for m in PAMAP2:
        assert _maxabs(raw[m].grad, raw_cpu[m].grad) <= tol, m
    for m, n in ref_norms.items():
        assert _maxabs(norms_dev[m].weight.grad, n.weight.grad) <= tol * 4, m     # column sums over the batch
        assert _maxabs(norms_dev[m].bias.grad, n.bias.grad) <= tol * 4, m
    for key, p in model.named_parameters():
        assert _maxabs(p.grad, sd[key].grad) <= tol, key
    # the fused kernel against the un-fused sequence (LayerNorm kernel first, then the fusion model)
    if precision == "bf16":
        assert ops.layer_norm_fused(model._plan(), ops.PRECISIONS["bf16"])
        model.eval()
        with torch.no_grad():
            a = model({k: v.detach() for k, v in raw.items()}, mask, input_norms=norms_dev)
            pre = {m: (ops.layer_norm(raw[m].detach(), norms_dev[m].weight, norms_dev[m].bias, 1e-5) if m in norms_dev
                       else raw[m].detach()) for m in PAMAP2}
            b = model(pre, mask)
        assert _maxabs(a, b) <= 2e-3     # same arithmetic up to where the bf16 rounding of the rows falls


def test_encode_fuse_pipeline_matches_reference_wiring():
    """pipeline.EncodeFuse = train.MultimodalFusionModule.forward (train.py:233-291) without Lightning: MLP encoders
    (eval mode) -> LayerNorm -> HybridFusion, a missing feature key skipped only where the fusion model allows it,
    tuple outputs split; against the same wiring on the CPU with torch modules + the fusion oracle."""
    pkg = load_pkg()
    pipeline = importlib.import_module(pkg.__name__ + ".pipeline")
    import encoders as dropin_encoders   # the drop-in (conftest puts its src/ on sys.path)
    import fusion as dropin_fusion
    torch.manual_seed(3)
    dims_in = {"imu_hand": 24, "imu_chest": 24, "heart_rate": 6}
    encs = torch.nn.ModuleDict({m: dropin_encoders.SimpleMLPEncoder(d, hidden_dim=32, output_dim=64, num_layers=2,
                                                                    dropout=0.0) for m, d in dims_in.items()})
    norms = torch.nn.ModuleDict({m: torch.nn.LayerNorm(64) for m in dims_in})
    fus = dropin_fusion.HybridFusion({m: 64 for m in dims_in}, hidden_dim=64, num_classes=7, num_heads=4, dropout=0.0)
    fus.precision = "bf16"
    model = pipeline.EncodeFuse(encs, fus, norms).cuda().eval()
    gen = torch.Generator().manual_seed(9)
    feats = {m: torch.randn(50, d, generator=gen).cuda() for m, d in dims_in.items()}
    mask = (torch.rand(50, 3, generator=gen) < 0.8).float().cuda()
    with torch.no_grad():
        logits, aux = model(feats, mask, return_attention=True)
        plain = model(feats, mask)
    assert torch.equal(plain, logits) and set(aux) == {"attention_maps", "fusion_weights"}
    # reference wiring on the CPU
    cpu = {m: encoder_oracle.mlp_encoder_forward({k: v.cpu() for k, v in encs[m].state_dict().items()},
                                                 feats[m].cpu(), 2, training=False) for m in dims_in}
    cpu = {m: torch.nn.functional.layer_norm(x, (64,), norms[m].weight.cpu(), norms[m].bias.cpu(), 1e-5)
           for m, x in cpu.items()}
    sd = {k: v.detach().cpu() for k, v in fus.state_dict().items()}
    ref, info = fusion_oracle.hybrid_fusion_forward(sd, list(dims_in), 4, cpu, mask.cpu())
    assert _maxabs(logits, ref) <= 1e-2
    assert _maxabs(aux["fusion_weights"], info["fusion_weights"]) <= 1e-2
    with pytest.raises(KeyError, match="Missing features for modality"):
        model({k: v for k, v in feats.items() if k != "heart_rate"}, mask)
    late = pipeline.EncodeFuse(encs, dropin_fusion.LateFusion({m: 64 for m in dims_in}, hidden_dim=32, num_classes=7,
                                                              dropout=0.0), norms).cuda().eval()
    with torch.no_grad():
        out = late(feats, mask)
    assert out.shape == (50, 7)
    with pytest.raises(ValueError, match="only available for HybridFusion"):
        late(feats, mask, return_attention=True)


def test_encode_fuse_trains_from_raw_windows_through_the_recurrence_kernels():
    """Boundary E in training mode: raw windows -> 2-layer LSTM SequenceEncoders (msf_lstm_forward in training mode,
    msf_lstm_backward) -> LayerNorm -> HybridFusion (tensor-core path) -> cross entropy, one backward pass through the
    drop-in modules.  Logits and the gradients that reach the LSTM parameters of every encoder are compared with the
    same wiring on the CPU (oracle encoders -> layer_norm -> fusion oracle -> autograd), bf16 tolerance."""
    pkg = load_pkg()
    pipeline = importlib.import_module(pkg.__name__ + ".pipeline")
    import encoders as dropin_encoders
    import fusion as dropin_fusion
    torch.manual_seed(21)
    feats_in = {"imu_hand": 17, "imu_chest": 17, "heart_rate": 1}
    B, T = 96, 20
    encs = torch.nn.ModuleDict({m: dropin_encoders.SequenceEncoder(f, hidden_dim=64, output_dim=64, num_layers=2,
                                                                   encoder_type="lstm", dropout=0.0)
                                for m, f in feats_in.items()})
    norms = torch.nn.ModuleDict({m: torch.nn.LayerNorm(64) for m in feats_in})
    fus = dropin_fusion.HybridFusion({m: 64 for m in feats_in}, hidden_dim=64, num_classes=7, num_heads=4, dropout=0.0)
    gen = torch.Generator().manual_seed(22)
    xs = {m: torch.randn(B, T, f, generator=gen) for m, f in feats_in.items()}
    mask = (torch.rand(B, 3, generator=gen) < 0.85).float()
    labels = torch.randint(0, 7, (B,), generator=gen)
    # reference wiring on the CPU with autograd through the oracles
    enc_sd = {m: {k: v.detach().clone().requires_grad_(True) for k, v in encs[m].state_dict().items()} for m in feats_in}
    cpu = {m: encoder_oracle.sequence_encoder_forward(enc_sd[m], xs[m], 2, "lstm") for m in feats_in}
    cpu = {m: torch.nn.functional.layer_norm(x, (64,), norms[m].weight.detach(), norms[m].bias.detach(), 1e-5)
           for m, x in cpu.items()}
    sd = {k: v.detach().clone() for k, v in fus.state_dict().items()}
    ref_logits, _ = fusion_oracle.hybrid_fusion_forward(sd, list(feats_in), 4, cpu, mask)
    torch.nn.functional.cross_entropy(ref_logits, labels).backward()
    # the drop-in on the device
    for e in encs.values():
        e.precision = "bf16"
    fus.precision = "bf16"
    model = pipeline.EncodeFuse(encs, fus, norms).cuda().train()
    logits = model({m: x.cuda() for m, x in xs.items()}, mask.cuda())
    torch.nn.functional.cross_entropy(logits, labels.cuda()).backward()
    assert _maxabs(logits, ref_logits) <= 2e-2
    for m in feats_in:
        for name, prm in encs[m].rnn.named_parameters():
            ref = enc_sd[m]["rnn." + name].grad
            assert prm.grad is not None and torch.isfinite(prm.grad).all(), (m, name)
            err = float((prm.grad.cpu().double() - ref.double()).norm() / ref.double().norm().clamp_min(1e-30))
            assert err <= 0.1, (m, name, err)
    # inference through the same grouped launches, and group == one encoder at a time (rows are independent)
    model.eval()
    with torch.no_grad():
        dev_x = {m: x.cuda() for m, x in xs.items()}
        eval_logits = model(dev_x, mask.cuda())
        names = list(feats_in)
        grouped = dropin_encoders.SequenceEncoder.forward_group([encs[m] for m in names], [dev_x[m] for m in names])
        for m, g in zip(names, grouped):
            assert encs[m].group_key(dev_x[m]) is not None
            assert torch.equal(g, encs[m](dev_x[m])), m
    assert _maxabs(eval_logits, ref_logits) <= 2e-2
