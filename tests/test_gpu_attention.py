"""Stand-alone CrossModalAttention module (generic q_len / k_len) vs golden
vectors from the reference (attention.py:68-146)."""
import sys

import pytest
import torch

from conftest import Golden, dropin_src

pytestmark = pytest.mark.gpu


def _module(g, device):
    sys.path.insert(0, dropin_src())
    from attention import CrossModalAttention
    att = CrossModalAttention(24, 12, hidden_dim=32, num_heads=int(g["heads"]), dropout=0.0)
    att.load_state_dict(g.group("sd"))
    return att.to(device).eval()


@pytest.mark.parametrize("device", ["cuda", "cpu"])
def test_generic_lengths_match_reference(device):
    g = Golden("attention_generic.npz")
    att = _module(g, device)
    q, k, v = (g.t(n).to(device).requires_grad_(True) for n in ("q3", "k3", "v3"))
    out, w = att(q, k, v, g.t("mask2").to(device))
    assert out.device.type == device
    assert float((out.cpu() - g.t("out3")).abs().max()) <= 1e-5
    assert float((w.cpu() - g.t("w3")).abs().max()) <= 1e-6
    out.square().sum().backward()
    for t, name in ((q, "gq3"), (k, "gk3"), (v, "gv3")):
        assert float((t.grad.cpu() - g.t(name)).abs().max()) <= 1e-5, name
    for key, ref in g.group("grad3").items():
        got = dict(att.named_parameters())[key].grad
        assert float((got.cpu() - ref).abs().max()) <= 2e-5, key


def test_two_dim_inputs_are_gates_with_zero_qk_grads():
    g = Golden("attention_generic.npz")
    att = _module(g, "cuda")
    q, k, v = (g.t(n).cuda().requires_grad_(True) for n in ("q2", "k2", "v2"))
    out, w = att(q, k, v, g.t("mask1").cuda())
    assert out.shape == g.t("out2").shape and w.shape == g.t("w2").shape
    assert float((out.cpu() - g.t("out2")).abs().max()) <= 1e-5
    assert torch.equal(w.cpu(), g.t("w2"))
    assert not torch.isnan(out).any()
    out.sum().backward()
    assert q.grad is not None and float(q.grad.abs().max()) == 0.0
    assert k.grad is not None and float(k.grad.abs().max()) == 0.0
    assert float((v.grad.cpu() - g.t("gv2")).abs().max()) <= 1e-5
    for key, ref in g.group("grad2").items():
        got = dict(att.named_parameters())[key].grad
        assert got is not None and float((got.cpu() - ref).abs().max()) <= 2e-5, key


def _maxabs(a, b):
    return float((a.detach().cpu().double() - b.detach().cpu().double()).abs().max())


def test_temporal_attention_matches_reference_golden():
    """TemporalAttention (src/attention.py:149-281) against the unmodified reference: attended steps, attention
    weights (masked keys get exactly zero), the reference's broadcast of the masked output, pool_sequence, and the
    dropout-free training gradients."""
    sys.path.insert(0, dropin_src())
    from attention import TemporalAttention
    g = Golden("frame_temporal_small.npz")
    ta = TemporalAttention(24, hidden_dim=32, num_heads=4, dropout=0.0)
    ta.load_state_dict(g.group("temporal/sd"))
    ta = ta.cuda().eval()
    x, mask = g.t("frame/x").cuda(), g.t("temporal/mask").cuda()
    out, w = ta(x)
    assert _maxabs(out, g.t("temporal/out")) <= 1e-5 and _maxabs(w, g.t("temporal/weights")) <= 1e-6
    out_m, w_m = ta(x, mask)
    assert tuple(out_m.shape) == tuple(g["temporal/out_mask"].shape)
    assert _maxabs(out_m, g.t("temporal/out_mask")) <= 1e-5 and _maxabs(w_m, g.t("temporal/weights_mask")) <= 1e-6
    ref_w = g.t("temporal/weights_mask")
    assert torch.equal(w_m.cpu() == 0, ref_w == 0)     # masked keys: exact zeros in the same places
    assert _maxabs(ta.pool_sequence(x, w_m), g.t("temporal/pooled")) <= 1e-5
    ta.train()
    xg = x.clone().requires_grad_(True)
    o, _ = ta(xg, mask)
    (o * torch.linspace(-1, 1, 32, device="cuda").view(1, 1, 32)).sum().backward()
    assert _maxabs(xg.grad, g.t("temporal/gradx")) <= 1e-5
    grads = dict(ta.named_parameters())
    for key, ref in g.group("temporal/grad").items():
        assert _maxabs(grads[key].grad, ref) <= 5e-5, key


def test_pairwise_modality_attention_matches_reference_golden():
    """PairwiseModalityAttention (src/attention.py:284-413): per-modality attended embeddings and the per-pair
    attention maps (1 for a present key modality, NaN-cleaned 0 for an absent one: exact)."""
    sys.path.insert(0, dropin_src())
    from attention import PairwiseModalityAttention
    g = Golden("frame_temporal_small.npz")
    dims = {"video": 12, "imu": 20, "hr": 8}
    pa = PairwiseModalityAttention(dims, hidden_dim=32, num_heads=4, dropout=0.0)
    pa.load_state_dict(g.group("pairwise/sd"))
    pa = pa.cuda().eval()
    feats = {m: g.t(f"pairwise/x/{m}").cuda() for m in dims}
    out, maps = pa(feats, g.t("pairwise/mask").cuda())
    for m in dims:
        assert _maxabs(out[m], g.t(f"pairwise/out/{m}")) <= 1e-5, m
    ref_maps = g.group("pairwise/map")
    assert set(maps) == set(ref_maps)
    for key, ref in ref_maps.items():
        assert torch.equal(maps[key].detach().cpu().reshape(ref.shape), ref), key
