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
