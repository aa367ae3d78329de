"""Pin the CPU oracle against golden vectors made from the unmodified reference
(oracle/make_golden.py).  CPU-only; runs in the build container and on the box."""
import numpy as np
import pytest
import torch

from conftest import Golden
from oracle import ece_oracle, fusion_oracle

FUSION_CASES = [
    "fusion_tiny.npz",
    "fusion_pamap_small.npz",
    "fusion_pamap_dropout.npz",
    "fusion_missing_pair.npz",
    "fusion_tc_shape.npz",
]


def _drops(g):
    if float(g["drop_p"]) == 0.0:
        return None
    return {
        "input": g.group("drop/input"),
        "proj": g.group("drop/proj"),
        "attn": g.group("drop/attn"),
        "cls": g.t("drop/cls"),
    }


@pytest.mark.parametrize("case", FUSION_CASES)
def test_fusion_eval_matches_reference(case):
    g = Golden(case)
    sd, feats, mask = g.group("sd"), g.group("x"), g.t("mask")
    heads = int(g["heads"])
    logits, info = fusion_oracle.hybrid_fusion_forward(sd, g.names, heads, feats, mask)
    # same ATen ops in the same order as fusion.py / attention.py -> bitwise on one build
    assert torch.allclose(logits, g.t("eval/logits"), atol=1e-6, rtol=0)
    assert torch.allclose(info["fusion_weights"], g.t("eval/fusion_weights"), atol=1e-7, rtol=0)
    for key, ref in g.group("eval/attn").items():
        got = info["attention_maps"][key]
        assert got.shape == ref.shape
        assert torch.equal(got, ref)
        assert set(torch.unique(got).tolist()) <= {0.0, 1.0}
    nomask, _ = fusion_oracle.hybrid_fusion_forward(sd, g.names, heads, feats, None)
    assert torch.allclose(nomask, g.t("eval/logits_nomask"), atol=1e-6, rtol=0)
    conf, pred = fusion_oracle.softmax_conf_pred(g.t("eval/logits"))
    assert torch.equal(pred, g.t("eval/pred"))
    assert torch.allclose(conf, g.t("eval/conf"), atol=1e-7, rtol=0)


SEEDED_CASES = ["fusion_config2_seeded.npz", "fusion_config5_seeded.npz"]


@pytest.mark.parametrize("case", SEEDED_CASES)
def test_full_size_shapes_match_reference(case):
    """BASELINE configs[1] (M=4, D=128, H=256, 4 heads, 25 classes) and the scaled variant configs[4] (M=8, D=256,
    H=512, 8 heads, 11 classes) at full width: the fixture holds the construction seed instead of the state dict
    (same seed + same constructor = the reference's parameters, checked through per-parameter sums), the inputs and
    the unmodified reference's logits, fusion weights, attention maps, loss and gradient views."""
    from helpers import module_from_seed

    g = Golden(case)
    model = module_from_seed(g)
    heads = int(g["heads"])
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in model.state_dict().items()}
    feats = {k: v.clone().requires_grad_(True) for k, v in g.group("x").items()}
    mask = g.t("mask")
    with torch.no_grad():
        logits, info = fusion_oracle.hybrid_fusion_forward(sd, g.names, heads, feats, mask)
    assert torch.allclose(logits, g.t("eval/logits"), atol=2e-6, rtol=0)
    assert torch.allclose(info["fusion_weights"], g.t("eval/fusion_weights"), atol=1e-7, rtol=0)
    keys = [str(k) for k in g["eval/attn_keys"]]
    assert sorted(info["attention_maps"]) == keys
    stack = torch.stack([info["attention_maps"][k].reshape(mask.shape[0], heads) for k in keys])
    assert torch.equal(stack, g.t("eval/attn_stack"))
    closed = fusion_oracle.hybrid_fusion_closed_form(
        {k: v.detach().double() for k, v in sd.items()}, g.names, heads,
        {k: v.detach().double() for k, v in feats.items()}, mask.double())
    assert torch.allclose(closed.float(), g.t("eval/logits"), atol=4e-6, rtol=0)
    # train mode (dropout 0): loss and gradients
    logits, _ = fusion_oracle.hybrid_fusion_forward(sd, g.names, heads, feats, mask)
    loss = fusion_oracle.cross_entropy_label_smoothing(logits, g.t("labels"), float(g["smoothing"]))
    loss.backward()
    assert abs(float(loss.detach()) - float(g["train/loss"])) <= 1e-6
    for key, p in sd.items():
        grad = torch.zeros_like(p).reshape(-1) if p.grad is None else p.grad.reshape(-1)
        ref_norm = float(g["gnorm/" + key])
        assert abs(float(grad.double().norm()) - ref_norm) <= 1e-5 * ref_norm + 1e-9, key
        assert abs(float(grad.double().sum()) - float(g["gsum/" + key])) <= 1e-4 * ref_norm + 1e-9, key
        assert torch.allclose(grad[:64], g.t("ghead/" + key), atol=1e-6 + 1e-4 * ref_norm, rtol=0), key
        if ".query_proj." in key or ".key_proj." in key:
            assert ref_norm == 0.0 and float(grad.abs().max()) == 0.0, key
    for m, ref in g.group("gradx").items():
        assert torch.allclose(feats[m].grad, ref, atol=1e-7, rtol=1e-4), m


@pytest.mark.parametrize("case", FUSION_CASES)
def test_fusion_closed_form_matches_reference(case):
    """q/k projections are dead, attention is a 0/1 gate."""
    g = Golden(case)
    sd = g.group("sd", torch.float64)
    feats = g.group("x", torch.float64)
    out = fusion_oracle.hybrid_fusion_closed_form(
        sd, g.names, int(g["heads"]), feats, g.t("mask", torch.float64)
    )
    assert torch.allclose(out.float(), g.t("eval/logits"), atol=2e-6, rtol=0)


@pytest.mark.parametrize("case", FUSION_CASES)
def test_fusion_train_grads_match_reference(case):
    g = Golden(case)
    sd = {k: v.clone().requires_grad_(True) for k, v in g.group("sd").items()}
    feats = {k: v.clone().requires_grad_(True) for k, v in g.group("x").items()}
    logits, info = fusion_oracle.hybrid_fusion_forward(
        sd, g.names, int(g["heads"]), feats, g.t("mask"), drops=_drops(g)
    )
    loss = fusion_oracle.cross_entropy_label_smoothing(
        logits, g.t("labels"), float(g["smoothing"])
    )
    loss.backward()
    assert torch.allclose(logits, g.t("train/logits"), atol=1e-6, rtol=0)
    assert torch.allclose(loss, g.t("train/loss"), atol=1e-6, rtol=0)
    for key, ref in g.group("train/attn").items():
        assert torch.allclose(info["attention_maps"][key], ref, atol=0, rtol=0)
    for key, ref in g.group("grad").items():
        assert sd[key].grad is not None, key
        assert torch.allclose(sd[key].grad, ref, atol=2e-7, rtol=1e-5), key
        if ".query_proj." in key or ".key_proj." in key:
            assert float(ref.abs().max()) == 0.0  # SURVEY §7 hard part 2
    for key, ref in g.group("gradx").items():
        assert torch.allclose(feats[key].grad, ref, atol=2e-7, rtol=1e-5), key


def test_optimizer_step_matches_torch_adamw():
    g = Golden("fusion_pamap_small.npz")
    grads = {k: v.clone() for k, v in g.group("grad").items()}
    total = fusion_oracle.clip_grad_norm(list(grads.values()), 1.0)
    assert abs(total - float(g["opt/grad_norm"])) < 1e-5
    for key, p0 in g.group("sd").items():
        p = p0.clone()
        m, v = torch.zeros_like(p), torch.zeros_like(p)
        fusion_oracle.adamw_step(p, grads[key], m, v, step=1)
        assert torch.allclose(p, g.t("opt/" + key), atol=1e-7, rtol=1e-6), key


def test_attention_generic_matches_reference():
    g = Golden("attention_generic.npz")
    heads = int(g["heads"])
    sd = {"a." + k: v.clone().requires_grad_(True) for k, v in g.group("sd").items()}
    q, k, v = (g.t(n).clone().requires_grad_(True) for n in ("q3", "k3", "v3"))
    out, w = fusion_oracle.cross_modal_attention(sd, "a", q, k, v, heads, g.t("mask2"))
    assert torch.allclose(out, g.t("out3"), atol=1e-6) and torch.allclose(w, g.t("w3"), atol=1e-7)
    out.square().sum().backward()
    for t, name in ((q, "gq3"), (k, "gk3"), (v, "gv3")):
        assert torch.allclose(t.grad, g.t(name), atol=1e-6), name
    for key, ref in g.group("grad3").items():
        assert torch.allclose(sd["a." + key].grad, ref, atol=1e-5, rtol=1e-5), key
    # HybridFusion-style 2-D call: weights are exactly {0,1}, q/k grads exactly zero
    sd = {"a." + k: v.clone().requires_grad_(True) for k, v in g.group("sd").items()}
    q, k, v = (g.t(n).clone().requires_grad_(True) for n in ("q2", "k2", "v2"))
    out, w = fusion_oracle.cross_modal_attention(sd, "a", q, k, v, heads, g.t("mask1"))
    assert torch.allclose(out, g.t("out2"), atol=1e-6) and torch.equal(w, g.t("w2"))
    out.sum().backward()
    assert torch.equal(q.grad, g.t("gq2")) and float(q.grad.abs().max()) == 0.0
    assert torch.equal(k.grad, g.t("gk2")) and float(k.grad.abs().max()) == 0.0
    assert torch.allclose(v.grad, g.t("gv2"), atol=1e-6)


@pytest.mark.parametrize("nb", [15, 10, 2])
def test_ece_bins_match_reference(nb):
    g = Golden("ece_seeded.npz")
    conf, pred, label = g["conf"], g["pred"].astype(np.int64), g["label"].astype(np.int64)
    edges = ece_oracle.linspace_f32(nb)
    assert np.array_equal(edges, g[f"edges_f32/{nb}"])  # torch.linspace bit pattern
    cnt, cor, cs = ece_oracle.bin_masks(conf, pred, label, edges)
    assert np.array_equal(cnt, g[f"count_f32/{nb}"])
    assert np.array_equal(cor, g[f"correct_f32/{nb}"])
    assert np.allclose(cs, g[f"confsum_f32/{nb}"], rtol=1e-12)
    ece, mce = ece_oracle.ece_from_bins(cnt, cor, cs, conf.shape[0])
    assert abs(ece - float(g[f"ece/{nb}"])) <= 1e-6 * max(1.0, abs(ece))
    assert abs(mce - float(g[f"mce/{nb}"])) <= 1e-6 * max(1.0, abs(mce))
    cnt64, _, _ = ece_oracle.bin_masks(conf, pred, label, ece_oracle.linspace_f64(nb))
    assert np.array_equal(cnt64, g[f"count_f64/{nb}"])
    # C restatement agrees with numpy, single- and multi-threaded
    for threads in (1, 4):
        c2, k2, s2 = ece_oracle.bin_masks_c(conf, pred, label, edges, threads)
        assert np.array_equal(c2, cnt) and np.array_equal(k2, cor)
        assert np.allclose(s2, cs, rtol=1e-12)


def test_ece_survey_kats():
    g = Golden("ece_kat.npz")
    assert float(g["ece2"]) == 0.25 and float(g["mce2"]) == 0.25
    e = ece_oracle.expected_calibration_error([0.8, 0.7], [0, 1], [0, 1], 2)
    m = ece_oracle.maximum_calibration_error([0.8, 0.7], [0, 1], [0, 1], 2)
    assert abs(e - 0.25) < 1e-7 and abs(m - 0.25) < 1e-7
    gen = torch.Generator().manual_seed(1234)
    logits = torch.randn(100000, 25, generator=gen) * 2
    labels = torch.randint(0, 25, (100000,), generator=gen)
    conf, pred = fusion_oracle.softmax_conf_pred(logits)
    cnt, cor, cs = ece_oracle.bin_masks_c(
        conf.numpy(), pred.numpy(), labels.numpy(), ece_oracle.linspace_f32(15)
    )
    assert cnt.tolist() == [0, 272, 6916, 16731, 18030, 15373, 11779, 8811, 6616,
                            5085, 3829, 2842, 2001, 1262, 453]
    ece, mce = ece_oracle.ece_from_bins(cnt, cor, cs, 100000)
    assert abs(ece - float(g["ece15_seed1234"])) < 1e-6
    assert abs(mce - float(g["mce15_seed1234"])) < 1e-6
