"""The drop-in modules export every public name of the reference's src/attention.py and src/uncertainty.py
(the reference's tests import them: tests/test_attention.py:22-26, tests/test_uncertainty.py:14), and the
helpers that are plain tensor programs (uncertainty.py:19-71,286-492) behave as the reference's tests pin them
(tests/test_uncertainty.py:62-160).  No GPU: nothing here launches a kernel."""
import sys

import pytest
import torch

from conftest import dropin_src

if dropin_src() not in sys.path:
    sys.path.insert(0, dropin_src())

import attention as dropin_attention  # noqa: E402
import uncertainty as dropin_uncertainty  # noqa: E402


def test_public_names_of_the_reference_modules_exist():
    for name in ("CrossModalAttention", "TemporalAttention", "PairwiseModalityAttention", "visualize_attention"):
        assert hasattr(dropin_attention, name), name
    for name in ("MCDropoutUncertainty", "CalibrationMetrics", "UncertaintyWeightedFusion", "TemperatureScaling",
                 "EnsembleUncertainty", "compute_calibration_metrics", "main"):
        assert hasattr(dropin_uncertainty, name), name
    import inspect
    assert inspect.signature(dropin_uncertainty.compute_calibration_metrics).parameters["device"].default == "cpu"


def test_constructors_keep_the_reference_attribute_surface():
    t = dropin_attention.TemporalAttention(feature_dim=12, hidden_dim=32, num_heads=4, dropout=0.1)
    assert (t.feature_dim, t.hidden_dim, t.num_heads, t.head_dim) == (12, 32, 4, 8) and t.scale == 8 ** -0.5
    assert [n for n, _ in t.named_children()] == ["query_proj", "key_proj", "value_proj", "out_proj", "dropout"]
    p = dropin_attention.PairwiseModalityAttention({"video": 6, "audio": 5, "imu": 4}, hidden_dim=16, num_heads=2)
    assert p.modality_names == ["video", "audio", "imu"] and p.num_modalities == 3 and p.hidden_dim == 16
    assert list(p.attention_layers) == ["video_to_audio", "video_to_imu", "audio_to_video", "audio_to_imu",
                                        "imu_to_video", "imu_to_audio"]
    with pytest.raises(ValueError, match="No modalities provided"):
        dropin_attention.PairwiseModalityAttention({})({}, modality_mask=None)
    pooled = t.pool_sequence(torch.randn(3, 5, 32), torch.rand(3, 4, 5, 5))
    assert pooled.shape == (3, 32)
    with pytest.raises(ValueError):
        t.pool_sequence(torch.randn(3, 5, 32), torch.rand(3, 5, 5))


class _Tiny(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.fc = torch.nn.Linear(4, 3)
        self.drop = torch.nn.Dropout(0.5)

    def forward(self, x):
        return self.fc(self.drop(x))


def test_mc_dropout_and_ensemble():
    torch.manual_seed(0)
    model = _Tiny().eval()
    mean, spread = dropin_uncertainty.MCDropoutUncertainty(model, num_samples=6)(torch.randn(5, 4))
    assert mean.shape == (5, 3) and spread.shape == (5,) and bool((spread >= 0).all()) and not model.training
    ens = dropin_uncertainty.EnsembleUncertainty([_Tiny(), _Tiny()])
    probs, spread = ens.predict_with_uncertainty(torch.randn(5, 4))
    assert probs.shape == (5, 3) and spread.shape == (5,) and torch.allclose(probs.sum(1), torch.ones(5), atol=1e-6)
    with pytest.raises(ValueError):
        dropin_uncertainty.EnsembleUncertainty([]).predict_with_uncertainty(torch.randn(1, 4))


def test_uncertainty_weighted_fusion_weights_and_fallbacks():
    fusion = dropin_uncertainty.UncertaintyWeightedFusion()
    preds = {"a": torch.tensor([[1.0, 0.0], [0.5, 0.5]]), "b": torch.tensor([[0.0, 1.0], [0.2, 0.8]])}
    unc = {"a": torch.tensor([0.1, 0.2]), "b": torch.tensor([0.3, 0.2])}
    fused, w = fusion(preds, unc, torch.ones(2, 2))
    assert fused.shape == (2, 2) and torch.allclose(w.sum(1), torch.ones(2), atol=1e-5)
    assert w[0, 0] > w[0, 1] and torch.allclose(w[1], torch.full((2,), 0.5), atol=1e-5)
    _, w0 = fusion(preds, unc, torch.zeros(2, 2))
    assert torch.allclose(w0, torch.full_like(w0, 0.5))                    # tests/test_uncertainty.py:94-103
    _, w1 = fusion(preds, unc, torch.tensor([[1.0, 0.0], [0.0, 1.0]]))
    assert torch.allclose(w1, torch.tensor([[1.0, 0.0], [0.0, 1.0]]), atol=1e-5)
    with pytest.raises(ValueError):
        fusion({}, {}, torch.ones(1, 1))
    with pytest.raises(KeyError):
        fusion(preds, {"a": unc["a"]}, torch.ones(2, 2))


def test_temperature_scaling_calibrates_and_follows_the_logits_device():
    torch.manual_seed(1)
    logits = torch.randn(256, 5) * 6.0                      # over-confident
    labels = torch.randint(0, 5, (256,))
    ts = dropin_uncertainty.TemperatureScaling()
    before = torch.nn.functional.cross_entropy(ts(logits), labels).item()
    ts.calibrate(logits, labels, lr=0.1, max_iter=50)
    after = torch.nn.functional.cross_entropy(ts(logits), labels).item()
    assert after < before and ts.temperature.item() > 1.0
    ts.temperature = torch.nn.Parameter(torch.ones(1, device="meta"))   # tests/test_uncertainty.py:112-114
    ts.calibrate(logits[:8], labels[:8], lr=0.1, max_iter=5)
    assert ts.temperature.device.type == "cpu" and ts.temperature.item() > 0
