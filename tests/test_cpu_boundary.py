"""CPU-only checks of the drop-in boundary: the C-ABI library loads and exports
every declared symbol, the arena layout equals the reference's parameter order,
module surfaces / error strings match, and nothing silently falls back."""
import ctypes
import os
import re
import runpy
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, Golden, dropin_src, load_pkg
from helpers import dropin_fusion, golden_dims, module_from_golden

HEADER = os.path.join(ROOT, "include", "msf_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(msf_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(pkg):
    lib = pkg.lib()
    names = declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/msf_b200.h but not exported"
        assert name in pkg.native.PROTOTYPES, f"{name} has no ctypes prototype"
    assert lib.msf_abi_version() == pkg.native.MSF_ABI_VERSION


def test_struct_layout_matches_header(pkg):
    n = pkg.native
    assert ctypes.sizeof(n.FusionShape) == 4 * 4 + 4 * 8 + 8
    # int32 x3, float, u64 x2, 2 ptr, 8 ptr, 3 ptr(+size_t), 3 ptr, 2 ptr, 8 ptr, 1 ptr, LayerNorm: 8 + 8 ptr, float (+pad)
    assert ctypes.sizeof(n.FusionCall) == 16 + 16 + 8 * (1 + 2 + 8 + 1 + 2 + 3 + 2 + 8 + 1) + 8 * 16 + 8


def test_header_is_plain_c_and_lstm_record_matches_field_by_field(pkg, tmp_path):
    """include/msf_b200.h is what a foreign-function binding reads: it must compile as C99 on its own (no C++, no torch
    types), and the ctypes mirror of msf_lstm_seq (the record of the recurrence entry points) must agree with the
    compiler about the offset of every field."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no C compiler")
    n = pkg.native
    fields = [name for name, _ in n.LstmSeq._fields_]
    src = tmp_path / "layout.c"
    lines = ['#include <stddef.h>', '#include <stdio.h>', '#include "msf_b200.h"', "int main(void) {",
             '  printf("%zu\\n", sizeof(msf_lstm_seq));']
    lines += [f'  printf("%zu\\n", offsetof(msf_lstm_seq, {f}));' for f in fields]
    lines += ["  return 0;", "}"]
    src.write_text("\n".join(lines) + "\n")
    exe = tmp_path / "layout"
    subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.dirname(HEADER), str(src), "-o", str(exe)],
                   check=True)
    out = [int(v) for v in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    assert out[0] == ctypes.sizeof(n.LstmSeq) == pkg.lib().msf_lstm_seq_bytes()
    assert out[1:] == [getattr(n.LstmSeq, f).offset for f in fields]


@pytest.mark.parametrize("case", ["fusion_tiny.npz", "fusion_pamap_small.npz", "fusion_missing_pair.npz"])
def test_arena_layout_is_reference_parameter_order(pkg, case):
    g = Golden(case)
    model = module_from_golden(g)
    ops = __import__(pkg.__name__ + ".ops", fromlist=["ops"])
    plan = model._plan()
    own = dict(model.named_parameters())
    assert [k for k, _, _ in plan.slots] == list(own.keys())  # registration order (fusion.py:291-328)
    # offsets are the running sum over ALL pair slots (deleted pairs keep theirs)
    H, M = plan.H, plan.M
    expect = sum(H * d + H for d in plan.dims) + M * (M - 1) * 4 * (H * H + H) + M * (H + 1) \
        + H * H + H + plan.C * H + plan.C
    assert plan.total == expect
    if plan.dense:
        off = 0
        for key, o, shape in plan.slots:
            assert o == off, key
            off += int(np.prod(shape))
        assert off == plan.total


def test_same_seed_construction_matches_reference_init():
    """Registration order == reference order, so the same seed gives the same weights."""
    g = Golden("fusion_pamap_small.npz")
    torch.manual_seed(11)  # seed used by oracle/make_golden.py for this fixture
    model = dropin_fusion.HybridFusion(golden_dims(g), hidden_dim=32, num_classes=25, num_heads=4, dropout=0.0)
    for key, ref in g.group("sd").items():
        assert torch.equal(model.state_dict()[key], ref), key


def test_error_surface_matches_reference():
    HF = dropin_fusion.HybridFusion
    with pytest.raises(ValueError, match="No modalities configured for HybridFusion."):
        HF({}, num_classes=3)({}, None)
    model = HF({"video": 4, "imu": 4}, num_classes=3)
    with pytest.raises(KeyError, match="Missing features for modality 'imu' in HybridFusion forward pass."):
        model({"video": torch.randn(2, 4)})
    with pytest.raises(ValueError, match="modality_mask must be provided for adaptive weighting."):
        model.compute_adaptive_weights({}, None)
    with pytest.raises(KeyError, match="Missing aggregated features for modality"):
        model.compute_adaptive_weights({"video": torch.randn(2, 256)}, torch.ones(2, 2))
    with pytest.raises(ValueError, match="Unknown fusion type"):
        dropin_fusion.build_fusion_model("ensemble", {"video": 4}, num_classes=3)
    with pytest.raises(AssertionError):
        from attention import CrossModalAttention
        CrossModalAttention(8, 8, hidden_dim=10, num_heads=4)
    assert isinstance(dropin_fusion.build_fusion_model("late", {"a": 4}, 3, num_heads=2), dropin_fusion.LateFusion)
    assert isinstance(dropin_fusion.build_fusion_model("hybrid", {"a": 4}, 3, num_heads=2), HF)


def test_module_surface():
    model = dropin_fusion.HybridFusion({"video": 4, "imu": 6}, hidden_dim=8, num_classes=3, num_heads=2)
    assert model.modality_names == ["video", "imu"] and model.num_modalities == 2 and model.hidden_dim == 8
    assert isinstance(model.projections, torch.nn.ModuleDict) and isinstance(model.classifier, torch.nn.Sequential)
    assert list(model.attention_modules.keys()) == ["video_to_imu", "imu_to_video"]
    att = model.attention_modules["video_to_imu"]
    assert att.num_heads == 2 and att.head_dim == 4 and abs(att.scale - 0.5) < 1e-12
    assert model.projections["video"](torch.randn(2, 4)).shape == (2, 8)  # callable entry (test_fusion.py:384)
    import copy
    clone = copy.deepcopy(model).cpu()  # train.py:79-89 compile cache deep-copies modules
    assert clone.state_dict().keys() == model.state_dict().keys()


def test_early_and_late_fusion_keep_behaviour():
    torch.manual_seed(0)
    late = dropin_fusion.LateFusion({"video": 4, "imu": 4}, num_classes=3, hidden_dim=8, dropout=0.0).eval()
    feats = {"video": torch.randn(2, 4), "imu": torch.randn(2, 4)}
    fused, per = late(feats, torch.tensor([[1.0, 0.0], [0.0, 0.0]]))
    assert torch.allclose(fused[0], per["video"][0], atol=1e-6)
    assert torch.allclose(fused[1], (per["video"][1] + per["imu"][1]) / 2, atol=1e-6)
    early = dropin_fusion.EarlyFusion({"video": 4, "imu": 4}, num_classes=3, hidden_dim=8)
    assert early(feats).shape == (2, 3)


def test_entrypoint_block_reports(capsys, monkeypatch):
    monkeypatch.chdir(os.path.join(ROOT, load_pkg().__name__))  # reference tests read Path("src/fusion.py")
    runpy.run_module("fusion", run_name="__main__")
    out = capsys.readouterr().out.lower()
    assert "testing fusion architectures" in out


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful without a GPU")
def test_no_silent_cpu_fallback(pkg):
    model = dropin_fusion.HybridFusion({"video": 4, "imu": 4}, num_classes=3)
    with pytest.raises(pkg.MsfError, match="no CPU or PyTorch-eager fallback"):
        model({"video": torch.randn(2, 4), "imu": torch.randn(2, 4)})
    import uncertainty
    with pytest.raises(pkg.MsfError, match="no CPU"):
        uncertainty.CalibrationMetrics.expected_calibration_error(
            torch.tensor([0.8, 0.7]), torch.tensor([0, 1]), torch.tensor([0, 1]), num_bins=2)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: no product source may reference it."""
    pkg_dir = os.path.join(ROOT, load_pkg().__name__)
    for base, _, files in os.walk(pkg_dir):
        if os.sep + "build" in base:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(base, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert "/root/reference" not in text, f
