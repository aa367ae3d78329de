"""Builders shared by the parity tests."""
import os
import sys

import torch

from conftest import Golden, dropin_src

if dropin_src() not in sys.path:
    sys.path.insert(0, dropin_src())

import fusion as dropin_fusion  # noqa: E402  (the drop-in, found via sys.path like the reference's tests)


def golden_dims(g):
    return {m: int(g["x/" + m].shape[1]) for m in g.names}


def module_from_golden(g, device=None, dropout=None, precision="fp32"):
    """Drop-in HybridFusion carrying the golden state_dict (pairs the fixture deleted are deleted here too)."""
    p = float(g["drop_p"]) if dropout is None else dropout
    model = dropin_fusion.HybridFusion(golden_dims(g), hidden_dim=int(g["hidden"]), num_classes=int(g["classes"]),
                                       num_heads=int(g["heads"]), dropout=p)
    sd = g.group("sd")
    for key in list(model.attention_modules.keys()):
        if f"attention_modules.{key}.value_proj.weight" not in sd:
            del model.attention_modules[key]
    model.load_state_dict(sd, strict=True)
    model.precision = precision
    if device is not None:
        model = model.to(device)
    return model


def module_from_seed(g, device=None, precision="fp32"):
    """Drop-in HybridFusion rebuilt from a seeded fixture (oracle/make_golden.py: fusion_seeded_case): same seed,
    same constructor arguments -> the reference's parameters; the fixture's per-parameter sums are checked."""
    dims = {m: int(d) for m, d in zip(g.names, g["dims"])}
    torch.manual_seed(int(g["seed"]))
    model = dropin_fusion.HybridFusion(dims, hidden_dim=int(g["hidden"]), num_classes=int(g["classes"]),
                                       num_heads=int(g["heads"]), dropout=0.0)
    for key, v in model.state_dict().items():
        assert float(v.double().sum()) == float(g["sdsum/" + key]), key
    model.precision = precision
    return model if device is None else model.to(device)


def seeded_case(dims, hidden, heads, classes, batch, seed, device="cpu", mask_p=0.8):
    """Random-init drop-in module + PAMAP2-shaped synthetic windows (SURVEY §8d config 2 recipe)."""
    torch.manual_seed(seed)
    model = dropin_fusion.HybridFusion(dims, hidden_dim=hidden, num_classes=classes, num_heads=heads, dropout=0.0)
    gen = torch.Generator().manual_seed(seed + 1)
    feats = {m: torch.randn(batch, d, generator=gen) for m, d in dims.items()}
    mask = (torch.rand(batch, len(dims), generator=gen) < mask_p).float()
    dead = mask.sum(1) == 0
    mask[dead, torch.randint(0, len(dims), (int(dead.sum()),), generator=gen)] = 1.0  # data.py:327-341
    labels = torch.randint(0, classes, (batch,), generator=gen)
    if device != "cpu":
        model = model.to(device)
        feats = {k: v.to(device) for k, v in feats.items()}
        mask, labels = mask.to(device), labels.to(device)
    return model, feats, mask, labels


PAMAP2 = {"imu_hand": 128, "imu_chest": 128, "imu_ankle": 128, "heart_rate": 128}
