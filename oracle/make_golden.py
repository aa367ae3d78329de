"""Generate tests/golden/*.npz by running the UNMODIFIED reference modules.

Run in the build container only (needs /root/reference; the GPU box has no
copy):  ``python oracle/make_golden.py``.  The reference's ``fusion.py``,
``attention.py`` and ``uncertainty.py`` are imported from
``/root/reference/src`` through ``sys.path`` exactly as the reference's own
tests do (tests/test_fusion.py:14-16); nothing is copied.  The resulting
fixtures pin the restatements in ``oracle/`` (tests/test_oracle_golden.py) and
are also replayed against the CUDA path (tests/test_gpu_*.py).

Dropout in train mode is injected by swapping the reference instance's
``nn.Dropout`` sub-modules for ``_QueueDrop`` (multiplies by a recorded mask);
the reference's forward code itself runs untouched.
"""

from __future__ import annotations

import itertools
import os
import sys

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

REF_SRC = "/root/reference/src"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def _import_reference():
    sys.path.insert(0, REF_SRC)
    import attention as ref_attention  # noqa
    import fusion as ref_fusion  # noqa
    import uncertainty as ref_uncertainty  # noqa

    for mod in (ref_attention, ref_fusion, ref_uncertainty):
        assert mod.__file__.startswith(REF_SRC), mod.__file__
    sys.path.pop(0)
    return ref_fusion, ref_attention, ref_uncertainty


class _QueueDrop(nn.Module):
    """Stand-in for nn.Dropout that applies pre-drawn masks in call order."""

    def __init__(self):
        super().__init__()
        self.queue = []

    def forward(self, x):
        if not self.training or not self.queue:
            return x
        return x * self.queue.pop(0)


def _draw(shape, p, gen):
    keep = (torch.rand(shape, generator=gen) >= p).float()
    return keep / (1.0 - p)


def _np(t):
    return t.detach().cpu().numpy().copy()  # copy: later in-place updates must not leak in


def _save(name, payload):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name)
    np.savez_compressed(path, **payload)
    print(f"wrote {path}  ({os.path.getsize(path) / 1024:.1f} KiB)")


def _masks(batch, m, gen, with_float=True):
    mask = (torch.rand(batch, m, generator=gen) < 0.7).float()
    mask[0] = 1.0  # all present
    mask[1] = 0.0  # all missing
    mask[2] = 0.0
    mask[2, m - 1] = 1.0  # single modality
    if with_float and batch > 3:
        mask[3] = 0.5  # non-binary availability (reference multiplies by it)
    return mask


def fusion_case(ref_fusion, name, dims, hidden, heads, classes, batch, seed,
                drop_p=0.0, delete_pairs=(), smoothing=0.05, optimizer=False):
    torch.manual_seed(seed)
    gen = torch.Generator().manual_seed(seed + 1)
    names = list(dims)
    model = ref_fusion.HybridFusion(
        dims, hidden_dim=hidden, num_classes=classes, num_heads=heads, dropout=drop_p
    )
    for key in delete_pairs:
        del model.attention_modules[key]
    feats = {m: torch.randn(batch, d, generator=gen) for m, d in dims.items()}
    mask = _masks(batch, len(names), gen)
    labels = torch.randint(0, classes, (batch,), generator=gen)
    payload = {
        "names": np.array(names),
        "heads": np.int64(heads),
        "classes": np.int64(classes),
        "hidden": np.int64(hidden),
        "drop_p": np.float64(drop_p),
        "smoothing": np.float64(smoothing),
        "mask": _np(mask),
        "labels": _np(labels),
    }
    for k, v in model.state_dict().items():
        payload["sd/" + k] = _np(v)
    for m in names:
        payload["x/" + m] = _np(feats[m])

    # ---- eval forward (fusion.py:331-427) --------------------------------
    model.eval()
    with torch.no_grad():
        logits, info = model(feats, mask, return_attention=True)
        logits_nomask = model(feats)
    payload["eval/logits"] = _np(logits)
    payload["eval/logits_nomask"] = _np(logits_nomask)
    payload["eval/fusion_weights"] = _np(info["fusion_weights"])
    for k, v in info["attention_maps"].items():
        payload["eval/attn/" + k] = _np(v)
    conf, pred = torch.max(F.softmax(logits, dim=1), dim=1)  # eval.py:89-90
    payload["eval/conf"] = _np(conf)
    payload["eval/pred"] = _np(pred)

    # ---- train forward + backward ---------------------------------------
    model.train()
    if drop_p > 0:
        # fusion.py:373 self.dropout (one call per modality), :294 projections[m][2],
        # attention.py:130 per pair (q-major order, present pairs only), fusion.py:326
        model.dropout = _QueueDrop()
        for m in names:
            d = _draw((batch, dims[m]), drop_p, gen)
            payload["drop/input/" + m] = _np(d)
            model.dropout.queue.append(d)
            qd = _QueueDrop()
            d = _draw((batch, hidden), drop_p, gen)
            payload["drop/proj/" + m] = _np(d)
            qd.queue.append(d)
            model.projections[m][2] = qd
        for q, k in itertools.permutations(names, 2):
            key = f"{q}_to_{k}"
            if key not in model.attention_modules:
                continue
            qd = _QueueDrop()
            d = _draw((batch, heads, 1, 1), drop_p, gen)
            payload["drop/attn/" + key] = _np(d)
            qd.queue.append(d)
            model.attention_modules[key].dropout = qd
        qd = _QueueDrop()
        d = _draw((batch, hidden), drop_p, gen)
        payload["drop/cls"] = _np(d)
        qd.queue.append(d)
        model.classifier[2] = qd
        model.train()

    xs = {m: feats[m].clone().requires_grad_(True) for m in names}
    logits, info = model(xs, mask, return_attention=True)
    loss = F.cross_entropy(logits, labels, label_smoothing=smoothing)  # train.py:185-186,310
    loss.backward()
    payload["train/logits"] = _np(logits)
    payload["train/loss"] = _np(loss)
    payload["train/fusion_weights"] = _np(info["fusion_weights"])
    for k, v in info["attention_maps"].items():
        payload["train/attn/" + k] = _np(v)
    for k, p in model.named_parameters():
        assert p.grad is not None, k
        payload["grad/" + k] = _np(p.grad)
    for m in names:
        payload["gradx/" + m] = _np(xs[m].grad)

    if optimizer:
        # train.py:378-382 AdamW(lr 1e-3, wd 1e-4) + clip norm 1.0 (base.yaml:74)
        total = torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        payload["opt/grad_norm"] = _np(total)
        opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=1e-4)
        opt.step()
        for k, p in model.named_parameters():
            payload["opt/" + k] = _np(p)
    _save(name, payload)


def fusion_seeded_case(ref_fusion, name, dims, hidden, heads, classes, batch, seed, smoothing=0.05):
    """Full-size shapes (BASELINE configs[1] / configs[4]): the state dict is too large for a fixture (13 MB / 240 MB),
    so the fixture holds the construction seed instead -- `torch.manual_seed(seed)` followed by the constructor gives
    the drop-in module the same parameters (same registration order and initialisers; the fixture's per-parameter sums
    let a test check that) -- plus the inputs, the reference's outputs and compact views of its gradients
    (norm, sum and the first 64 entries of every parameter gradient; the input gradients in full)."""
    torch.manual_seed(seed)
    gen = torch.Generator().manual_seed(seed + 1)
    names = list(dims)
    model = ref_fusion.HybridFusion(dims, hidden_dim=hidden, num_classes=classes, num_heads=heads, dropout=0.0)
    feats = {m: torch.randn(batch, d, generator=gen) for m, d in dims.items()}
    mask = _masks(batch, len(names), gen, with_float=False)
    labels = torch.randint(0, classes, (batch,), generator=gen)
    payload = {
        "names": np.array(names), "dims": np.array([dims[m] for m in names], dtype=np.int64),
        "heads": np.int64(heads), "classes": np.int64(classes), "hidden": np.int64(hidden),
        "seed": np.int64(seed), "smoothing": np.float64(smoothing), "mask": _np(mask), "labels": _np(labels),
    }
    for m in names:
        payload["x/" + m] = _np(feats[m])
    for k, v in model.state_dict().items():
        payload["sdsum/" + k] = np.float64(v.double().sum().item())
    model.eval()
    with torch.no_grad():
        logits, info = model(feats, mask, return_attention=True)
    payload["eval/logits"] = _np(logits)
    payload["eval/fusion_weights"] = _np(info["fusion_weights"])
    payload["eval/attn_stack"] = _np(torch.stack([info["attention_maps"][k].reshape(batch, heads)
                                                  for k in sorted(info["attention_maps"])]))
    payload["eval/attn_keys"] = np.array(sorted(info["attention_maps"]))
    model.train()
    xs = {m: feats[m].clone().requires_grad_(True) for m in names}
    logits = model(xs, mask)
    loss = F.cross_entropy(logits, labels, label_smoothing=smoothing)
    loss.backward()
    payload["train/loss"] = _np(loss)
    for k, p in model.named_parameters():
        g = p.grad.reshape(-1)
        payload["gnorm/" + k] = np.float64(g.double().norm().item())
        payload["gsum/" + k] = np.float64(g.double().sum().item())
        payload["ghead/" + k] = _np(g[:64])
    for m in names:
        payload["gradx/" + m] = _np(xs[m].grad)
    _save(name, payload)


def attention_case(ref_attention, name, seed):
    """Generic CrossModalAttention (attention.py:68-146), q_len/k_len > 1."""
    torch.manual_seed(seed)
    gen = torch.Generator().manual_seed(seed + 1)
    b, ql, kl, qd, kd, hid, heads = 5, 3, 6, 24, 12, 32, 4
    att = ref_attention.CrossModalAttention(qd, kd, hidden_dim=hid, num_heads=heads, dropout=0.0)
    att.eval()
    q3 = torch.randn(b, ql, qd, generator=gen, requires_grad=True)
    k3 = torch.randn(b, kl, kd, generator=gen, requires_grad=True)
    v3 = torch.randn(b, kl, kd, generator=gen, requires_grad=True)
    m2 = (torch.rand(b, kl, generator=gen) < 0.6).float()
    m2[0] = 0.0  # a fully masked row -> NaN softmax -> 0 (attention.py:127-129)
    out3, w3 = att(q3, k3, v3, m2)
    out3.square().sum().backward()
    payload = {"heads": np.int64(heads)}
    for k, v in att.state_dict().items():
        payload["sd/" + k] = _np(v)
    payload.update({
        "q3": _np(q3), "k3": _np(k3), "v3": _np(v3), "mask2": _np(m2),
        "out3": _np(out3), "w3": _np(w3),
        "gq3": _np(q3.grad), "gk3": _np(k3.grad), "gv3": _np(v3.grad),
    })
    for k, p in att.named_parameters():
        payload["grad3/" + k] = _np(p.grad)
    att.zero_grad()
    # 2-D (HybridFusion-style) call with a 1-D key mask
    q2 = torch.randn(b, qd, generator=gen, requires_grad=True)
    k2 = torch.randn(b, kd, generator=gen, requires_grad=True)
    v2 = torch.randn(b, kd, generator=gen, requires_grad=True)
    m1 = torch.tensor([1.0, 1.0, 0.0, 1.0, 0.0])
    out2, w2 = att(q2, k2, v2, m1)
    out2.sum().backward()
    payload.update({
        "q2": _np(q2), "k2": _np(k2), "v2": _np(v2), "mask1": _np(m1),
        "out2": _np(out2), "w2": _np(w2),
        "gq2": _np(q2.grad), "gk2": _np(k2.grad), "gv2": _np(v2.grad),
    })
    for k, p in att.named_parameters():
        payload["grad2/" + k] = _np(p.grad)
    _save(name, payload)


def ece_case(ref_uncertainty, name, seed, n, classes):
    """ECE/MCE floats from the real reference (uncertainty.py:84-171) plus the
    per-bin counts obtained by evaluating the reference's own mask expressions
    (uncertainty.py:113-117) with its ``torch.linspace`` edges."""
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(n, classes, generator=g) * 2
    labels = torch.randint(0, classes, (n,), generator=g)
    conf, pred = torch.max(F.softmax(logits, dim=1), dim=1)
    # adversarial tail: exact fp32/f64 edges, 0, 1, NaN, out of range
    extra = [0.0, 1.0, float("nan"), 1.0000001, -1e-9, 0.5, 0.2, 0.6, 1.0 / 3.0]
    for nb in (15, 10, 2):
        extra += torch.linspace(0.0, 1.0, nb + 1).tolist()
        extra += np.linspace(0.0, 1.0, nb + 1).astype(np.float32).tolist()
    extra_t = torch.tensor(extra, dtype=torch.float32)
    conf = torch.cat([conf, extra_t])
    pred = torch.cat([pred, torch.arange(extra_t.numel()) % classes])
    labels = torch.cat([labels, (torch.arange(extra_t.numel()) * 7) % classes])
    payload = {"conf": _np(conf), "pred": _np(pred).astype(np.int16),
               "label": _np(labels).astype(np.int16)}
    CM = ref_uncertainty.CalibrationMetrics
    for nb in (15, 10, 2):
        payload[f"ece/{nb}"] = np.float64(CM.expected_calibration_error(conf, pred, labels, nb))
        payload[f"mce/{nb}"] = np.float64(CM.maximum_calibration_error(conf, pred, labels, nb))
        bounds = torch.linspace(0.0, 1.0, steps=nb + 1)  # uncertainty.py:109
        payload[f"edges_f32/{nb}"] = _np(bounds)
        cnt, cor, csum = [], [], []
        for lower, upper in zip(bounds[:-1], bounds[1:]):
            if upper == 1.0:
                in_bin = (conf >= lower) & (conf <= upper)
            else:
                in_bin = (conf >= lower) & (conf < upper)
            cnt.append(int(in_bin.sum()))
            cor.append(int((pred[in_bin] == labels[in_bin]).sum()))
            csum.append(float(conf[in_bin].double().sum()))
        payload[f"count_f32/{nb}"] = np.array(cnt, dtype=np.int64)
        payload[f"correct_f32/{nb}"] = np.array(cor, dtype=np.int64)
        payload[f"confsum_f32/{nb}"] = np.array(csum, dtype=np.float64)
        # reliability-diagram binning, numpy float64 edges (uncertainty.py:222-241)
        c64 = conf.numpy()
        edges = np.linspace(0.0, 1.0, nb + 1)
        cnt = []
        for idx, (lo, up) in enumerate(zip(edges[:-1], edges[1:])):
            if idx == nb - 1:
                in_bin = (c64 >= lo) & (c64 <= up)
            else:
                in_bin = (c64 >= lo) & (c64 < up)
            cnt.append(int(np.sum(in_bin)))
        payload[f"count_f64/{nb}"] = np.array(cnt, dtype=np.int64)
    _save(name, payload)


def survey_kat(ref_uncertainty):
    """KATs quoted in SURVEY.md §8c, regenerated rather than trusted."""
    CM = ref_uncertainty.CalibrationMetrics
    conf = torch.tensor([0.8, 0.7]); pred = torch.tensor([0, 1]); lab = torch.tensor([0, 1])
    e2 = CM.expected_calibration_error(conf, pred, lab, num_bins=2)
    m2 = CM.maximum_calibration_error(conf, pred, lab, num_bins=2)
    g = torch.Generator().manual_seed(1234)
    logits = torch.randn(100000, 25, generator=g) * 2
    labels = torch.randint(0, 25, (100000,), generator=g)
    c, p = torch.max(F.softmax(logits, dim=1), dim=1)
    e15 = CM.expected_calibration_error(c, p, labels, 15)
    m15 = CM.maximum_calibration_error(c, p, labels, 15)
    print("KAT ece2", e2, "mce2", m2, "ece15", e15, "mce15", m15)
    _save("ece_kat.npz", {"ece2": np.float64(e2), "mce2": np.float64(m2),
                          "ece15_seed1234": np.float64(e15), "mce15_seed1234": np.float64(m15)})


def encoder_cases():
    """SequenceEncoder (lstm with and without lengths, gru) and SimpleMLPEncoder of the unmodified
    reference (src/encoders.py), eval-mode outputs plus dropout-free training gradients."""
    sys.path.insert(0, REF_SRC)
    import encoders as ref_enc  # noqa
    assert ref_enc.__file__.startswith(REF_SRC), ref_enc.__file__
    sys.path.pop(0)
    gen = torch.Generator().manual_seed(31)
    payload = {}
    B, T, F_in, H, D = 6, 14, 17, 32, 16
    x = torch.randn(B, T, F_in, generator=gen) * 2.0
    lengths = torch.tensor([14, 9, 1, 14, 5, 12])
    payload["seq/x"], payload["seq/lengths"] = _np(x), _np(lengths)
    for kind in ("lstm", "gru"):
        torch.manual_seed(41 if kind == "lstm" else 42)
        enc = ref_enc.SequenceEncoder(F_in, hidden_dim=H, output_dim=D, num_layers=2, encoder_type=kind, dropout=0.0)
        enc.eval()
        for k, v in enc.state_dict().items():
            payload[f"{kind}/sd/{k}"] = _np(v)
        payload[f"{kind}/out"] = _np(enc(x))
        payload[f"{kind}/out_lengths"] = _np(enc(x, lengths))   # ragged windows: the reference packs them (encoders.py:140-156)
        enc.train()
        xg = x.clone().requires_grad_(True)
        out = enc(xg)
        w = torch.linspace(-1, 1, D).unsqueeze(0)
        (out * w).sum().backward()
        payload[f"{kind}/gradx"] = _np(xg.grad)
        for k, p in enc.named_parameters():
            payload[f"{kind}/grad/{k}"] = _np(p.grad)
    # SimpleMLPEncoder: heart-rate style features, batch-norm batch statistics in train mode
    torch.manual_seed(43)
    mlp = ref_enc.SimpleMLPEncoder(12, hidden_dim=24, output_dim=D, num_layers=2, dropout=0.0, batch_norm=True)
    xm = torch.randn(10, 12, generator=gen)
    payload["mlp/x"] = _np(xm)
    for k, v in mlp.state_dict().items():
        payload[f"mlp/sd/{k}"] = _np(v)
    mlp.eval()
    payload["mlp/out_eval"] = _np(mlp(xm))
    mlp.train()
    xg = xm.clone().requires_grad_(True)
    out = mlp(xg)
    payload["mlp/out_train"] = _np(out)
    (out * torch.linspace(-1, 1, D).unsqueeze(0)).sum().backward()
    payload["mlp/gradx"] = _np(xg.grad)
    for k, p in mlp.named_parameters():
        payload[f"mlp/grad/{k}"] = _np(p.grad)
    # LayerNorm glue of train.py:170-171,267-268 on an encoder output
    torch.manual_seed(44)
    ln = nn.LayerNorm(D)
    with torch.no_grad():
        ln.weight.uniform_(0.5, 1.5)
        ln.bias.uniform_(-0.2, 0.2)
    payload["ln/weight"], payload["ln/bias"] = _np(ln.weight), _np(ln.bias)
    payload["ln/out"] = _np(ln(torch.from_numpy(payload["lstm/out"])))
    _save("encoders_small.npz", payload)


def frame_and_temporal_cases():
    """FrameEncoder (all three poolings, frame masks incl. a fully masked clip) of src/encoders.py:211-336 and
    TemporalAttention / PairwiseModalityAttention of src/attention.py:149-281,284-372 of the unmodified reference:
    eval-mode outputs and dropout-free training gradients.  A file of its own (frame_temporal_small.npz) so the older
    fixtures regenerate bit-identically."""
    sys.path.insert(0, REF_SRC)
    import attention as ref_att  # noqa
    import encoders as ref_enc  # noqa
    assert ref_enc.__file__.startswith(REF_SRC) and ref_att.__file__.startswith(REF_SRC)
    sys.path.pop(0)
    gen = torch.Generator().manual_seed(51)
    payload = {}
    B, T, F_in, H, D = 5, 9, 24, 32, 16
    frames = torch.randn(B, T, F_in, generator=gen)
    mask = torch.ones(B, T)
    mask[1, 4:] = 0
    mask[2, 0] = 0
    mask[3, :] = 0          # a clip with no valid frame
    payload["frame/x"], payload["frame/mask"] = _np(frames), _np(mask)
    w = torch.linspace(-1, 1, D).unsqueeze(0)
    for pool in ("attention", "average", "max"):
        torch.manual_seed(52)
        enc = ref_enc.FrameEncoder(F_in, hidden_dim=H, output_dim=D, temporal_pooling=pool, dropout=0.0)
        enc.eval()
        for k, v in enc.state_dict().items():
            payload[f"frame/{pool}/sd/{k}"] = _np(v)
        payload[f"frame/{pool}/out"] = _np(enc(frames))
        payload[f"frame/{pool}/out_mask"] = _np(enc(frames, mask))
        enc.train()
        xg = frames.clone().requires_grad_(True)
        (enc(xg, mask) * w).sum().backward()
        payload[f"frame/{pool}/gradx"] = _np(xg.grad)
        for k, prm in enc.named_parameters():
            payload[f"frame/{pool}/grad/{k}"] = _np(prm.grad)
    # TemporalAttention: self-attention over the time steps of one modality
    torch.manual_seed(53)
    ta = ref_att.TemporalAttention(F_in, hidden_dim=H, num_heads=4, dropout=0.0)
    ta.eval()
    for k, v in ta.state_dict().items():
        payload[f"temporal/sd/{k}"] = _np(v)
    tmask = mask.clone()
    tmask[3, :2] = 1        # the reference's softmax over an all-masked row is NaN-cleaned; keep one such row out
    payload["temporal/mask"] = _np(tmask)
    out, wts = ta(frames)
    payload["temporal/out"], payload["temporal/weights"] = _np(out), _np(wts)
    out_m, wts_m = ta(frames, tmask)
    payload["temporal/out_mask"], payload["temporal/weights_mask"] = _np(out_m), _np(wts_m)
    payload["temporal/pooled"] = _np(ta.pool_sequence(frames, wts_m))
    ta.train()
    xg = frames.clone().requires_grad_(True)
    o, _ = ta(xg, tmask)
    (o * torch.linspace(-1, 1, H).view(1, 1, H)).sum().backward()
    payload["temporal/gradx"] = _np(xg.grad)
    for k, prm in ta.named_parameters():
        payload[f"temporal/grad/{k}"] = _np(prm.grad)
    # PairwiseModalityAttention over three modality embeddings
    dims = {"video": 12, "imu": 20, "hr": 8}
    torch.manual_seed(54)
    pa = ref_att.PairwiseModalityAttention(dims, hidden_dim=H, num_heads=4, dropout=0.0)
    pa.eval()
    for k, v in pa.state_dict().items():
        payload[f"pairwise/sd/{k}"] = _np(v)
    feats = {m: torch.randn(B, d, generator=gen) for m, d in dims.items()}
    mmask = torch.tensor([[1, 1, 1], [1, 0, 1], [0, 1, 1], [1, 1, 0], [1, 0, 0]], dtype=torch.float32)
    for m, t in feats.items():
        payload[f"pairwise/x/{m}"] = _np(t)
    payload["pairwise/mask"] = _np(mmask)
    got = pa(feats, mmask)
    attended, maps = got if isinstance(got, tuple) else (got, {})
    for m, t in attended.items():
        payload[f"pairwise/out/{m}"] = _np(t)
    for key, t in maps.items():
        payload[f"pairwise/map/{key}"] = _np(t)
    _save("frame_temporal_small.npz", payload)


def main():
    ref_fusion, ref_attention, ref_uncertainty = _import_reference()
    # the reference's own test configuration (tests/test_fusion.py:50-80)
    fusion_case(ref_fusion, "fusion_tiny.npz", {"video": 4, "imu": 4}, 8, 1, 3, 6, seed=0)
    pamap = {"imu_hand": 16, "imu_chest": 16, "imu_ankle": 16, "heart_rate": 16}
    fusion_case(ref_fusion, "fusion_pamap_small.npz", pamap, 32, 4, 25, 40, seed=11, optimizer=True)
    fusion_case(ref_fusion, "fusion_pamap_dropout.npz", pamap, 32, 4, 25, 40, seed=12, drop_p=0.1)
    fusion_case(ref_fusion, "fusion_missing_pair.npz", {"a": 8, "b": 12, "c": 4}, 16, 2, 5, 10,
                seed=13, delete_pairs=("a_to_b", "c_to_a"))
    # tensor-core eligible shape (hidden % 64 == 0), ragged batch vs a 128-row tile
    fusion_case(ref_fusion, "fusion_tc_shape.npz", {"imu": 64, "hr": 64}, 64, 4, 25, 130, seed=14)
    # full-size shapes by construction seed: BASELINE configs[1] (PAMAP2) and configs[4] (scaled variant)
    pamap_full = {"imu_hand": 128, "imu_chest": 128, "imu_ankle": 128, "heart_rate": 128}
    fusion_seeded_case(ref_fusion, "fusion_config2_seeded.npz", pamap_full, 256, 4, 25, 48, seed=31)
    scaled = {f"video_{i}": 256 for i in range(2)}
    scaled.update({f"imu_{i}": 256 for i in range(6)})
    fusion_seeded_case(ref_fusion, "fusion_config5_seeded.npz", scaled, 512, 8, 11, 24, seed=32)
    attention_case(ref_attention, "attention_generic.npz", seed=21)
    ece_case(ref_uncertainty, "ece_seeded.npz", seed=1234, n=20000, classes=25)
    survey_kat(ref_uncertainty)
    encoder_cases()
    frame_and_temporal_cases()


if __name__ == "__main__":
    main()
