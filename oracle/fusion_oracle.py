"""CPU oracle: functional restatement of the reference HybridFusion path.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Plain torch CPU ops in
fp32 or fp64 over an explicit ``state_dict``; no nn.Module, no CUDA.  Every
function cites the reference lines it restates (paths relative to
``/root/reference``).  Pinned against the unmodified reference by
``tests/golden/fusion_*.npz`` (made by ``oracle/make_golden.py``).

Dropout is injectable: the reference draws CPU ``bernoulli_`` streams that no
GPU can reproduce, so every dropout site takes an explicit multiplicative mask
(values ``0`` or ``1/(1-p)``); ``None`` means eval mode / ``p = 0``.
"""

from __future__ import annotations

import math
from typing import Dict, List, Mapping, Optional, Sequence, Tuple

import torch

StateDict = Mapping[str, torch.Tensor]


def _linear(sd: StateDict, prefix: str, x: torch.Tensor) -> torch.Tensor:
    """``nn.Linear``: ``x @ W.T + b`` with W of shape (out, in)."""
    w = sd[prefix + ".weight"].to(x.dtype)
    b = sd[prefix + ".bias"].to(x.dtype)
    return x @ w.t() + b


# ---------------------------------------------------------------------------
# CrossModalAttention  (src/attention.py:68-146)
# ---------------------------------------------------------------------------
def cross_modal_attention(
    sd: StateDict,
    prefix: str,
    query: torch.Tensor,
    key: torch.Tensor,
    value: torch.Tensor,
    num_heads: int,
    mask: Optional[torch.Tensor] = None,
    attn_drop: Optional[torch.Tensor] = None,
) -> Tuple[torch.Tensor, torch.Tensor]:
    """Literal restatement, generic over q_len / k_len (attention.py:88-146).

    ``attn_drop`` is the dropout multiplier applied to the attention weights
    (attention.py:130), broadcastable to ``(B, heads, q_len, k_len)``.
    """
    batch = query.size(0)
    squeeze_query = squeeze_key = False
    if query.dim() == 2:  # attention.py:92-94
        query, squeeze_query = query.unsqueeze(1), True
    if key.dim() == 2:  # attention.py:95-97
        key, squeeze_key = key.unsqueeze(1), True
    if value.dim() == 2:  # attention.py:98-99
        value = value.unsqueeze(1)
    q_len, k_len = query.size(1), key.size(1)

    hidden = sd[prefix + ".query_proj.weight"].shape[0]
    head_dim = hidden // num_heads
    scale = head_dim ** -0.5  # attention.py:66

    q = _linear(sd, prefix + ".query_proj", query)  # attention.py:104
    k = _linear(sd, prefix + ".key_proj", key)  # attention.py:105
    v = _linear(sd, prefix + ".value_proj", value)  # attention.py:106
    q = q.view(batch, q_len, num_heads, head_dim).transpose(1, 2)
    k = k.view(batch, k_len, num_heads, head_dim).transpose(1, 2)
    v = v.view(batch, k_len, num_heads, head_dim).transpose(1, 2)

    scores = torch.matmul(q, k.transpose(-2, -1)) * scale  # attention.py:118
    if mask is not None:  # attention.py:120-124
        if mask.dim() == 1:
            mask = mask.unsqueeze(1)
        mask = mask.unsqueeze(1).unsqueeze(2)
        scores = scores.masked_fill(mask == 0, float("-inf"))
    weights = torch.softmax(scores, dim=-1)  # attention.py:126
    weights = torch.nan_to_num(weights, nan=0.0, posinf=0.0, neginf=0.0)
    if attn_drop is not None:  # attention.py:130
        weights = weights * attn_drop.to(weights.dtype)

    attended = torch.matmul(weights, v)  # attention.py:132-134
    attended = attended.transpose(1, 2).contiguous().view(batch, q_len, hidden)
    attended = _linear(sd, prefix + ".out_proj", attended)  # attention.py:140
    if squeeze_query:
        attended = attended.squeeze(1)
    if squeeze_key:
        weights = weights[:, :, :, :1]
    return attended, weights


# ---------------------------------------------------------------------------
# HybridFusion.compute_adaptive_weights  (src/fusion.py:429-479)
# ---------------------------------------------------------------------------
def adaptive_weights(
    sd: StateDict,
    modality_names: Sequence[str],
    aggregated: Mapping[str, torch.Tensor],
    mask: torch.Tensor,
) -> torch.Tensor:
    scores = [
        _linear(sd, f"gating_layers.{m}", aggregated[m]) for m in modality_names
    ]  # fusion.py:452-459
    score = torch.cat(scores, dim=1)  # fusion.py:461
    mask = mask.to(score.dtype)
    masked = score.masked_fill(mask <= 0, float("-inf"))  # fusion.py:464
    w = torch.softmax(masked, dim=1)  # fusion.py:465
    w = torch.nan_to_num(w, nan=0.0, posinf=0.0, neginf=0.0)  # fusion.py:466
    w = w * mask  # fusion.py:467
    sum_w = w.sum(dim=1, keepdim=True)
    mask_sum = mask.sum(dim=1, keepdim=True)
    fallback = torch.where(  # fusion.py:471-475
        mask_sum > 0,
        mask / (mask_sum + 1e-8),
        torch.full_like(mask, 1.0 / len(modality_names)),
    )
    return torch.where(sum_w > 0, w / (sum_w + 1e-8), fallback)  # fusion.py:476-478


# ---------------------------------------------------------------------------
# HybridFusion.forward  (src/fusion.py:331-427)
# ---------------------------------------------------------------------------
def hybrid_fusion_forward(
    sd: StateDict,
    modality_names: Sequence[str],
    num_heads: int,
    features: Mapping[str, torch.Tensor],
    mask: Optional[torch.Tensor] = None,
    drops: Optional[Mapping[str, object]] = None,
) -> Tuple[torch.Tensor, Dict[str, object]]:
    """Returns ``(logits, {"attention_maps", "fusion_weights", "aggregated"})``.

    ``drops`` (train mode) may hold ``"input"``: {m: (B, D_m)}, ``"proj"``:
    {m: (B, H)}, ``"attn"``: {"q_to_k": (B, heads, 1, 1)}, ``"cls"``: (B, H).
    """
    drops = drops or {}
    names = list(modality_names)
    ref = features[names[0]]
    batch, dtype = ref.size(0), ref.dtype
    if mask is None:  # fusion.py:357-362
        mask = torch.ones(batch, len(names), dtype=dtype)
    else:
        mask = mask.to(dtype=dtype)

    projected: Dict[str, torch.Tensor] = {}
    for idx, m in enumerate(names):  # fusion.py:365-374
        x = features[m] * mask[:, idx].unsqueeze(-1)
        if "input" in drops:
            x = x * drops["input"][m].to(dtype)
        p = torch.relu(_linear(sd, f"projections.{m}.0", x))
        if "proj" in drops:
            p = p * drops["proj"][m].to(dtype)
        projected[m] = p

    lists = {m: [projected[m]] for m in names}
    attention_maps: Dict[str, torch.Tensor] = {}
    for q in names:  # fusion.py:383-404
        for k_idx, k in enumerate(names):
            if q == k:
                continue
            key = f"{q}_to_{k}"
            if f"attention_modules.{key}.value_proj.weight" not in sd:
                continue  # fusion.py:388-389
            attn_drop = drops["attn"][key] if "attn" in drops else None
            attended, w = cross_modal_attention(
                sd,
                f"attention_modules.{key}",
                projected[q],
                projected[k],
                projected[k],
                num_heads,
                mask=mask[:, k_idx],
                attn_drop=attn_drop,
            )
            lists[q].append(attended)
            attention_maps[key] = w

    aggregated = {}
    for idx, m in enumerate(names):  # fusion.py:406-408
        stacked = torch.stack(lists[m], dim=0).mean(dim=0)
        aggregated[m] = stacked * mask[:, idx].unsqueeze(-1)

    fw = adaptive_weights(sd, names, aggregated, mask)  # fusion.py:410-412
    modality_tensor = torch.stack([aggregated[m] for m in names], dim=1)
    fused = (modality_tensor * fw.unsqueeze(-1)).sum(dim=1)  # fusion.py:416-418
    hidden = torch.relu(_linear(sd, "classifier.0", fused))  # fusion.py:419
    if "cls" in drops:
        hidden = hidden * drops["cls"].to(dtype)
    logits = _linear(sd, "classifier.3", hidden)
    return logits, {
        "attention_maps": attention_maps,
        "fusion_weights": fw,
        "aggregated": aggregated,
        "fused": fused,
    }


def hybrid_fusion_closed_form(
    sd: StateDict,
    modality_names: Sequence[str],
    num_heads: int,
    features: Mapping[str, torch.Tensor],
    mask: torch.Tensor,
) -> torch.Tensor:
    """Eval-mode closed form of SURVEY.md §8 a-2: with q_len = k_len = 1 the
    softmax over one key is 1 (or NaN -> 0 when the key is masked), so the
    query/key projections drop out and attention is the gate ``1[mask_k != 0]``.
    Used only to cross-check the literal restatement above."""
    names = list(modality_names)
    dtype = features[names[0]].dtype
    mask = mask.to(dtype)
    proj = {
        m: torch.relu(
            _linear(sd, f"projections.{m}.0", features[m] * mask[:, i : i + 1])
        )
        for i, m in enumerate(names)
    }
    agg = {}
    for qi, q in enumerate(names):
        total, count = proj[q], 1
        for ki, k in enumerate(names):
            pre = f"attention_modules.{q}_to_{k}"
            if q == k or pre + ".value_proj.weight" not in sd:
                continue
            gate = (mask[:, ki : ki + 1] != 0).to(dtype)
            v = _linear(sd, pre + ".value_proj", proj[k]) * gate
            total = total + _linear(sd, pre + ".out_proj", v)
            count += 1
        agg[q] = total / count * mask[:, qi : qi + 1]
    fw = adaptive_weights(sd, names, agg, mask)
    fused = sum(agg[m] * fw[:, i : i + 1] for i, m in enumerate(names))
    hidden = torch.relu(_linear(sd, "classifier.0", fused))
    return _linear(sd, "classifier.3", hidden)


# ---------------------------------------------------------------------------
# Loss / confidence / optimizer pieces adjacent to the path
# ---------------------------------------------------------------------------
def cross_entropy_label_smoothing(
    logits: torch.Tensor, labels: torch.Tensor, smoothing: float = 0.0
) -> torch.Tensor:
    """``nn.CrossEntropyLoss(label_smoothing=s)`` mean reduction
    (src/train.py:185-186,310): ``(1-s)·nll + s·mean_c(-log p_c)``."""
    logp = torch.log_softmax(logits, dim=1)
    nll = -logp.gather(1, labels.view(-1, 1).long()).squeeze(1)
    smooth = -logp.mean(dim=1)
    return ((1.0 - smoothing) * nll + smoothing * smooth).mean()


def softmax_conf_pred(logits: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """``probs = softmax(logits, 1); conf, pred = max(probs, 1)``
    (src/eval.py:89-90, src/train.py:341-342); first index wins ties."""
    probs = torch.softmax(logits, dim=1)
    conf, pred = torch.max(probs, dim=1)
    return conf, pred


def clip_grad_norm(grads: List[torch.Tensor], max_norm: float) -> float:
    """``torch.nn.utils.clip_grad_norm_`` as Lightning applies it
    (src/train.py:416-430, ``gradient_clip_norm`` base.yaml:74). In place."""
    total = math.sqrt(sum(float((g.double() ** 2).sum()) for g in grads))
    coef = min(1.0, max_norm / (total + 1e-6))
    for g in grads:
        g.mul_(coef)
    return total


def adamw_step(
    p: torch.Tensor,
    g: torch.Tensor,
    m: torch.Tensor,
    v: torch.Tensor,
    step: int,
    lr: float = 1e-3,
    beta1: float = 0.9,
    beta2: float = 0.999,
    eps: float = 1e-8,
    weight_decay: float = 1e-4,
) -> None:
    """``torch.optim.AdamW`` single-tensor update (src/train.py:378-382;
    lr 1e-3, wd 1e-4 from config/base.yaml). ``step`` is 1-based. In place."""
    p.mul_(1.0 - lr * weight_decay)
    m.mul_(beta1).add_(g, alpha=1.0 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1.0 - beta2)
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-(lr / bc1))
