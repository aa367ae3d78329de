"""Attention modules — drop-in for the reference's ``src/attention.py``.

``CrossModalAttention`` keeps the reference's constructor, sub-module names
(``query_proj / key_proj / value_proj / out_proj / dropout``), attributes
(``hidden_dim, num_heads, head_dim, scale``) and return convention
(attention.py:16-146); its forward runs on the msf_b200 kernels: the four
projections through ``msf_linear_*`` and the score/softmax/weights.V core
through ``msf_attention_core_*`` (generic over q_len and k_len).

Inside ``HybridFusion`` these modules are parameter containers only: there the
q_len = k_len = 1 attention degenerates to a per-(window, head) gate that the
fused kernels apply directly (see fusion.py and DESIGN.md).

``TemporalAttention`` and ``PairwiseModalityAttention`` (attention.py:149-413)
are not on the path ``north_star`` names (nothing in train.py / eval.py calls
them); they are provided so that the module's public surface is complete, and
they run on the same kernels: the projections through ``msf_linear_*``, the
attention core through ``msf_attention_core_*``, ``PairwiseModalityAttention``
by composing ``CrossModalAttention``.  ``visualize_attention`` is host plotting.

Device policy as in fusion.py: CUDA runs in place, CPU tensors are staged
through the current CUDA device, no CUDA device -> error.
"""
from __future__ import annotations

import importlib
import os
import sys
from pathlib import Path
from typing import Dict, Mapping, Optional, Sequence, Tuple

import torch
import torch.nn as nn

_HERE = os.path.dirname(os.path.realpath(__file__))   # realpath: src/ may be reached through a symlink
_ROOT = os.path.dirname(os.path.dirname(_HERE))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
ops = importlib.import_module(os.path.basename(os.path.dirname(_HERE)) + ".ops")


def _as_3d(t: torch.Tensor) -> Tuple[torch.Tensor, bool]:
    return (t.unsqueeze(1), True) if t.dim() == 2 else (t, False)


class CrossModalAttention(nn.Module):
    """Modality A (query) attends to modality B (key/value)."""

    def __init__(self, query_dim: int, key_dim: int, hidden_dim: int = 256, num_heads: int = 4,
                 dropout: float = 0.1):
        super().__init__()
        self.hidden_dim = hidden_dim
        self.num_heads = num_heads
        self.head_dim = hidden_dim // num_heads  # computed before the check, like attention.py:51,57
        assert hidden_dim % num_heads == 0, (
            f"hidden_dim ({hidden_dim}) must be divisible by num_heads ({num_heads})")
        self.query_proj = nn.Linear(query_dim, hidden_dim)
        self.key_proj = nn.Linear(key_dim, hidden_dim)
        self.value_proj = nn.Linear(key_dim, hidden_dim)
        self.out_proj = nn.Linear(hidden_dim, hidden_dim)
        self.dropout = nn.Dropout(dropout)
        self.scale = self.head_dim ** -0.5

    def forward(self, query: torch.Tensor, key: torch.Tensor, value: torch.Tensor,
                mask: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """Returns ``(attended, attention_weights)``; 2-D inputs are treated as
        length-1 sequences and the outputs squeezed back (attention.py:92-99,142-145)."""
        home, out_dtype = query.device, query.dtype
        dev = home if home.type == "cuda" else ops.require_cuda("CrossModalAttention.forward")
        query, squeeze_q = _as_3d(query)
        key, squeeze_k = _as_3d(key)
        value, _ = _as_3d(value)
        batch, k_len = query.size(0), key.size(1)

        def put(t):
            return t.to(device=dev, dtype=torch.float32)

        key_mask = None
        if mask is not None:  # (B,) -> (B, 1); (B, k_len) stays (attention.py:120-123)
            key_mask = mask.unsqueeze(1) if mask.dim() == 1 else mask
            key_mask = key_mask.expand(batch, k_len)
        p = float(self.dropout.p)
        training = bool(self.training and p > 0.0)
        seed = int(torch.randint(0, 2**62, (1,)).item()) if training else 0
        with torch.cuda.device(dev):
            q = ops.linear(put(query), put(self.query_proj.weight), put(self.query_proj.bias))
            k = ops.linear(put(key), put(self.key_proj.weight), put(self.key_proj.bias))
            v = ops.linear(put(value), put(self.value_proj.weight), put(self.value_proj.bias))
            ctx, weights = ops.attention_core(q, k, v, key_mask, self.num_heads, p, training, seed)
            attended = ops.linear(ctx, put(self.out_proj.weight), put(self.out_proj.bias))
        attended = attended.to(device=home, dtype=out_dtype)
        weights = weights.to(device=home, dtype=out_dtype)
        if squeeze_q:
            attended = attended.squeeze(1)
        if squeeze_k:
            weights = weights[:, :, :, :1]
        return attended, weights


def _staged(dev):
    def put(t):
        return t.to(device=dev, dtype=torch.float32)
    return put


class TemporalAttention(nn.Module):
    """Multi-head self-attention over the time steps of one sequence (attention.py:149-281)."""

    def __init__(self, feature_dim: int, hidden_dim: int = 256, num_heads: int = 4, dropout: float = 0.1):
        super().__init__()
        self.feature_dim, self.hidden_dim, self.num_heads = feature_dim, hidden_dim, num_heads
        self.head_dim = hidden_dim // num_heads
        self.query_proj = nn.Linear(feature_dim, hidden_dim)
        self.key_proj = nn.Linear(feature_dim, hidden_dim)
        self.value_proj = nn.Linear(feature_dim, hidden_dim)
        self.out_proj = nn.Linear(hidden_dim, hidden_dim)
        self.dropout = nn.Dropout(dropout)
        self.scale = self.head_dim ** -0.5

    def forward(self, sequence: torch.Tensor, mask: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """``sequence`` (batch, steps, feature_dim), ``mask`` (batch, steps) or (steps,) of valid steps ->
        ``(attended, weights)`` with weights (batch, heads, steps, steps).  Masked keys get zero weight; a masked
        output is multiplied by the mask reshaped to (batch', 1, 1, steps, 1), exactly as attention.py:228-256
        does (so with a mask the attended tensor comes back broadcast to (batch', 1, batch, steps, hidden))."""
        home, out_dtype = sequence.device, sequence.dtype
        dev = home if home.type == "cuda" else ops.require_cuda("TemporalAttention.forward")
        put = _staged(dev)
        batch, steps, _ = sequence.shape
        key_mask = None
        if mask is not None:
            key_mask = mask.unsqueeze(0) if mask.dim() == 1 else mask
        p = float(self.dropout.p)
        training = bool(self.training and p > 0.0)
        seed = int(torch.randint(0, 2**62, (1,)).item()) if training else 0
        with torch.cuda.device(dev):
            x = put(sequence)
            q = ops.linear(x, put(self.query_proj.weight), put(self.query_proj.bias))
            k = ops.linear(x, put(self.key_proj.weight), put(self.key_proj.bias))
            v = ops.linear(x, put(self.value_proj.weight), put(self.value_proj.bias))
            core_mask = None if key_mask is None else put(key_mask).expand(batch, steps).contiguous()
            ctx, weights = ops.attention_core(q, k, v, core_mask, self.num_heads, p, training, seed)
            attended = ops.linear(ctx, put(self.out_proj.weight), put(self.out_proj.bias))
        attended = attended.to(device=home, dtype=out_dtype)
        weights = weights.to(device=home, dtype=out_dtype)
        if key_mask is not None:
            attended = attended * key_mask.to(attended.dtype)[:, None, None, :, None]
        return attended, weights

    def pool_sequence(self, sequence: torch.Tensor, attention_weights: torch.Tensor) -> torch.Tensor:
        """Fixed-size summary: the steps weighted by the attention they receive, averaged over heads and query
        positions and re-normalised (attention.py:258-281)."""
        if attention_weights.dim() != 4:
            raise ValueError(f"Expected attention weights with 4 dims, got {attention_weights.shape}")
        received = attention_weights.mean(dim=(1, 2))                       # (batch, steps)
        received = received / (received.sum(dim=1, keepdim=True) + 1e-8)
        return torch.einsum("bs,bsh->bh", received, sequence)


class PairwiseModalityAttention(nn.Module):
    """Every modality attends to every other one through its own ``CrossModalAttention`` and averages what it
    got with its own projection (attention.py:284-413) — HybridFusion's first half."""

    def __init__(self, modality_dims: Mapping[str, int], hidden_dim: int = 256, num_heads: int = 4,
                 dropout: float = 0.1):
        super().__init__()
        dims = dict(modality_dims)
        self.modality_names = list(dims)
        self.num_modalities = len(dims)
        self.hidden_dim = hidden_dim
        self.projections = nn.ModuleDict(
            {m: nn.Sequential(nn.Linear(d, hidden_dim), nn.ReLU(), nn.Dropout(dropout)) for m, d in dims.items()})
        self.attention_layers = nn.ModuleDict(
            {f"{q}_to_{k}": CrossModalAttention(hidden_dim, hidden_dim, hidden_dim, num_heads, dropout)
             for q in self.modality_names for k in self.modality_names if q != k})

    def forward(self, modality_features: Mapping[str, torch.Tensor], modality_mask: Optional[torch.Tensor] = None
                ) -> Tuple[Dict[str, torch.Tensor], Dict[str, torch.Tensor]]:
        if not self.modality_names:
            raise ValueError("No modalities provided for PairwiseModalityAttention.")
        first = modality_features[self.modality_names[0]]
        if modality_mask is None:
            modality_mask = torch.ones(first.size(0), self.num_modalities, device=first.device, dtype=first.dtype)
        else:
            modality_mask = modality_mask.to(device=first.device, dtype=first.dtype)
        tokens = {}
        for m in self.modality_names:
            lin, act, drop = self.projections[m]
            x = modality_features[m].to(first.device)
            if isinstance(lin, nn.Linear) and isinstance(act, nn.ReLU) and x.dim() == 2:
                dev = x.device if x.device.type == "cuda" else ops.require_cuda("PairwiseModalityAttention.forward")
                put = _staged(dev)
                with torch.cuda.device(dev):
                    y = ops.linear(put(x), put(lin.weight), put(lin.bias), relu=True)
                tokens[m] = drop(y.to(device=x.device, dtype=x.dtype))
            else:   # a swapped-in sub-module: compose whatever is there
                tokens[m] = self.projections[m](x)
        gathered = {m: [tokens[m]] for m in self.modality_names}
        maps: Dict[str, torch.Tensor] = {}
        for qi, q in enumerate(self.modality_names):
            for ki, k in enumerate(self.modality_names):
                name = f"{q}_to_{k}"
                if q == k or name not in self.attention_layers:
                    continue
                got, w = self.attention_layers[name](tokens[q], tokens[k], tokens[k], mask=modality_mask[:, ki])
                gathered[q].append(got)
                maps[name] = w
        out = {m: torch.stack(gathered[m]).mean(0) * modality_mask[:, i:i + 1]
               for i, m in enumerate(self.modality_names)}
        return out, maps


def visualize_attention(attention_weights, modality_names: Sequence[str], save_path: Path | str | None = None) -> None:
    """Heat map of query-modality x key-modality attention (attention.py:416-485): leading dimensions beyond
    two are averaged away, the figure is saved to ``save_path`` or shown."""
    import matplotlib.pyplot as plt
    import numpy as np

    grid = torch.as_tensor(attention_weights.detach() if isinstance(attention_weights, torch.Tensor)
                           else attention_weights, dtype=torch.float32).cpu()
    while grid.dim() < 2:
        grid = grid.unsqueeze(0)
    while grid.dim() > 2:
        grid = grid.mean(dim=0)
    values = np.atleast_2d(grid.numpy())    # a backend that hands back a squeezed vector still draws one row
    rows, cols = values.shape
    fig, ax = plt.subplots(figsize=(4 + 0.5 * cols, 4))
    image = ax.imshow(values, cmap="viridis", aspect="auto")
    ax.set_xticks(np.arange(cols))
    ax.set_yticks(np.arange(rows))
    ax.set_xticklabels(list(modality_names[:cols]), rotation=45, ha="right")
    ax.set_yticklabels(list(modality_names[:rows]))
    ax.set_xlabel("Key Modality")
    ax.set_ylabel("Query Modality")
    ax.set_title("Cross-Modal Attention Weights")
    plt.colorbar(image, ax=ax, fraction=0.046, pad=0.04)
    plt.tight_layout()
    if save_path is None:
        plt.show()
        return
    target = Path(save_path)
    target.parent.mkdir(parents=True, exist_ok=True)
    fig.savefig(target, dpi=300, bbox_inches="tight")
    plt.close(fig)


if __name__ == "__main__":
    # Simple test
    print("Testing attention mechanisms...")
    print("\nTesting CrossModalAttention...")
    try:
        layer = CrossModalAttention(query_dim=512, key_dim=64, hidden_dim=256, num_heads=4)
        a, b = torch.randn(4, 512), torch.randn(4, 64)
        out, w = layer(a, b, b)
        assert out.shape == (4, 256)
        print(f"✓ CrossModalAttention working! Output shape: {out.shape}")
    except NotImplementedError:
        print("✗ CrossModalAttention not implemented yet")
    except Exception as err:  # noqa: BLE001
        print(f"✗ CrossModalAttention error: {err}")
    print("\nTesting TemporalAttention...")
    try:
        steps = TemporalAttention(128, 256, 4)
        seq = torch.randn(4, 10, 128)
        attended_seq, w = steps(seq)
        assert attended_seq.shape == (4, 10, 256)
        print(f"✓ TemporalAttention working! Output shape: {attended_seq.shape}")
    except NotImplementedError:
        print("✗ TemporalAttention not implemented yet")
    except Exception as err:  # noqa: BLE001
        print(f"✗ TemporalAttention error: {err}")
