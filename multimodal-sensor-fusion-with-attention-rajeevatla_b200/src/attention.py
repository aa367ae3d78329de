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

Device policy as in fusion.py: CUDA runs in place, CPU tensors are staged
through the current CUDA device, no CUDA device -> error.
"""
from __future__ import annotations

import importlib
import os
import sys
from typing import Optional, Tuple

import torch
import torch.nn as nn

_HERE = os.path.dirname(os.path.abspath(__file__))
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


# Simple test
if __name__ == "__main__":
    print("Testing attention mechanisms...")
    try:
        layer = CrossModalAttention(query_dim=512, key_dim=64)
        a, b = torch.randn(4, 512), torch.randn(4, 64)
        out, w = layer(a, b, b)
        assert out.shape == (4, 256)
        print(f"✓ CrossModalAttention working! Output shape: {out.shape}")
    except NotImplementedError:
        print("✗ CrossModalAttention not implemented yet")
    except Exception as err:  # noqa: BLE001
        print(f"✗ CrossModalAttention error: {err}")
