"""CPU oracle: functional restatement of the reference encoders on the hot path.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Plain torch CPU ops over
an explicit ``state_dict``; every function cites the reference lines it restates
(paths relative to ``/root/reference``).  Pinned against the unmodified
reference by ``tests/golden/encoders_*.npz`` (``oracle/make_golden.py``).

The recurrences of ``nn.LSTM`` / ``nn.GRU`` live in PyTorch ATen (torch >= 2.9 per
the reference's ``pyproject.toml:26``; 2.11.0 here), not in the reference's
sources: they are restated from the published cell equations with PyTorch's
gate order (LSTM: i, f, g, o; GRU: r, z, n) and anchored on the reference's
call site ``src/encoders.py:135-166``.
"""
from __future__ import annotations

from typing import Mapping, Optional

import torch

StateDict = Mapping[str, torch.Tensor]


def lstm_last_hidden(sd: StateDict, prefix: str, x: torch.Tensor, num_layers: int,
                     lengths: Optional[torch.Tensor] = None,
                     layer_masks: Optional[Mapping[int, torch.Tensor]] = None) -> torch.Tensor:
    """``h_n[-1]`` of a ``batch_first`` multi-layer LSTM.  Eval mode by default; ``layer_masks[l]`` (B, T, H) are the
    multipliers (0 or 1/(1-p)) of nn.LSTM's training-mode dropout on the INPUT of layer l >= 1 (the outputs of layer
    l-1; encoders.py:54-65 passes ``dropout`` to nn.LSTM), injected so that both sides use the same draws.

    With ``lengths`` the state of row b stops updating after ``lengths[b]`` steps, which is what
    ``pack_padded_sequence`` does (encoders.py:141-156); encoders.py:160-164 takes ``hidden[0][-1]``.
    """
    B, T, _ = x.shape
    inp = x
    h = None
    for layer in range(num_layers):
        w_ih = sd[f"{prefix}.weight_ih_l{layer}"].to(x.dtype)
        w_hh = sd[f"{prefix}.weight_hh_l{layer}"].to(x.dtype)
        b = (sd[f"{prefix}.bias_ih_l{layer}"] + sd[f"{prefix}.bias_hh_l{layer}"]).to(x.dtype)
        H = w_hh.shape[1]
        h = torch.zeros(B, H, dtype=x.dtype)
        c = torch.zeros(B, H, dtype=x.dtype)
        outs = []
        for t in range(T):
            gates = inp[:, t] @ w_ih.t() + h @ w_hh.t() + b
            i, f, g, o = gates.split(H, dim=1)
            c_new = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
            h_new = torch.sigmoid(o) * torch.tanh(c_new)
            if lengths is not None:
                live = (t < lengths).to(x.dtype).unsqueeze(1)
                c_new = live * c_new + (1 - live) * c
                h_new = live * h_new + (1 - live) * h
            h, c = h_new, c_new
            outs.append(h)
        inp = torch.stack(outs, dim=1)
        if layer_masks is not None and (layer + 1) in layer_masks:
            inp = inp * layer_masks[layer + 1].to(x.dtype)
    return h


def gru_last_hidden(sd: StateDict, prefix: str, x: torch.Tensor, num_layers: int,
                    lengths: Optional[torch.Tensor] = None,
                    layer_masks: Optional[Mapping[int, torch.Tensor]] = None) -> torch.Tensor:
    """``h_n[-1]`` of a ``batch_first`` multi-layer GRU (gate order r, z, n); with ``lengths`` the state of row b
    stops after ``lengths[b]`` steps (pack_padded_sequence, encoders.py:141-156)."""
    B, T, _ = x.shape
    inp = x
    h = None
    for layer in range(num_layers):
        w_ih = sd[f"{prefix}.weight_ih_l{layer}"].to(x.dtype)
        w_hh = sd[f"{prefix}.weight_hh_l{layer}"].to(x.dtype)
        b_ih = sd[f"{prefix}.bias_ih_l{layer}"].to(x.dtype)
        b_hh = sd[f"{prefix}.bias_hh_l{layer}"].to(x.dtype)
        H = w_hh.shape[1]
        h = torch.zeros(B, H, dtype=x.dtype)
        outs = []
        for t in range(T):
            gi = inp[:, t] @ w_ih.t() + b_ih
            gh = h @ w_hh.t() + b_hh
            i_r, i_z, i_n = gi.split(H, dim=1)
            h_r, h_z, h_n = gh.split(H, dim=1)
            r = torch.sigmoid(i_r + h_r)
            z = torch.sigmoid(i_z + h_z)
            n = torch.tanh(i_n + r * h_n)
            h_new = (1 - z) * n + z * h
            if lengths is not None:
                live = (t < lengths).to(x.dtype).unsqueeze(1)
                h_new = live * h_new + (1 - live) * h
            h = h_new
            outs.append(h)
        inp = torch.stack(outs, dim=1)
        if layer_masks is not None and (layer + 1) in layer_masks:   # nn.GRU's training-mode inter-layer dropout, injected
            inp = inp * layer_masks[layer + 1].to(x.dtype)
    return h


def sequence_encoder_forward(sd: StateDict, x: torch.Tensor, num_layers: int, encoder_type: str = "lstm",
                             lengths: Optional[torch.Tensor] = None,
                             layer_masks: Optional[Mapping[int, torch.Tensor]] = None) -> torch.Tensor:
    """``SequenceEncoder.forward`` for the rnn variants in eval mode (encoders.py:115-166):
    last hidden state of the top layer -> (dropout = identity) -> ``projection``."""
    if x.dim() != 3:
        raise ValueError(f"Expected 3D input sequence, got shape {x.shape}")
    if encoder_type == "lstm":
        final = lstm_last_hidden(sd, "rnn", x, num_layers, lengths, layer_masks)
    elif encoder_type == "gru":
        final = gru_last_hidden(sd, "rnn", x, num_layers, lengths, layer_masks)
    else:
        raise ValueError(f"Unsupported encoder type: {encoder_type}")
    return final @ sd["projection.weight"].to(x.dtype).t() + sd["projection.bias"].to(x.dtype)


def mlp_encoder_forward(sd: StateDict, x: torch.Tensor, num_layers: int, batch_norm: bool = True,
                        training: bool = False, eps: float = 1e-5) -> torch.Tensor:
    """``SimpleMLPEncoder.forward`` (encoders.py:339-397) with dropout = identity.  BatchNorm1d uses batch
    statistics (biased variance) in training mode and the running statistics otherwise."""
    if x.dim() != 2:
        raise ValueError(f"Expected 2D feature tensor, got shape {x.shape}")
    idx = 0
    for _ in range(num_layers):
        x = x @ sd[f"encoder.{idx}.weight"].to(x.dtype).t() + sd[f"encoder.{idx}.bias"].to(x.dtype)
        idx += 1
        if batch_norm:
            if training:
                mean, var = x.mean(dim=0), x.var(dim=0, unbiased=False)
            else:
                mean, var = sd[f"encoder.{idx}.running_mean"].to(x.dtype), sd[f"encoder.{idx}.running_var"].to(x.dtype)
            x = (x - mean) / torch.sqrt(var + eps) * sd[f"encoder.{idx}.weight"].to(x.dtype) \
                + sd[f"encoder.{idx}.bias"].to(x.dtype)
            idx += 1
        x = torch.relu(x)
        idx += 2  # ReLU, Dropout
    return x @ sd[f"encoder.{idx}.weight"].to(x.dtype).t() + sd[f"encoder.{idx}.bias"].to(x.dtype)


def layer_norm(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """Per-modality ``nn.LayerNorm`` applied to the encoder output (src/train.py:170-171,267-268)."""
    mu = x.mean(dim=-1, keepdim=True)
    var = x.var(dim=-1, unbiased=False, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * weight + bias
