"""Modality encoders — drop-in for the reference's ``src/encoders.py``.

Same classes, constructor signatures, sub-module / parameter names (checkpoint
keys ``encoders.<m>.rnn.weight_ih_l0`` ... ``encoders.<m>.projection.*``), error
strings and factory routing (encoders.py:16-451).  What runs where:

* every ``nn.Linear`` of the encoders (the ``projection`` of ``SequenceEncoder``,
  the layers of ``SimpleMLPEncoder``, ``FrameEncoder``'s processor / projection)
  goes through the msf_b200 dense-layer kernels (``msf_linear_forward/backward``),
  with the ReLU fused where it follows directly;
* the recurrence of ``SequenceEncoder`` (``nn.LSTM`` / ``nn.GRU``, encoders.py:67-85,
  135-166) and the CNN / Transformer variants use PyTorch's own modules (library
  code) — the persistent LSTM kernel is the first "next" row of SURVEY.md §8f;
* sub-modules are looked up at call time, so the structural mutations the
  reference's tests perform (``encoder.rnn = None``, ``projection = nn.Identity()``,
  ``encoder_type = "bogus"``) behave as in the reference.

Device policy as in fusion.py: CUDA tensors run in place, CPU tensors are staged
through the current CUDA device for the kernel-backed layers, and without a CUDA
device those layers raise (there is no CPU fallback).
"""
from __future__ import annotations

import importlib
import os
import sys
from typing import Any, Dict, Optional, cast

import torch
import torch.nn as nn

_HERE = os.path.dirname(os.path.realpath(__file__))   # realpath: src/ may be reached through a symlink
_ROOT = os.path.dirname(os.path.dirname(_HERE))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
ops = importlib.import_module(os.path.basename(os.path.dirname(_HERE)) + ".ops")


def _dense(layer: nn.Module, x: torch.Tensor, relu: bool = False) -> torch.Tensor:
    """``layer(x)`` (optionally followed by ReLU).  ``nn.Linear`` runs on the msf_b200
    kernels; anything else (``nn.Identity``, a test double) is simply called."""
    if not isinstance(layer, nn.Linear):
        y = layer(x)
        return torch.relu(y) if relu else y
    home, dtype = x.device, x.dtype
    dev = home if home.type == "cuda" else ops.require_cuda("encoder projection")
    with torch.cuda.device(dev):
        y = ops.linear(x.to(device=dev, dtype=torch.float32),
                       layer.weight.to(device=dev, dtype=torch.float32),
                       None if layer.bias is None else layer.bias.to(device=dev, dtype=torch.float32), relu)
    return y.to(device=home, dtype=dtype)


def _lstm_tensor_core_ok(enc, sequence: torch.Tensor, lengths) -> bool:
    """The hand-written recurrence (msf_lstm_forward / msf_lstm_backward: bf16 operands, fp32 state, <= 1e-2) is used
    when the caller opted into bf16 (``encoder.precision = "bf16"`` or MSF_PRECISION=bf16): uni-directional LSTM of
    any depth, CUDA input, hidden % 64 == 0, input_dim <= 64; per-window ``lengths`` (the reference packs ragged
    windows, src/encoders.py:140-152), stacked layers and the training mode (gradients wanted for the LSTM's
    parameters) on the persistent kernels (hidden <= 256).  GRU encoders (src/encoders.py:66-72): the
    same kernels (hidden <= 256), inference and training.  A gradient with respect to the input sequence is not provided
    on this path (the encoders' inputs are data): such a call takes the fp32 kernels (_lstm_fp32)."""
    prec = getattr(enc, "precision", None) or os.environ.get("MSF_PRECISION", "fp32")
    rnn = enc.rnn
    if not (prec == "bf16" and enc.encoder_type in ("lstm", "gru") and sequence.is_cuda
            and isinstance(rnn, (nn.LSTM, nn.GRU)) and not rnn.bidirectional and getattr(rnn, "proj_size", 0) == 0
            and rnn.hidden_size % 64 == 0 and rnn.input_size <= 64):
        return False
    if torch.is_grad_enabled() and sequence.requires_grad:
        return False
    persistent = rnn.hidden_size <= 256 and not os.environ.get("MSF_LSTM_STEPS")
    if _lstm_wants_grad(rnn):
        return persistent
    if lengths is not None or rnn.num_layers > 1 or isinstance(rnn, nn.GRU):
        return persistent
    return True


def _lstm_wants_grad(rnn: nn.Module) -> bool:
    return torch.is_grad_enabled() and any(p.requires_grad for p in rnn.parameters())


def _lstm_layers(rnn: nn.Module):
    return [tuple(getattr(rnn, f"{name}_l{l}", None) for name in ("weight_ih", "weight_hh", "bias_ih", "bias_hh"))
            for l in range(rnn.num_layers)]


def _lstm_tensor_core(rnn: nn.Module, sequence: torch.Tensor, lengths: Optional[torch.Tensor] = None,
                      seed: Optional[int] = None) -> torch.Tensor:
    """``h_n[-1]`` of ``rnn`` on the hand-written kernels.  Training mode (gradients wanted): nn.LSTM's inter-layer
    dropout is drawn from the library's Philox stream keyed by ``seed`` (default: one draw from torch's generator)."""
    layers = _lstm_layers(rnn)
    with torch.cuda.device(sequence.device):
        if _lstm_wants_grad(rnn):   # training: the tapes of the forward pass feed msf_lstm_backward
            p = float(rnn.dropout) if (rnn.training and rnn.num_layers > 1) else 0.0
            if seed is None:
                seed = int(torch.randint(0, 2**62, (1,)).item()) if p > 0.0 else 0
            flat = [t for layer in layers for t in layer]
            return ops.LstmLastHidden.apply(sequence.detach(), lengths, p, int(seed), *flat,
                                            cell="gru" if isinstance(rnn, nn.GRU) else "lstm")
        if rnn.num_layers > 1 or isinstance(rnn, nn.GRU):
            return ops.lstm_forward_stack(sequence, layers, rnn.hidden_size, lengths,
                                          "gru" if isinstance(rnn, nn.GRU) else "lstm")
        packed = ops.lstm_pack_weights(*layers[0])
        return ops.lstm_forward([ops.lstm_pack_input(sequence.to(torch.float32))], [packed], rnn.hidden_size,
                                None if lengths is None else [lengths])[0]


def _lstm_fp32_ok(enc, sequence: torch.Tensor) -> bool:
    """The default (fp32) precision of an LSTM / GRU encoder on a CUDA input runs the recurrence on the library's own
    fp32 kernels (msf_lstm_f32_* / msf_gru_f32_*: FFMA GEMMs + cell kernels, <= 1e-5 of the reference, gradients for the
    parameters and the input), any depth, with or without ``lengths``.  MSF_LSTM_LIBRARY=1 keeps torch.nn.LSTM / GRU."""
    rnn = enc.rnn
    return (enc.encoder_type in ("lstm", "gru") and sequence.is_cuda and isinstance(rnn, (nn.LSTM, nn.GRU))
            and not rnn.bidirectional and getattr(rnn, "proj_size", 0) == 0 and not os.environ.get("MSF_LSTM_LIBRARY"))


def _lstm_fp32(rnn: nn.Module, sequence: torch.Tensor, lengths: Optional[torch.Tensor] = None,
               seed: Optional[int] = None) -> torch.Tensor:
    p = float(rnn.dropout) if (rnn.training and rnn.num_layers > 1) else 0.0
    if seed is None:
        seed = int(torch.randint(0, 2**62, (1,)).item()) if p > 0.0 else 0
    flat = [t for layer in _lstm_layers(rnn) for t in layer]
    with torch.cuda.device(sequence.device):
        return ops.LstmLastHiddenF32.apply(sequence, lengths, p, int(seed), "gru" if isinstance(rnn, nn.GRU) else "lstm", *flat)


def _rnn_fp32(rnn: nn.Module, inp):
    """Run the library recurrence in true fp32: cuDNN's RNN kernels default to TF32, which is
    ~1e-4 away from the reference's CPU arithmetic."""
    with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
        return rnn(inp)


def _bn_act(bn: nn.BatchNorm1d, y: torch.Tensor, relu: bool, drop: Optional[nn.Dropout]) -> torch.Tensor:
    """``Dropout(ReLU(BatchNorm1d(y)))`` as one fused forward (and one fused backward) on msf_bn_act_*
    (encoders.py:374-377 of the reference).  Statistics, running-average update and ``num_batches_tracked`` follow
    nn.BatchNorm1d; under data parallelism they are per process, as in the reference (no SyncBatchNorm)."""
    home, dtype = y.device, y.dtype
    dev = home if home.type == "cuda" else ops.require_cuda("SimpleMLPEncoder batch norm")
    put = lambda t: None if t is None else t.to(device=dev, dtype=torch.float32)  # noqa: E731
    training = bn.training
    use_batch = training or not bn.track_running_stats or bn.running_mean is None
    momentum = bn.momentum
    if training and bn.track_running_stats and bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(1)
        if momentum is None:   # cumulative moving average
            momentum = 1.0 / float(bn.num_batches_tracked)
    momentum = 0.0 if momentum is None else float(momentum)
    update = training and bn.track_running_stats and bn.running_mean is not None
    rm = rv = None
    if update or not use_batch:
        rm, rv = put(bn.running_mean), put(bn.running_var)
        if update and rm.data_ptr() != bn.running_mean.data_ptr():
            rm, rv = rm.clone(), rv.clone()
    p = float(drop.p) if (drop is not None and drop.training) else 0.0
    with torch.cuda.device(dev):
        out = ops.batch_norm_act(put(y), put(bn.weight), put(bn.bias), rm, rv, momentum, bn.eps, use_batch, relu, p)
    if update and rm.data_ptr() != bn.running_mean.data_ptr():   # staged copies: hand the moved-on statistics back
        with torch.no_grad():
            bn.running_mean.copy_(rm)
            bn.running_var.copy_(rv)
    return out.to(device=home, dtype=dtype)


def _run_sequential(seq: nn.Sequential, x: torch.Tensor) -> torch.Tensor:
    """An ``nn.Sequential`` with its Linear layers on the kernels; ``Linear -> ReLU`` pairs are fused, and
    ``BatchNorm1d [-> ReLU] [-> Dropout]`` behind a Linear runs as one fused kernel sequence (_bn_act)."""
    mods = list(seq)
    i = 0
    while i < len(mods):
        if isinstance(mods[i], nn.BatchNorm1d) and x.dim() == 2:
            j = i + 1
            relu = j < len(mods) and isinstance(mods[j], nn.ReLU)
            j += 1 if relu else 0
            drop = mods[j] if j < len(mods) and isinstance(mods[j], nn.Dropout) and not mods[j].inplace else None
            j += 1 if drop is not None else 0
            x = _bn_act(mods[i], x, relu, drop)
            i = j
            continue
        fuse = isinstance(mods[i], nn.Linear) and i + 1 < len(mods) and isinstance(mods[i + 1], nn.ReLU)
        x = _dense(mods[i], x, relu=fuse)
        i += 2 if fuse else 1
    return x


class SequenceEncoder(nn.Module):
    """Time-series encoder: LSTM / GRU / 1-D CNN / Transformer -> fixed-size embedding."""

    encoder_type: str
    hidden_dim: int
    output_dim: int

    def __init__(self, input_dim: int, hidden_dim: int = 256, output_dim: int = 128, num_layers: int = 2,
                 encoder_type: str = "lstm", dropout: float = 0.1):
        super().__init__()
        me = cast(Any, self)
        me.encoder_type, me.hidden_dim, me.output_dim = encoder_type, hidden_dim, output_dim
        self.dropout_layer = nn.Dropout(dropout)
        me.rnn = me.conv_net = me.pool = me.input_projection = me.transformer = None
        self.projection = nn.Identity()
        rnn_drop = dropout if num_layers > 1 else 0.0
        if encoder_type == "lstm":
            me.rnn = nn.LSTM(input_dim, hidden_dim, num_layers=num_layers, batch_first=True, dropout=rnn_drop)
        elif encoder_type == "gru":
            me.rnn = nn.GRU(input_dim, hidden_dim, num_layers=num_layers, batch_first=True, dropout=rnn_drop)
        elif encoder_type == "cnn":
            me.conv_net = nn.Sequential(
                nn.Conv1d(input_dim, hidden_dim, kernel_size=3, padding=1), nn.BatchNorm1d(hidden_dim), nn.ReLU(),
                nn.Conv1d(hidden_dim, hidden_dim, kernel_size=3, padding=1), nn.BatchNorm1d(hidden_dim), nn.ReLU())
            me.pool = nn.AdaptiveAvgPool1d(1)
        elif encoder_type == "transformer":
            me.input_projection = nn.Linear(input_dim, hidden_dim)
            layer = nn.TransformerEncoderLayer(d_model=hidden_dim, nhead=4 if hidden_dim % 4 == 0 else 1,
                                               dropout=dropout, batch_first=True)
            me.transformer = nn.TransformerEncoder(layer, num_layers=num_layers)
        else:
            raise ValueError(f"Unknown encoder type: {encoder_type}")
        self.projection = nn.Linear(hidden_dim, output_dim)

    def forward(self, sequence: torch.Tensor, lengths: Optional[torch.Tensor] = None) -> torch.Tensor:
        """``(batch, seq_len, input_dim)`` [+ ``lengths (batch,)``] -> ``(batch, output_dim)``."""
        if sequence.dim() != 3:
            raise ValueError(f"Expected 3D input sequence, got shape {sequence.shape}")
        batch, seq_len, _ = sequence.shape

        if self.encoder_type in ("lstm", "gru"):
            if self.rnn is None:
                raise RuntimeError("RNN module not initialized.")
            if _lstm_tensor_core_ok(self, sequence, lengths):
                last = _lstm_tensor_core(self.rnn, sequence, lengths,
                                         getattr(self, "lstm_dropout_seed", None)).to(sequence.dtype)
                return _dense(self.projection, self.dropout_layer(last))
            if _lstm_fp32_ok(self, sequence):
                last = _lstm_fp32(self.rnn, sequence, lengths, getattr(self, "lstm_dropout_seed", None)).to(sequence.dtype)
                return _dense(self.projection, self.dropout_layer(last))
            if lengths is not None:  # ragged windows: pack, like encoders.py:141-156
                lens = lengths.to(device=sequence.device).to(torch.int64).cpu()
                packed = nn.utils.rnn.pack_padded_sequence(sequence, lens, batch_first=True, enforce_sorted=False)
                _, hidden = _rnn_fp32(self.rnn, packed)
            else:
                _, hidden = _rnn_fp32(self.rnn, sequence)
            state = hidden[0] if self.encoder_type == "lstm" else hidden
            return _dense(self.projection, self.dropout_layer(state[-1]))

        if self.encoder_type == "cnn":
            if self.conv_net is None or self.pool is None:
                raise RuntimeError("CNN modules not initialized.")
            x = self.pool(self.conv_net(sequence.transpose(1, 2))).squeeze(-1)
            return _dense(self.projection, self.dropout_layer(x))

        if self.encoder_type == "transformer":
            if self.input_projection is None or self.transformer is None:
                raise RuntimeError("Transformer modules not initialized.")
            x = _dense(self.input_projection, sequence)
            pad = None
            if lengths is not None:
                lens = lengths.to(device=sequence.device, dtype=torch.long)
                pad = torch.arange(seq_len, device=sequence.device).unsqueeze(0).expand(batch, -1) >= lens.unsqueeze(1)
            out = self.transformer(x, src_key_padding_mask=pad)
            if pad is not None:
                valid = (~pad).unsqueeze(-1).float()
                pooled = (out * valid).sum(dim=1) / valid.sum(dim=1).clamp_min(1.0)
            else:
                pooled = out.mean(dim=1)
            return _dense(self.projection, self.dropout_layer(pooled))

        raise ValueError(f"Unsupported encoder type: {self.encoder_type}")

    # ---- several recurrent encoders in one go (pipeline.EncodeFuse) -------------------------------------------
    def group_key(self, sequence: torch.Tensor):
        """Encoders whose keys are equal (and not None) can share the launches of the tensor-core recurrence: same
        cell, depth, hidden size, batch, length, mode.  None: this encoder runs by itself."""
        if (self.encoder_type not in ("lstm", "gru") or self.rnn is None or sequence.dim() != 3
                or not _lstm_tensor_core_ok(self, sequence, None)):
            return None
        rnn = self.rnn
        return (type(rnn).__name__, rnn.num_layers, rnn.hidden_size, tuple(sequence.shape[:2]), str(sequence.device),
                _lstm_wants_grad(rnn), bool(rnn.training), float(rnn.dropout))

    @staticmethod
    def forward_group(encoders, sequences):
        """``[enc(seq) for enc, seq in zip(encoders, sequences)]`` for up to 4 encoders with equal ``group_key``: every
        layer of the recurrence (forward and, in training mode, backward) is ONE persistent launch for all of them
        (ops.lstm_forward_stack_group / ops.LstmLastHiddenGroup) instead of one per encoder."""
        n = len(encoders)
        rnn0 = encoders[0].rnn
        layers_list = [_lstm_layers(e.rnn) for e in encoders]
        with torch.cuda.device(sequences[0].device):
            if _lstm_wants_grad(rnn0):
                p = float(rnn0.dropout) if (rnn0.training and rnn0.num_layers > 1) else 0.0
                seeds = []
                for e in encoders:
                    seed = getattr(e, "lstm_dropout_seed", None)
                    seeds.append(int(seed) if seed is not None else
                                 (int(torch.randint(0, 2**62, (1,)).item()) if p > 0.0 else 0))
                flat = [t for layers in layers_list for layer in layers for t in layer]
                lasts = ops.LstmLastHiddenGroup.apply(n, rnn0.num_layers, None, p, seeds,
                                                      "gru" if isinstance(rnn0, nn.GRU) else "lstm",
                                                      *[s.detach() for s in sequences], *flat)
            else:
                lasts = ops.lstm_forward_stack_group(list(sequences), layers_list, rnn0.hidden_size, None,
                                                     "gru" if isinstance(rnn0, nn.GRU) else "lstm")
        return [_dense(e.projection, e.dropout_layer(h.to(s.dtype))) for e, h, s in zip(encoders, lasts, sequences)]


class FrameEncoder(nn.Module):
    """Frame-feature encoder with attention / average / max temporal pooling (encoders.py:211-336)."""

    temporal_pooling: str

    def __init__(self, frame_dim: int, hidden_dim: int = 256, output_dim: int = 128,
                 temporal_pooling: str = "attention", dropout: float = 0.1):
        super().__init__()
        me = cast(Any, self)
        me.temporal_pooling = temporal_pooling
        self.frame_processor = nn.Sequential(nn.Linear(frame_dim, hidden_dim), nn.ReLU(), nn.Dropout(dropout))
        me.attention = None
        if temporal_pooling == "attention":
            me.attention = nn.Linear(hidden_dim, 1)
        elif temporal_pooling not in ("average", "max"):
            raise ValueError(f"Unknown pooling: {temporal_pooling}")
        self.projection = nn.Sequential(nn.Linear(hidden_dim, hidden_dim), nn.ReLU(), nn.Dropout(dropout),
                                        nn.Linear(hidden_dim, output_dim))

    def forward(self, frames: torch.Tensor, mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        if frames.dim() != 3:
            raise ValueError(f"Expected 3D frame tensor, got shape {frames.shape}")
        processed = _run_sequential(self.frame_processor, frames)
        if self.temporal_pooling == "attention":
            pooled = self.attention_pool(processed, mask)
        elif self.temporal_pooling in ("average", "max"):
            pooled = _frame_pool(processed, mask, self.temporal_pooling)
        else:
            raise ValueError(f"Unknown pooling strategy: {self.temporal_pooling}")
        return _run_sequential(self.projection, pooled)

    def attention_pool(self, frames: torch.Tensor, mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Scores -> frame mask -> softmax over the frames -> NaN -> 0 -> weighted sum (encoders.py:312-336), one
        fused kernel each way (msf_frame_pool_*)."""
        if self.attention is None:
            raise RuntimeError("Attention layer not initialized.")
        if not isinstance(self.attention, nn.Linear):   # a test double: the reference's own op sequence
            scores = self.attention(frames)
            if mask is not None:
                scores = scores.masked_fill(mask.to(scores.device).unsqueeze(-1) == 0, float("-inf"))
            weights = torch.nan_to_num(torch.softmax(scores, dim=1), nan=0.0, posinf=0.0, neginf=0.0)
            return (weights * frames).sum(dim=1)
        return _frame_pool(frames, mask, "attention", self.attention.weight, self.attention.bias)


def _frame_pool(frames: torch.Tensor, mask: Optional[torch.Tensor], mode: str, weight=None, bias=None) -> torch.Tensor:
    home, dtype = frames.device, frames.dtype
    dev = home if home.type == "cuda" else ops.require_cuda("FrameEncoder pooling")
    with torch.cuda.device(dev):
        out = ops.frame_pool(frames.to(device=dev, dtype=torch.float32), None if mask is None else mask.to(dev), mode,
                             None if weight is None else weight.to(dev), None if bias is None else bias.to(dev))
    return out.to(device=home, dtype=dtype)


class SimpleMLPEncoder(nn.Module):
    """``[Linear -> BatchNorm1d -> ReLU -> Dropout] x num_layers -> Linear`` on pre-extracted features."""

    def __init__(self, input_dim: int, hidden_dim: int = 256, output_dim: int = 128, num_layers: int = 2,
                 dropout: float = 0.1, batch_norm: bool = True):
        super().__init__()
        layers, width = [], input_dim
        for _ in range(num_layers):
            layers.append(nn.Linear(width, hidden_dim))
            if batch_norm:
                layers.append(nn.BatchNorm1d(hidden_dim))
            layers += [nn.ReLU(), nn.Dropout(dropout)]
            width = hidden_dim
        layers.append(nn.Linear(width, output_dim))
        self.encoder = nn.Sequential(*layers)

    def forward(self, features: torch.Tensor) -> torch.Tensor:
        if features.dim() != 2:
            raise ValueError(f"Expected 2D feature tensor, got shape {features.shape}")
        return _run_sequential(self.encoder, features)


_SEQUENCE_NAMES = ("imu", "audio", "mocap", "accelerometer")


def build_encoder(modality: str, input_dim: int, output_dim: int,
                  encoder_config: Optional[Dict[str, Any]] = None) -> nn.Module:
    """Factory used by ``train.MultimodalFusionModule`` (encoders.py:400-451): an explicit ``type`` wins,
    then the modality name decides, unknown names get the MLP encoder."""
    config: Dict[str, Any] = dict(encoder_config) if encoder_config else {}
    kind = config.pop("type", None)
    name = modality.lower()
    if kind is None:
        if name in ("video", "frames"):
            kind = "frame"
        elif name in _SEQUENCE_NAMES or name.startswith("imu_"):
            kind = "sequence"
        else:
            kind = "mlp"
    if kind == "frame":
        return FrameEncoder(frame_dim=input_dim, output_dim=output_dim, **config)
    if kind == "sequence":
        return SequenceEncoder(input_dim=input_dim, output_dim=output_dim, **config)
    if kind == "mlp":
        return SimpleMLPEncoder(input_dim=input_dim, output_dim=output_dim, **config)
    # an unrecognised explicit type falls through to the name heuristics, like the reference
    return build_encoder(modality, input_dim, output_dim, config)


if __name__ == "__main__":
    print("Testing encoders...")
    demo_batch, demo_len, demo_in, demo_out = 4, 100, 64, 128

    print("\nTesting SequenceEncoder...")
    for kind_name in ("lstm", "gru", "cnn"):
        try:
            enc = SequenceEncoder(input_dim=demo_in, output_dim=demo_out, encoder_type=kind_name)
            got = enc(torch.randn(demo_batch, demo_len, demo_in))
            assert got.shape == (demo_batch, demo_out)
            print(f"✓ {kind_name} encoder working! Output shape: {got.shape}")
        except NotImplementedError:
            print(f"✗ {kind_name} encoder not implemented yet")
        except Exception as err:  # noqa: BLE001 - demo block reports and continues
            print(f"✗ {kind_name} encoder error: {err}")

    print("\nTesting FrameEncoder...")
    try:
        enc = FrameEncoder(frame_dim=512, output_dim=demo_out, temporal_pooling="attention")
        got = enc(torch.randn(demo_batch, 30, 512))
        assert got.shape == (demo_batch, demo_out)
        print(f"✓ FrameEncoder working! Output shape: {got.shape}")
    except NotImplementedError:
        print("✗ FrameEncoder not implemented yet")
    except Exception as err:  # noqa: BLE001
        print(f"✗ FrameEncoder error: {err}")

    print("\nTesting SimpleMLPEncoder...")
    try:
        enc = SimpleMLPEncoder(input_dim=demo_in, output_dim=demo_out)
        got = enc(torch.randn(demo_batch, demo_in))
        assert got.shape == (demo_batch, demo_out)
        print(f"✓ SimpleMLPEncoder working! Output shape: {got.shape}")
    except NotImplementedError:
        print("✗ SimpleMLPEncoder not implemented yet")
    except Exception as err:  # noqa: BLE001
        print(f"✗ SimpleMLPEncoder error: {err}")
