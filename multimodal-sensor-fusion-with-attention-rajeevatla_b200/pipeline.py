"""Encoders -> per-modality LayerNorm -> fusion model: the data path of
``train.MultimodalFusionModule.forward`` (src/train.py:233-291) without Lightning.

The reference's trainer cannot be imported here (pytorch_lightning / hydra are absent) and is out of scope; what is
on the hot path is its forward wiring, and that is what this module provides to callers of the drop-in modules:

    model = EncodeFuse(encoders, fusion_model, layer_norms)        # same three ModuleDict / Module attributes
    logits = model(features, mask)                                  # train.py:233-291
    logits, aux = model(features, mask, return_attention=True)     # HybridFusion only, like train.py:247-250

Semantics kept from the reference: modalities are encoded in ``encoders`` order, a modality whose key is missing from
``features`` is skipped (train.py:264-265), LayerNorm is applied only where ``layer_norms`` has the key
(train.py:267-268), a tuple returned by the fusion model is split into ``(logits, aux)`` (train.py:281-286).

What is different: with a ``HybridFusion`` the LayerNorm is not a separate pass over the encoder outputs — the raw
encoder outputs and the LayerNorm modules are handed to ``HybridFusion.forward(input_norms=...)``, which normalises
each row inside the projection kernel on the tensor-core path (``proj_gemm.cu``) and through ``msf_layer_norm_*``
otherwise; gradients reach the encoders and the LayerNorm parameters either way.
"""
from __future__ import annotations

from typing import Mapping, Optional

import torch
import torch.nn as nn


class EncodeFuse(nn.Module):
    def __init__(self, encoders: Mapping[str, nn.Module], fusion_model: nn.Module,
                 layer_norms: Optional[Mapping[str, nn.Module]] = None):
        super().__init__()
        self.encoders = encoders if isinstance(encoders, nn.ModuleDict) else nn.ModuleDict(dict(encoders))
        self.layer_norms = (layer_norms if isinstance(layer_norms, nn.ModuleDict)
                            else nn.ModuleDict(dict(layer_norms or {})))
        self.use_layer_norm = len(self.layer_norms) > 0
        self.fusion_model = fusion_model

    def _fuses_norms(self) -> bool:
        return type(self.fusion_model).__name__ == "HybridFusion" and hasattr(self.fusion_model, "_plan")

    def _encode(self, features: Mapping[str, torch.Tensor]):
        """``{m: encoders[m](features[m])}`` in ``encoders`` order (train.py:261-269).  Recurrent encoders on the
        tensor-core path that agree in cell, depth, hidden size, batch and length (``SequenceEncoder.group_key``) are
        run together, up to 4 per launch (``SequenceEncoder.forward_group``): the three IMU encoders and the heart-rate
        encoder of the PAMAP2 configuration share every launch of the recurrence."""
        todo = [(m, enc) for m, enc in self.encoders.items() if m in features]
        out, groups = {}, {}
        for m, enc in todo:
            key_fn = getattr(enc, "group_key", None)
            key = key_fn(features[m]) if callable(key_fn) else None
            if key is not None:
                groups.setdefault((type(enc), key), []).append(m)
        for (cls, _), names in groups.items():
            for i in range(0, len(names), 4):
                chunk = names[i:i + 4]
                if len(chunk) < 2:
                    continue
                res = cls.forward_group([self.encoders[m] for m in chunk], [features[m] for m in chunk])
                out.update(dict(zip(chunk, res)))
        return {m: out[m] if m in out else enc(features[m]) for m, enc in todo}

    def forward(self, features: Mapping[str, torch.Tensor], mask: Optional[torch.Tensor] = None,
                return_attention: bool = False):
        if return_attention and type(self.fusion_model).__name__ != "HybridFusion":
            raise ValueError("Attention information is only available for HybridFusion.")
        encoded = self._encode(features)
        norms = {m: self.layer_norms[m] for m in encoded if self.use_layer_norm and m in self.layer_norms}
        kwargs = {}
        if norms and self._fuses_norms():
            kwargs["input_norms"] = norms          # normalised inside the fusion model's projection kernel
        else:
            encoded = {m: (norms[m](x) if m in norms else x) for m, x in encoded.items()}
        if return_attention:
            kwargs["return_attention"] = True
        out = self.fusion_model(encoded, mask, **kwargs)
        logits, aux = (out[0], out[1] if len(out) > 1 else None) if isinstance(out, tuple) else (out, None)
        return (logits, aux) if return_attention else logits
