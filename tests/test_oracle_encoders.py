"""CPU: the encoder restatements in oracle/encoder_oracle.py against golden vectors made by the
unmodified reference encoders (oracle/make_golden.py: encoder_cases), plus the drop-in's module
surface (names, routing, error strings) which needs no GPU."""
import sys

import pytest
import torch
import torch.nn as nn

from conftest import Golden, dropin_src
from oracle import encoder_oracle

if dropin_src() not in sys.path:
    sys.path.insert(0, dropin_src())
import encoders as dropin_encoders  # noqa: E402

TOL = 1e-5


def _maxabs(a, b):
    return float((a.double() - b.double()).abs().max())


@pytest.mark.parametrize("kind", ["lstm", "gru"])
def test_sequence_encoder_oracle_matches_reference(kind):
    g = Golden("encoders_small.npz")
    sd = g.group(f"{kind}/sd")
    x = g.t("seq/x")
    out = encoder_oracle.sequence_encoder_forward(sd, x, 2, kind)
    assert _maxabs(out, g.t(f"{kind}/out")) <= TOL
    # gradients through the restated recurrence (dropout 0): d(sum(out * w)) / dx and / d params
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    xg = x.clone().requires_grad_(True)
    out = encoder_oracle.sequence_encoder_forward(sdg, xg, 2, kind)
    (out * torch.linspace(-1, 1, out.shape[1]).unsqueeze(0)).sum().backward()
    assert _maxabs(xg.grad, g.t(f"{kind}/gradx")) <= TOL
    for key, ref in g.group(f"{kind}/grad").items():
        assert _maxabs(sdg[key].grad, ref) <= 5 * TOL, key


@pytest.mark.parametrize("kind", ["lstm", "gru"])
def test_lengths_stop_the_state(kind):
    """Ragged windows through the unmodified reference (SequenceEncoder.forward(x, lengths): packed sequence)."""
    g = Golden("encoders_small.npz")
    out = encoder_oracle.sequence_encoder_forward(g.group(f"{kind}/sd"), g.t("seq/x"), 2, kind, g.t("seq/lengths"))
    assert _maxabs(out, g.t(f"{kind}/out_lengths")) <= TOL


@pytest.mark.parametrize("kind", ["lstm", "gru"])
def test_ragged_windows_match_the_reference_call(kind):
    """The reference packs ragged windows (src/encoders.py:140-156: pack_padded_sequence -> nn.LSTM / nn.GRU -> h_n[-1]).
    The oracle's ``lengths`` handling (state stands still after a window's last valid step), which the GPU parity tests
    lean on for both cells, is held to exactly that call sequence of PyTorch's own CPU recurrence here."""
    torch.manual_seed(51)
    rnn = (nn.LSTM if kind == "lstm" else nn.GRU)(5, 12, num_layers=2, batch_first=True)
    x = torch.randn(9, 11, 5)
    lengths = torch.tensor([11, 1, 4, 7, 11, 2, 9, 3, 6])
    packed = nn.utils.rnn.pack_padded_sequence(x, lengths, batch_first=True, enforce_sorted=False)
    with torch.no_grad():
        _, hidden = rnn(packed)
    ref = (hidden[0] if kind == "lstm" else hidden)[-1]
    sd = {"rnn." + k: v.detach() for k, v in rnn.state_dict().items()}
    fn = encoder_oracle.lstm_last_hidden if kind == "lstm" else encoder_oracle.gru_last_hidden
    assert _maxabs(fn(sd, "rnn", x, 2, lengths), ref) <= TOL


@pytest.mark.parametrize("kind", ["lstm", "gru"])
def test_injected_inter_layer_dropout_is_applied_to_the_upper_layers_input(kind):
    """``layer_masks`` (the training-mode dropout nn.LSTM / nn.GRU applies between layers, injected so that the GPU
    kernels' Philox draws and the oracle agree): equal to running the layers one by one through PyTorch's own
    single-layer recurrences with the mask multiplied in between."""
    torch.manual_seed(52)
    cls = nn.LSTM if kind == "lstm" else nn.GRU
    rnn = cls(5, 12, num_layers=2, batch_first=True)
    x = torch.randn(7, 9, 5)
    mask = (torch.rand(7, 9, 12) > 0.3).float() / 0.7
    lower, upper = cls(5, 12, batch_first=True), cls(12, 12, batch_first=True)
    lower.load_state_dict({k[:-1] + "0": v for k, v in rnn.state_dict().items() if k.endswith("l0")})
    upper.load_state_dict({k[:-1] + "0": v for k, v in rnn.state_dict().items() if k.endswith("l1")})
    with torch.no_grad():
        out0, _ = lower(x)
        _, hidden = upper(out0 * mask)
    ref = (hidden[0] if kind == "lstm" else hidden)[-1]
    sd = {"rnn." + k: v.detach() for k, v in rnn.state_dict().items()}
    fn = encoder_oracle.lstm_last_hidden if kind == "lstm" else encoder_oracle.gru_last_hidden
    assert _maxabs(fn(sd, "rnn", x, 2, None, {1: mask}), ref) <= TOL
    assert _maxabs(fn(sd, "rnn", x, 2, None, {1: mask}), fn(sd, "rnn", x, 2)) > 1e-3   # the mask matters


def test_mlp_encoder_and_layernorm_oracle_match_reference():
    g = Golden("encoders_small.npz")
    sd = g.group("mlp/sd")
    assert _maxabs(encoder_oracle.mlp_encoder_forward(sd, g.t("mlp/x"), 2, True, training=False), g.t("mlp/out_eval")) <= TOL
    assert _maxabs(encoder_oracle.mlp_encoder_forward(sd, g.t("mlp/x"), 2, True, training=True), g.t("mlp/out_train")) <= TOL
    assert _maxabs(encoder_oracle.layer_norm(g.t("lstm/out"), g.t("ln/weight"), g.t("ln/bias")), g.t("ln/out")) <= TOL


def test_dropin_state_dict_keys_and_same_seed_init():
    g = Golden("encoders_small.npz")
    torch.manual_seed(41)  # seed used by make_golden for the lstm case: same registration order -> same weights
    enc = dropin_encoders.SequenceEncoder(17, hidden_dim=32, output_dim=16, num_layers=2, encoder_type="lstm", dropout=0.0)
    ref = g.group("lstm/sd")
    assert list(enc.state_dict().keys()) == list(ref.keys())
    for k, v in ref.items():
        assert torch.equal(enc.state_dict()[k], v), k
    torch.manual_seed(43)
    mlp = dropin_encoders.SimpleMLPEncoder(12, hidden_dim=24, output_dim=16, num_layers=2, dropout=0.0)
    for k, v in g.group("mlp/sd").items():
        assert torch.equal(mlp.state_dict()[k], v), k


def test_factory_routing_and_errors():
    E = dropin_encoders
    assert isinstance(E.build_encoder("imu_hand", 17, 128), E.SequenceEncoder)
    assert isinstance(E.build_encoder("heart_rate", 1, 128), E.SimpleMLPEncoder)
    assert isinstance(E.build_encoder("video", 512, 128), E.FrameEncoder)
    assert isinstance(E.build_encoder("heart_rate", 1, 128, {"type": "sequence", "encoder_type": "gru"}), E.SequenceEncoder)
    assert isinstance(E.build_encoder("audio", 8, 16, {"type": "bogus"}), E.SequenceEncoder)  # falls back to the name
    with pytest.raises(ValueError, match="Unknown encoder type"):
        E.SequenceEncoder(4, encoder_type="nope")
    with pytest.raises(ValueError, match="Unknown pooling"):
        E.FrameEncoder(4, temporal_pooling="nope")
    enc = E.SequenceEncoder(4, hidden_dim=8, output_dim=4, num_layers=1)
    with pytest.raises(ValueError, match="Expected 3D input sequence"):
        enc(torch.zeros(2, 4))
    enc.rnn = None
    with pytest.raises(RuntimeError, match="RNN module not initialized."):
        enc(torch.zeros(2, 3, 4))
    enc.encoder_type = "bogus"
    with pytest.raises(ValueError, match="Unsupported encoder type"):
        enc(torch.zeros(2, 3, 4))
    with pytest.raises(ValueError, match="Expected 2D feature tensor"):
        E.SimpleMLPEncoder(4)(torch.zeros(2, 3, 4))
    # no CUDA device here: a kernel-backed layer must fail loudly instead of falling back
    if not torch.cuda.is_available():
        with pytest.raises(Exception, match="no CPU or PyTorch-eager fallback"):
            E.SimpleMLPEncoder(4, hidden_dim=8, output_dim=4)(torch.zeros(2, 4))
    # a swapped-in projection is simply called (tests/test_encoders.py hot-swapping)
    enc2 = E.SequenceEncoder(4, hidden_dim=8, output_dim=4, num_layers=1)
    enc2.projection = nn.Identity()
    assert enc2(torch.zeros(2, 3, 4)).shape == (2, 8)
