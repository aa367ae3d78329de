"""CPU: the host-side operand packing of the recurrence kernels (ops.lstm_pack_* / gru_as_four_gates: pure tensor
code, no kernel launch).  Each test replays ONE time step with the packed operands the way the kernels read them
(gate-interleaved rows 4u+g, 64-wide k-blocks, zero-padded input columns, the ones column of the bias gradient,
the four-gate form of the GRU) and compares with PyTorch's own cells — so a packing regression shows up without a GPU."""
import importlib

import pytest
import torch
import torch.nn as nn

from conftest import load_pkg

ops = importlib.import_module(load_pkg().__name__ + ".ops")


def _unpack_step(w_hh, w_ih, bias, x64, h):
    """pre[b, 4u+g] exactly as lstm_seq_kernel accumulates it: sum over the k-blocks of h plus the x k-block."""
    KBH, N4, _ = w_hh.shape
    pre = x64.float() @ w_ih.float().t() + bias
    for kb in range(KBH):
        pre = pre + h[:, 64 * kb:64 * kb + 64].float() @ w_hh[kb].float().t()
    return pre.view(h.shape[0], N4 // 4, 4)   # (B, unit, gate)


@pytest.mark.parametrize("feat,hidden", [(17, 64), (1, 128), (64, 256)])
def test_lstm_packing_reproduces_the_cell(feat, hidden):
    torch.manual_seed(61)
    cell = nn.LSTMCell(feat, hidden)
    x, h, c = torch.randn(6, feat), torch.randn(6, hidden) * 0.5, torch.randn(6, hidden)
    w_hh, w_ih, bias = ops.lstm_pack_weights(cell.weight_ih, cell.weight_hh, cell.bias_ih, cell.bias_hh)
    assert w_hh.shape == (hidden // 64, 4 * hidden, 64) and w_ih.shape == (4 * hidden, 64) and w_hh.dtype == torch.bfloat16
    xp = ops.lstm_pack_input(x.unsqueeze(1), ones_column=True)          # (T=1, B, 64)
    assert xp.shape == (1, 6, 64) and torch.equal(xp[0, :, :feat].float(), x.to(torch.bfloat16).float())
    if feat < 64:
        assert float(xp[0, 0, feat]) == 1.0 and float(w_ih[:, feat].abs().max()) == 0.0   # ones column, zero weight
    pre = _unpack_step(w_hh, w_ih, bias, xp[0], h.to(torch.bfloat16))
    i, f, g, o = (pre[:, :, k] for k in range(4))
    c_new = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
    h_new = torch.sigmoid(o) * torch.tanh(c_new)
    with torch.no_grad():
        h_ref, c_ref = cell(x, (h, c))
    assert float((h_new - h_ref).abs().max()) <= 2e-2 and float((c_new - c_ref).abs().max()) <= 3e-2   # bf16 operands
    # the transposed operand of the backward GEMM: out[n][4u+g] = weight_hh[g*H+u][n]
    wt = ops.lstm_pack_weights_t(cell.weight_hh)
    assert wt.shape == (hidden, 4 * hidden)
    u, gate, n = 5, 2, 9
    assert float(wt[n, 4 * u + gate]) == float(cell.weight_hh[gate * hidden + u, n].detach().to(torch.bfloat16))


@pytest.mark.parametrize("feat,hidden", [(17, 64), (3, 128)])
def test_gru_four_gate_form_reproduces_the_cell(feat, hidden):
    """gru_as_four_gates: gates (r, z, n_x, n_h) with the input's and the recurrent share of the candidate gate in
    separate columns; the kernel's cell h' = n + z (h - n), n = tanh(a_nx + r a_nh) on those columns is nn.GRUCell."""
    torch.manual_seed(62)
    cell = nn.GRUCell(feat, hidden)
    x, h = torch.randn(5, feat), torch.randn(5, hidden) * 0.5
    w_ih4, w_hh4, b_ih4, b_hh4 = ops.gru_as_four_gates(cell.weight_ih, cell.weight_hh, cell.bias_ih, cell.bias_hh)
    assert w_ih4.shape == (4 * hidden, feat) and w_hh4.shape == (4 * hidden, hidden)
    assert float(w_ih4[3 * hidden:].abs().max()) == 0.0 and float(w_hh4[2 * hidden:3 * hidden].abs().max()) == 0.0
    a = (x @ w_ih4.t() + b_ih4 + h @ w_hh4.t() + b_hh4).view(5, 4, hidden)   # fp32, gate-major
    r, z = torch.sigmoid(a[:, 0]), torch.sigmoid(a[:, 1])
    # the kernel adds b_in to a_nx and b_hn to a_nh through the summed bias: a[:, 2] = x W_in + b_in, a[:, 3] = h W_hn + b_hn
    n = torch.tanh(a[:, 2] + r * a[:, 3])
    h_new = n + z * (h - n)
    with torch.no_grad():
        ref = cell(x, h)
    assert float((h_new - ref).abs().max()) <= 1e-6
    # and through the LSTM packers (what the kernel actually loads): rows 4u+g, bf16
    w_hh, w_ih, bias = ops.lstm_pack_weights(w_ih4, w_hh4, b_ih4, b_hh4)
    pre = _unpack_step(w_hh, w_ih, bias, ops.lstm_pack_input(x.unsqueeze(1))[0], h.to(torch.bfloat16))
    r, z = torch.sigmoid(pre[:, :, 0]), torch.sigmoid(pre[:, :, 1])
    n = torch.tanh(pre[:, :, 2] + r * pre[:, :, 3])
    assert float((n + z * (h - n) - ref).abs().max()) <= 2e-2


def test_upper_layer_packing_and_lengths_validation():
    torch.manual_seed(63)
    rnn = nn.LSTM(64, 64, num_layers=2)
    w_hh, w_ih, bias = ops.lstm_pack_upper(rnn.weight_ih_l1, rnn.weight_hh_l1, rnn.bias_ih_l1, rnn.bias_hh_l1)
    assert w_hh.shape == (1, 256, 64) and w_ih.shape == (256, 64) and bias.shape == (256,)
    u, gate = 7, 3
    assert float(w_ih[4 * u + gate, 11]) == float(rnn.weight_ih_l1[gate * 64 + u, 11].detach().to(torch.bfloat16))
    assert abs(float(bias[4 * u + gate]) - float((rnn.bias_ih_l1[gate * 64 + u] + rnn.bias_hh_l1[gate * 64 + u]).detach())) <= 1e-6
    with pytest.raises(Exception, match="lengths must be"):
        ops._lstm_lengths(torch.tensor([0, 3]), 2, 5, torch.device("cpu"))
    with pytest.raises(Exception, match="lengths must be"):
        ops._lstm_lengths(torch.tensor([1, 6]), 2, 5, torch.device("cpu"))
    assert ops._lstm_lengths(torch.tensor([1, 5]), 2, 5, torch.device("cpu")).dtype == torch.int32
