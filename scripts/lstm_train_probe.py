"""Timing probe of the LSTM training pass (ops.lstm_train_forward + ops.lstm_backward) for the 4 PAMAP2 encoders:
per-launch CUDA-event times (msf_prof_*) at B = 4096.  MSF_LSTM_DBG=16 prints the cycle stamps of CTA 0.

    python scripts/lstm_train_probe.py [T] [B]
"""
import ctypes
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "multimodal-sensor-fusion-with-attention-rajeevatla_b200"
ops = importlib.import_module(PKG + ".ops")
nat = importlib.import_module(PKG + "._native")

T = int(sys.argv[1]) if len(sys.argv) > 1 else 256
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
H = 256
dev = torch.device("cuda", 0)
torch.manual_seed(0)
feats = [17, 17, 17, 1]
rnns = [torch.nn.LSTM(f, H, batch_first=True).to(dev) for f in feats]
xs = [torch.randn(B, T, f, device=dev) for f in feats]
d_h = [torch.randn(B, H, device=dev) / B for _ in feats]
weights = [(r.weight_ih_l0, r.weight_hh_l0, r.bias_ih_l0, r.bias_hh_l0) for r in rnns]


def one_pass():
    return ops.lstm_backward(ops.lstm_train_forward(xs, weights, H), d_h)


one_pass()
torch.cuda.synchronize()
nat.check(nat.lib().msf_prof_enable(1))
one_pass()
buf = ctypes.create_string_buffer(1 << 16)
nat.check(nat.lib().msf_prof_report(buf, len(buf)))
nat.check(nat.lib().msf_prof_enable(0))
for row in buf.value.decode().splitlines():
    label, cnt, ms, fl = row.split("\t")
    print(f"B={B} T={T}: {label}: {cnt} launches, {float(ms):.3f} ms, {float(ms) * 1e3 / T:.2f} us per time step, "
          f"{float(fl) / (float(ms) * 1e-3) / 1e12:.0f} TFLOP/s (MSF_LSTM_DBG={os.environ.get('MSF_LSTM_DBG', '')})")

# one stacked encoder (the reference's default num_layers = 2): inference and training pass through the layer-wise path
LAYERS = int(sys.argv[3]) if len(sys.argv) > 3 else 2
if LAYERS > 1:
    rnn = torch.nn.LSTM(17, H, num_layers=LAYERS, batch_first=True, dropout=0.1).to(dev)
    layers = [tuple(getattr(rnn, f"{n}_l{l}") for n in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")) for l in range(LAYERS)]
    x = xs[0]

    def timed(fn, reps=3):
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    ms_inf = timed(lambda: ops.lstm_forward_stack(x, layers, H))
    ms_trn = timed(lambda: ops.lstm_backward_stack(ops.lstm_train_forward_stack(x, layers, H, None, 0.1, 7), d_h[0]))
    print(f"B={B} T={T}: one {LAYERS}-layer encoder: inference {ms_inf:.2f} ms, training pass (forward + backward) {ms_trn:.2f} ms")
