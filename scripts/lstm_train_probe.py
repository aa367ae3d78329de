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
