"""Time the LSTM recurrence alone (msf_lstm_forward) at the raw-window benchmark shape: 4 encoders, B windows, T steps,
hidden 256.  MSF_LSTM_STEPS=1: launch per step; MSF_LSTM_DBG: lstm_seq.cu debug switches (16 = per-step cycle stamps)."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_pkg
pkg = load_pkg()
ops = importlib.import_module(pkg.__name__ + ".ops")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
T = int(sys.argv[2]) if len(sys.argv) > 2 else 256
H = 256
torch.manual_seed(0)
feats = [17, 17, 17, 1]
rnns = [torch.nn.LSTM(f, H, batch_first=True).cuda() for f in feats]
packed = [ops.lstm_pack_weights(r.weight_ih_l0, r.weight_hh_l0, r.bias_ih_l0, r.bias_hh_l0) for r in rnns]
xs = [ops.lstm_pack_input(torch.randn(B, T, f, device="cuda")) for f in feats]
for _ in range(2):
    ops.lstm_forward(xs, packed, H)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(3):
    ops.lstm_forward(xs, packed, H)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 3
flop = B * sum(T * 2 * (f + H) * 4 * H for f in feats)
print(f"B={B} T={T}: {ms:.3f} ms per pass, {ms * 1e3 / T:.2f} us per step, {flop / ms / 1e9:.0f} TFLOP/s "
      f"(env: {os.environ.get('MSF_LSTM_STEPS', '')} dbg {os.environ.get('MSF_LSTM_DBG', '')})")
