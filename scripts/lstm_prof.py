"""Per-launch time of the LSTM step kernel (msf_prof events) for a few groupings."""
import importlib, os, sys, ctypes
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_pkg
pkg = load_pkg()
ops = importlib.import_module(pkg.__name__ + ".ops")
N = importlib.import_module(pkg.__name__ + "._native")
lib = pkg.lib()
H, T = 256, 64
for B, n in ((4096, 4), (4096, 1), (1024, 4), (512, 1)):
    torch.manual_seed(0)
    packed, xs = [], []
    for i in range(n):
        F = 17
        lstm = torch.nn.LSTM(F, H, batch_first=True).cuda()
        packed.append(ops.lstm_pack_weights(lstm.weight_ih_l0, lstm.weight_hh_l0, lstm.bias_ih_l0, lstm.bias_hh_l0))
        xs.append(ops.lstm_pack_input(torch.randn(B, T, F, device="cuda")))
    ops.lstm_forward(xs, packed, H)
    torch.cuda.synchronize()
    N.check(lib.msf_prof_enable(1))
    ops.lstm_forward(xs, packed, H)
    buf = ctypes.create_string_buffer(1 << 16)
    N.check(lib.msf_prof_report(buf, len(buf)))
    N.check(lib.msf_prof_enable(0))
    for line in buf.value.decode().splitlines():
        label, cnt, ms, fl = line.split("\t")
        us = float(ms) * 1e3 / int(cnt)
        print(f"B={B} n={n}: {label} {us:.1f} us/launch, {float(fl) / int(cnt) / us / 1e6:.0f} TFLOP/s, tiles {n * ((B + 127) // 128) * 4}")
