"""ncu target for the recurrence kernels: one inference pass (lstm_seq_kernel<false>), one training-mode forward
(lstm_seq_kernel<true>) and one backward (lstm_bwd_kernel + weight-gradient GEMMs) of the 4 PAMAP2 encoders at
B = 4096 and a short sequence (the per-step behaviour does not depend on T).

    ncu --set full --clock-control none --import-source on -k regex:"lstm_seq_kernel|lstm_bwd_kernel" -c 3 \
        -o gpurun_out/lstm_kernels python scripts/lstm_prof_target.py [T]
"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "multimodal-sensor-fusion-with-attention-rajeevatla_b200"
ops = importlib.import_module(PKG + ".ops")

T = int(sys.argv[1]) if len(sys.argv) > 1 else 48
B, H = 4096, 256
dev = torch.device("cuda", 0)
torch.manual_seed(0)
feats = [17, 17, 17, 1]
rnns = [torch.nn.LSTM(f, H, batch_first=True).to(dev) for f in feats]
xs = [torch.randn(B, T, f, device=dev) for f in feats]
d_h = [torch.randn(B, H, device=dev) / B for _ in feats]
weights = [(r.weight_ih_l0, r.weight_hh_l0, r.bias_ih_l0, r.bias_hh_l0) for r in rnns]
with torch.no_grad():
    ops.lstm_forward([ops.lstm_pack_input(x) for x in xs], [ops.lstm_pack_weights(*w) for w in weights], H)
    ops.lstm_backward(ops.lstm_train_forward(xs, weights, H), d_h)
torch.cuda.synchronize()
print("done")
