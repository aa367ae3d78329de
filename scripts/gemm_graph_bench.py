"""Steady-state per-launch time of msf_gemm_bf16 inside a CUDA graph (no Python / launch overhead)."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_pkg
ops = importlib.import_module(load_pkg().__name__ + ".ops")

def bench(m, n, k, mn=False, out=torch.bfloat16, reps=20, bias=True):
    if mn:
        a = torch.randn(k, m, device="cuda").bfloat16(); b = torch.randn(k, n, device="cuda").bfloat16()
    else:
        a = torch.randn(m, k, device="cuda").bfloat16(); b = torch.randn(n, k, device="cuda").bfloat16()
    bs = torch.randn(n, device="cuda") if bias else None
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        ops.gemm_bf16(a, b, mn_major=mn, bias=bs, out_dtype=out)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        for _ in range(reps):
            d = ops.gemm_bf16(a, b, mn_major=mn, bias=bs, out_dtype=out)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10):
        g.replay()
    e.record(); torch.cuda.synchronize()
    us = s.elapsed_time(e) * 1000 / (10 * reps)
    fl = 2.0 * m * n * k
    print("m=%5d n=%4d k=%5d mn=%d out=%s: %7.2f us/launch  %7.1f TFLOP/s" % (m, n, k, mn, str(out)[6:], us, fl / us / 1e6))

for shape in [(128, 32, 64), (128, 256, 64), (128, 256, 256), (4096, 32, 256), (4096, 256, 256), (4096, 256, 128),
              (18944, 256, 256), (37888, 256, 256), (4096, 256, 768)]:
    bench(*shape)
bench(4096, 32, 256, out=torch.float32)
bench(256, 256, 4096, mn=True, out=torch.float32, bias=False)
bench(256, 128, 4096, mn=True, out=torch.float32, bias=False)
