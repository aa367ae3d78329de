"""Probe of the cta_group::2 GEMM building block (msf_debug_pair_gemm) against torch."""
import importlib, os, sys, ctypes
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_pkg
pkg = load_pkg()
N = importlib.import_module(pkg.__name__ + "._native")
lib = pkg.lib()
torch.manual_seed(0)
for M, K in ((256, 64), (256, 256), (512, 512), (1000, 4096)):
    a = torch.randn(M, K, device="cuda").bfloat16()
    b = torch.randn(256, K, device="cuda").bfloat16()
    d = torch.full((M, 256), float("nan"), device="cuda")
    N.check(lib.msf_debug_pair_gemm(a.data_ptr(), b.data_ptr(), d.data_ptr(), M, K, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    ref = a.float() @ b.float().t()
    err = float((d - ref).abs().max())
    print(f"M={M} K={K}: max abs err {err:.3e} (ref max {float(ref.abs().max()):.1f}) nan={int(torch.isnan(d).sum())}", flush=True)
# timing
M, K = 16384, 4096
a = torch.randn(M, K, device="cuda").bfloat16(); b = torch.randn(256, K, device="cuda").bfloat16()
d = torch.empty(M, 256, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    N.check(lib.msf_debug_pair_gemm(a.data_ptr(), b.data_ptr(), d.data_ptr(), M, K, st))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    N.check(lib.msf_debug_pair_gemm(a.data_ptr(), b.data_ptr(), d.data_ptr(), M, K, st))
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / 20
print(f"pair gemm {M}x256x{K}: {us:.1f} us, {2.0 * M * 256 * K / us / 1e6:.0f} TFLOP/s on {M // 256 * 2} CTAs")

# tensor-pipe issue rate with resident operands
for ctas in (1, 148):
    for n in (64, 128, 256):
        c = ctypes.c_int64()
        N.check(lib.msf_debug_mma_rate(n, 2000, 4, ctas, ctypes.byref(c), torch.cuda.current_stream().cuda_stream))
        per = c.value / (2000 * 4)
        print(f"mma rate: ctas={ctas} M=128 N={n} K=16: {per:.1f} cycles/MMA -> {128 * n * 16 * 2 / per * 1.965e9 * 148 / 1e12:.0f} TFLOP/s chip-equivalent")
