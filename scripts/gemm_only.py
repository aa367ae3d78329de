"""Run the tcgen05 GEMM building block a few times (profiling target)."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_pkg
ops = importlib.import_module(load_pkg().__name__ + ".ops")
m, n, k = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (4096, 256, 256)))
a = torch.randn(m, k, device="cuda").bfloat16(); b = torch.randn(n, k, device="cuda").bfloat16()
bias = torch.randn(n, device="cuda")
for _ in range(3):
    d = ops.gemm_bf16(a, b, bias=bias, out_dtype=torch.bfloat16)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(20):
    d = ops.gemm_bf16(a, b, bias=bias, out_dtype=torch.bfloat16)
e.record(); torch.cuda.synchronize()
print("gemm %dx%dx%d: %.2f us/launch" % (m, n, k, s.elapsed_time(e) * 1000 / 20))
