import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_pkg
ops = importlib.import_module(load_pkg().__name__ + ".ops")
m, n, k = 37888, 256, 256
a = torch.randn(m, k, device="cuda").bfloat16(); b = torch.randn(n, k, device="cuda").bfloat16()
bias = torch.randn(n, device="cuda")
for _ in range(5):
    d = ops.gemm_bf16(a, b, bias=bias, out_dtype=torch.bfloat16)
torch.cuda.synchronize()
print("ok")
