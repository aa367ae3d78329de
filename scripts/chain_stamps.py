"""In-kernel wait accounting of the chained pair-GEMM kernel (chain3_kernel, CTA 0 of the last launch = the
backward chain of a train step) from the timeline build:

    MSF_B200_LIB=<pkg>/libmsf_b200_timeline.so python scripts/chain_stamps.py [B]
"""
import ctypes, importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_pkg
from helpers import PAMAP2, seeded_case
pkg = load_pkg()
engine = importlib.import_module(pkg.__name__ + ".engine")
N = importlib.import_module(pkg.__name__ + "._native")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
model, feats, mask, labels = seeded_case(PAMAP2, 256, 4, 25, B, seed=7, device="cuda")
eng = engine.FusionEngine(model, B, precision="bf16", seed=5, use_graph=True)
eng.load_batch(feats, mask, labels)
for _ in range(10):
    eng.train_step_resident()
torch.cuda.synchronize()
lib = pkg.lib()
out = (ctypes.c_int64 * 32)()
N.check(lib.msf_debug_chain_stamps(out))
v = list(out)
clk = 1.965e3  # cycles per us at the maximum SM clock
t0 = v[0]
us = lambda c: c / clk
print("chain3 stamps of CTA 0 (us; absolute stamps relative to kernel start after the cluster sync)")
print(f"  first operands landed (MMA thread)   {us(v[12] - t0):7.2f}")
print(f"  MMA loop end                         {us(v[4] - t0):7.2f}")
print(f"  MMA waits: operands {us(v[1]):6.2f}  T drained {us(v[2]):6.2f}  staged blocks {us(v[3]):6.2f}"
      f"  -> issuing {us(v[4] - v[12] - v[1] - v[2] - v[3]):6.2f}")
print(f"  epilogue first block start           {us(v[7] - t0):7.2f}")
print(f"  epilogue end                         {us(v[9] - t0):7.2f}")
print(f"  epilogue waits: G1 {us(v[5]):6.2f}  free staging block {us(v[6]):6.2f}  ACC {us(v[8]):6.2f}"
      f"  -> working {us(v[9] - v[7] - v[5] - v[6] - v[8]):6.2f}")
print(f"  W producer waits for free slots {us(v[10]):6.2f}, ends at {us(v[11] - t0):7.2f}")
rel = lambda i: us(v[i] - t0) if v[i] else float("nan")
print(f"  epilogue thread 128, first item: aux blocks in ACC / ReLU bits ready at {rel(16):7.2f}")
for i in range(3):
    print(f"    pass {i}: T ready {rel(17 + 2 * i):7.2f}  staged {rel(18 + 2 * i):7.2f}  ({us(v[18 + 2 * i] - v[17 + 2 * i]):5.2f} us)")
print(f"    final pass: ACC ready {rel(23):7.2f}  end {rel(9):7.2f}  ({us(v[9] - v[23]):5.2f} us)")
