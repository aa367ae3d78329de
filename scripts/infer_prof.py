"""Per-launch times of one inference pass at B = 65536 (msf_prof events, eager)."""
import importlib, os, sys, ctypes
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_pkg
from helpers import PAMAP2, seeded_case
pkg = load_pkg()
ops = importlib.import_module(pkg.__name__ + ".ops")
N = importlib.import_module(pkg.__name__ + "._native")
lib = pkg.lib()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
model, feats, mask, labels = seeded_case(PAMAP2, 256, 4, 25, B, seed=7, device="cuda")
plan = model._plan()
own = dict(model.named_parameters())
arena = plan.gather([own[k].detach() for k, _, _ in plan.slots])
a16 = plan.pack_bf16(arena)
xs = [feats[m].contiguous() for m in plan.names]
ws = torch.empty(plan.workspace_bytes(B, N.MSF_PREC_BF16), dtype=torch.uint8, device="cuda")
kw = dict(precision=N.MSF_PREC_BF16, arena_bf16=a16, workspace=ws)
full = torch.ones_like(mask)
for hint, m in ((0, full), (0b0101, None)):
    if m is None:
        m = torch.zeros_like(mask); m[:, [0, 2]] = 1.0
    for _ in range(3):
        ops.fusion_infer_pass_raw(plan, arena, xs, m, present_hint=hint, **kw)
    torch.cuda.synchronize()
    N.check(lib.msf_prof_enable(1))
    for _ in range(5):
        ops.fusion_infer_pass_raw(plan, arena, xs, m, present_hint=hint, **kw)
    buf = ctypes.create_string_buffer(1 << 16)
    N.check(lib.msf_prof_report(buf, len(buf)))
    N.check(lib.msf_prof_enable(0))
    print("hint", bin(hint))
    for line in buf.value.decode().splitlines():
        label, n, ms, fl = line.split("\t")
        us = float(ms) * 1e3 / int(n)
        print(f"  {label:45s} {us:8.1f} us  {float(fl) / int(n) / us / 1e6:7.0f} TFLOP/s")
