"""Target for profiling the fused head kernel alone: a few train / infer passes at B = 4096 (config 2)."""
import importlib, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_pkg
from helpers import PAMAP2, seeded_case
pkg = load_pkg()
ops = importlib.import_module(pkg.__name__ + ".ops")
N = importlib.import_module(pkg.__name__ + "._native")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
model, feats, mask, labels = seeded_case(PAMAP2, 256, 4, 25, B, seed=7, device="cuda")
plan = model._plan()
own = dict(model.named_parameters())
arena = plan.gather([own[k].detach() for k, _, _ in plan.slots])
a16 = plan.pack_bf16(arena)
xs = [feats[m].contiguous() for m in plan.names]
ws = torch.empty(plan.workspace_bytes(B, N.MSF_PREC_BF16), dtype=torch.uint8, device="cuda")
kw = dict(precision=N.MSF_PREC_BF16, arena_bf16=a16, workspace=ws)
for _ in range(3):
    ops.fusion_train_pass_raw(plan, arena, xs, mask, labels, smoothing=0.05, training=True, p=0.1, seed=1, **kw)
    ops.fusion_infer_pass_raw(plan, arena, xs, mask, **kw)
torch.cuda.synchronize()
lib = pkg.lib()
import ctypes
N.check(lib.msf_prof_enable(1))
for _ in range(10):
    ops.fusion_train_pass_raw(plan, arena, xs, mask, labels, smoothing=0.05, training=True, p=0.1, seed=1, **kw)
    ops.fusion_infer_pass_raw(plan, arena, xs, mask, **kw)
buf = ctypes.create_string_buffer(1 << 16)
N.check(lib.msf_prof_report(buf, len(buf)))
N.check(lib.msf_prof_enable(0))
for line in buf.value.decode().splitlines():
    label, n, ms, fl = line.split("\t")
    print(f"{label:45s} {float(ms) * 1e3 / int(n):8.1f} us")

names = ["P0 start", "P0 end", "E1 acq", "E1 end", "E2 acq", "E2 end", "E3 acq", "E3 end", "E4 acq", "E4 end", "P5 end"]
for what in ("train", "infer"):
    if what == "train":
        ops.fusion_train_pass_raw(plan, arena, xs, mask, labels, smoothing=0.05, training=True, p=0.1, seed=1, **kw)
    else:
        ops.fusion_infer_pass_raw(plan, arena, xs, mask, **kw)
    st = (ctypes.c_int64 * 16)()
    N.check(lib.msf_debug_head_stamps(st))
    n = 11 if what == "train" else 6
    print(what, " ".join(f"{names[i]}:{(st[i] - st[0]) / 1965.0:.1f}us" for i in range(n)),
          f"| entry:{(st[11] - st[0]) / 1965.0:.1f}us all-done:{(st[12] - st[0]) / 1965.0:.1f}us ticket:{(st[13] - st[0]) / 1965.0:.1f}us")

os.environ["MSF_NO_HEAD"] = "1"
ops.fusion_forward_raw(plan, arena, xs, mask, precision=N.MSF_PREC_BF16, training=True, p=0.1, seed=1, arena_bf16=a16, workspace=ws)
os.environ.pop("MSF_NO_HEAD")
st = (ctypes.c_int64 * 16)()
N.check(lib.msf_debug_chain_stamps(st))
us = lambda x: x / 1965.0
print("chain fwd CTA0: MMA stage-wait %.1f  T-drained-wait %.1f  u-staged-wait %.1f  MMA end @%.1f | epi: G1-wait %.1f u-free-wait %.1f compute %.1f final-wait %.1f end @%.1f | producer: free-stage-wait %.1f end @%.1f"
      % (us(st[1]), us(st[2]), us(st[3]), us(st[4] - st[0]), us(st[5]), us(st[6]), us(st[7]), us(st[8]), us(st[9] - st[0]), us(st[10]), us(st[11] - st[0])))

st = (ctypes.c_int64 * 16)()
N.check(lib.msf_debug_proj_stamps(st))
lab = ["entry", "setup done", "consts", "x landed", "X phase end", "philox end", "gemm done", "epilogue end", "all done"]
print("proj CTA0:", " ".join(f"{lab[i]}:{us(st[i] - st[0]):.1f}" for i in range(9)))
