"""Timeline of ONE graph-replayed train step (BASELINE configs[1]) from the debug library built with
`MSF_BUILD_VARIANT=timeline python <pkg>/build.py`: per kernel, when its first CTA entered, when pdl_wait()
returned (first / last CTA) and when its CTAs ended (first / last), in microseconds from the step's first event.

    MSF_B200_LIB=<pkg>/libmsf_b200_timeline.so python scripts/step_timeline.py [B]
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
for _ in range(20):
    eng.train_step_resident()
torch.cuda.synchronize()
lib = pkg.lib()
lib.msf_debug_timeline.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int]
lib.msf_debug_timeline.restype = ctypes.c_int
buf = ctypes.create_string_buffer(1 << 18)
CH = "chain2_gemm.cu"
WG = "tc_gemm.cu#1" if os.environ.get("MSF_WG") == "v1" else "wg2_gemm.cu#0"
ORDER = ["proj_gemm.cu#0", CH + "#0", "head_gemm.cu#0", "fusion_bf16.cu#0", CH + "#1", "fusion_bf16.cu#1", WG,
         "opt_pack.cu#0"]
for rep in range(3):
    N.check(lib.msf_debug_timeline(buf, len(buf), 1))   # clear
    eng.train_step_resident()
    torch.cuda.synchronize()
    N.check(lib.msf_debug_timeline(buf, len(buf), 1))
    rows = {}
    ctas = {}
    for line in buf.value.decode().splitlines():
        name, *v = line.split("\t")
        if name.endswith(".cta"):
            ctas[name[:-4]] = [tuple(int(y) for y in x.split(":")) for x in v]
        else:
            rows[name] = [int(x) for x in v]
    t0 = min(v[0] for v in rows.values())
    print(f"--- step {rep}: kernel, CTAs | first entry | pdl_wait returned first..last | CTA end first..last (us)")
    prev_end = None
    for name in ORDER + sorted(set(rows) - set(ORDER)):
        if name not in rows:
            continue
        e, w0, w1, d0, d1, n, k6, k7 = rows[name]
        us = lambda t: (t - t0) / 1e3
        waited = f"{us(w0):7.1f}..{us(w1):7.1f}" if w1 else "      (no wait)  "
        gap = "" if prev_end is None or not w1 else f"  wait-return after predecessor's last end: {us(w0) - prev_end:+.1f}"
        print(f"{name:20s} {n:4d} | {us(e):7.1f} | {waited} | {us(d0):7.1f}..{us(d1):7.1f}{gap}")
        if k6 or k7:
            print(f"{'':20s}      marks: {us(k6):7.1f} {us(k7):7.1f}")
        if not name.startswith("fusion_bf16.cu"):
            prev_end = us(d1)
    if rep == 2 and os.environ.get("MSF_TL_CTAS"):   # per-CTA end times of the last step: block index (SM id) in order of completion
        for name in ORDER:
            if name in ctas:
                order = sorted(range(len(ctas[name])), key=lambda b: ctas[name][b][0])
                print(name, "ends:", " ".join(f"{b}({ctas[name][b][1]}):{(ctas[name][b][0] - t0) / 1e3:.1f}" for b in order))
