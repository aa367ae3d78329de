"""A/B the train-step bench over several builds of the library on ONE box (box-to-box spread is larger than most
kernel changes): runs `bench.py --no-cpu-baseline` alternately with MSF_B200_LIB pointing at each build and prints
one line per run.  Typical use inside a single gpurun call:

    cp <pkg>/libmsf_b200.so ab_tmp/lib_A.so        # build A, then edit + rebuild -> lib_B.so ...
    python scripts/ab_bench.py --rounds 2 ab_tmp/lib_A.so ab_tmp/lib_B.so

(`ab_tmp/` is git-ignored but travels to the GPU box.)  This is how the scheduling changes of round 1 were chosen
(profiles/README.md, "Block-scheduler findings")."""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ap = argparse.ArgumentParser()
ap.add_argument("libs", nargs="+")
ap.add_argument("--rounds", type=int, default=2)
ap.add_argument("--steps", type=int, default=304)
ap.add_argument("--warmup", type=int, default=32)
ap.add_argument("--extra", default="", help="further bench.py arguments, e.g. '--steps-per-graph 1'")
args = ap.parse_args()
for r in range(args.rounds):
    for lib in args.libs:
        env = dict(os.environ, MSF_B200_LIB=os.path.abspath(lib))
        cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--steps", str(args.steps), "--warmup", str(args.warmup),
               "--no-cpu-baseline", *args.extra.split()]
        res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
        lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
        if res.returncode != 0 or not lines:
            print(f"{lib}: FAILED rc={res.returncode} {res.stderr[-300:]!r}", flush=True)
            continue
        d = json.loads(lines[-1])
        print(f"{os.path.basename(lib):24s} round {r}: {d['ms_per_step'] * 1e3:7.1f} us/step   e2e "
              f"{d['e2e']['ms_per_step'] * 1e3:7.1f} us   launches/step {d['gpu_launches_per_step']}   "
              f"final loss {d['final_loss']:.4f}", flush=True)
