set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2/pytest11.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest11.log
tail -12 gpurun_out/r2/pytest11.log | cut -c1-220
timeout 600 python bench.py > gpurun_out/r2/bench11.json 2> gpurun_out/r2/bench11.err; echo "bench rc=$?"; tail -3 gpurun_out/r2/bench11.err | cut -c1-300
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2/*11*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["value"], d.get("roofline",{}).get("frac"), d.get("roofline",{}).get("avg_launch_us"), d.get("strong_32768"), d.get("run"), d.get("e2e"))
    except Exception as e: print(f, "ERR", e)
PY
