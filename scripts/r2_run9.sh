set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2/pytest10.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest9.log
tail -12 gpurun_out/r2/pytest10.log | cut -c1-200
timeout 600 python bench.py > gpurun_out/r2/bench10.json 2> gpurun_out/r2/bench10.err; echo "bench rc=$?"; tail -3 gpurun_out/r2/bench10.err
for v in v2 v3; do
  MSF_CHAIN=$v timeout 300 python bench.py --workload infer_sweep --steps 10 > gpurun_out/r2/sweep10_$v.json 2> gpurun_out/r2/sweep10_$v.err; echo "sweep $v rc=$?"
  MSF_CHAIN=$v timeout 300 python bench.py --workload infer_sweep --steps 10 --no-mask-hint > gpurun_out/r2/sweep10d_$v.json 2> gpurun_out/r2/sweep10d_$v.err; echo "sweep dense $v rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2/*10*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["value"], d.get("roofline",{}).get("frac"), d.get("roofline",{}).get("avg_launch_us"), d.get("strong_32768"), d.get("run",{}).get("repeats"))
    except Exception as e: print(f, "ERR", e)
PY
