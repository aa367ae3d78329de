cd $GRAFT_REPO_ROOT
O=gpurun_out/r2; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/pytest31.log 2>&1; echo "pytest rc=$?"
grep -v "^$" $O/pytest31.log | tail -30 | cut -c1-250
timeout 600 python bench.py --shape scaled --no-strong > $O/bench31_scaled.json 2> $O/bench31_scaled.err; echo "bench scaled rc=$?"; tail -5 $O/bench31_scaled.err | cut -c1-300
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2/bench31*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d["roofline"]
        print(f, d["ms_per_step"], d["value"], "e2e", d["e2e"]["ms_per_step"], d["e2e"].get("host_feature_dtype"), "frac", r.get("frac"), r.get("whole_step_frac"), d.get("cpu_baseline"))
        print([ (p["launch"][:18],p["us_per_launch"], p["launches"]) for p in r.get("per_launch",[])])
    except Exception as e: print(f, "ERR", e)
PY
