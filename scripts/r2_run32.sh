cd $GRAFT_REPO_ROOT
O=gpurun_out/r2; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_fullwidth.py -m gpu -q > $O/pytest32.log 2>&1; echo "pytest rc=$?"
grep -v "^$" $O/pytest32.log | tail -12 | cut -c1-250
for v in "" "MSF_WG=v1"; do
env $v timeout 600 python bench.py --shape scaled --no-strong --no-cpu-baseline > $O/bench32_scaled_$v.json 2> $O/bench32_scaled_$v.err; echo "bench scaled rc=$?"; tail -2 $O/bench32_scaled_$v.err | cut -c1-300
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2/bench32*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d["roofline"]
        print(f, d["ms_per_step"], d["value"], "frac", r.get("whole_step_frac"))
        print([ (p["launch"][:18],p["us_per_launch"], p["launches"]) for p in r.get("per_launch",[]) if "WG" in p["launch"]])
    except Exception as e: print(f, "ERR", e)
PY
