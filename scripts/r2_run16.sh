set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/pytest16.log 2>&1; echo "pytest rc=$?" >> $O/pytest16.log
tail -6 $O/pytest16.log | cut -c1-220
TL=$PWD/multimodal-sensor-fusion-with-attention-rajeevatla_b200/libmsf_b200_timeline.so
for B in 4096 2048; do
MSF_B200_LIB=$TL timeout 300 python scripts/step_timeline.py $B > $O/timeline16_$B.txt 2>&1; echo "timeline rc=$?"
grep -A10 "step 2" $O/timeline16_$B.txt | cut -c1-160
done
timeout 600 python bench.py --no-cpu-baseline --no-strong > $O/bench16.json 2> $O/bench16.err; echo "bench rc=$?"; tail -3 $O/bench16.err | cut -c1-300
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2/bench16*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d.get("roofline",{})
        print(f, d["ms_per_step"], d["value"], r.get("frac"), r.get("avg_launch_us"), d.get("e2e"))
        print([ (p["launch"][:14],p["us_per_launch"]) for p in r.get("per_launch",[])])
    except Exception as e: print(f, "ERR", e)
PY
