set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2; mkdir -p $O
nvidia-smi -L > $O/gpu.txt
timeout 900 python -m pytest tests -m gpu -q > $O/pytest13.log 2>&1; echo "pytest rc=$?" >> $O/pytest13.log
tail -12 $O/pytest13.log | cut -c1-220
timeout 600 python bench.py > $O/bench13.json 2> $O/bench13.err; echo "bench rc=$?"; tail -3 $O/bench13.err | cut -c1-300
TL=$PWD/multimodal-sensor-fusion-with-attention-rajeevatla_b200/libmsf_b200_timeline.so
MSF_B200_LIB=$TL timeout 300 python scripts/step_timeline.py > $O/timeline13.txt 2>&1; echo "timeline rc=$?"
MSF_CHAIN=v2 MSF_B200_LIB=$TL timeout 300 python scripts/chain_stamps.py > $O/stamps13.txt 2>&1; echo "stamps rc=$?"
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-strong > $O/plain13.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches13.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-strong > $O/ncu13.log 2>&1
echo "ncu rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2/bench13*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d.get("roofline",{})
        print(f, d["ms_per_step"], d["value"], r.get("frac"), r.get("avg_launch_us"), d.get("e2e"), d.get("run"))
        print([ (p["launch"][:14],p["us_per_launch"]) for p in r.get("per_launch",[])])
    except Exception as e: print(f, "ERR", e)
PY
grep -A12 "step 2" $O/timeline13.txt | cut -c1-160
tail -16 $O/stamps13.txt
