cd $GRAFT_REPO_ROOT
O=gpurun_out/r2; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest28.log 2>&1; echo "pytest rc=$?"
tail -3 $O/pytest28.log | cut -c1-250
TL=$PWD/multimodal-sensor-fusion-with-attention-rajeevatla_b200/libmsf_b200_timeline.so
MSF_B200_LIB=$TL timeout 300 python scripts/step_timeline.py 4096 > $O/timeline28.txt 2>&1; echo "timeline rc=$?"
grep -A10 "step 2" $O/timeline28.txt | cut -c1-160
timeout 600 python bench.py --no-cpu-baseline --no-strong > $O/bench28.json 2> $O/bench28.err; echo "bench rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2/bench28*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["value"], "e2e", d["e2e"]["ms_per_step"])
    except Exception as e: print(f, "ERR", e)
PY
