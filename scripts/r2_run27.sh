cd $GRAFT_REPO_ROOT
O=gpurun_out/r2; mkdir -p $O
TL=$PWD/multimodal-sensor-fusion-with-attention-rajeevatla_b200/libmsf_b200_timeline.so
MSF_B200_LIB=$TL timeout 300 python scripts/step_timeline.py 4096 > $O/timeline27.txt 2>&1; echo "timeline rc=$?"
grep -A10 "step 2" $O/timeline27.txt | cut -c1-160
MSF_CS_PAD=1 MSF_B200_LIB=$TL timeout 300 python scripts/step_timeline.py 4096 > $O/timeline27_pad.txt 2>&1; echo "timeline rc=$?"
grep -A10 "step 2" $O/timeline27_pad.txt | cut -c1-160
for v in "" "MSF_CS_PAD=1"; do
env $v timeout 600 python bench.py --no-cpu-baseline --no-strong > $O/bench27_$v.json 2> $O/bench27_$v.err; echo "bench rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2/bench27*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["value"], "e2e", d["e2e"]["ms_per_step"])
    except Exception as e: print(f, "ERR", e)
PY
