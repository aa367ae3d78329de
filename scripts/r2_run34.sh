cd $GRAFT_REPO_ROOT
O=gpurun_out/r2; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_head_fused.py -m gpu -q -x > $O/pytest34.log 2>&1; echo "pytest rc=$?"
grep -v "^$" $O/pytest34.log | tail -25 | cut -c1-250
timeout 600 python bench.py --workload infer_sweep > $O/bench34_sweep.json 2> $O/bench34_sweep.err; echo "sweep rc=$?"; tail -3 $O/bench34_sweep.err | cut -c1-300
timeout 600 python bench.py --workload infer_sweep --no-fold --no-cpu-baseline > $O/bench34_sweep_nofold.json 2> /dev/null; echo "sweep nofold rc=$?"
timeout 900 python bench.py --workload raw_infer > $O/bench34_raw.json 2> $O/bench34_raw.err; echo "raw rc=$?"; tail -3 $O/bench34_raw.err | cut -c1-300
TL=$PWD/multimodal-sensor-fusion-with-attention-rajeevatla_b200/libmsf_b200_timeline.so
MSF_B200_LIB=$TL timeout 300 python scripts/step_timeline.py 4096 > $O/timeline34.txt 2>&1; echo "timeline rc=$?"
grep -A10 "step 2" $O/timeline34.txt | cut -c1-160
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2/bench34*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        e=d.get("e2e",{}); r=d.get("roofline",{})
        print(f.split('/')[-1], "ms", round(d.get("ms_per_step",0),4), "value", round(d.get("value",0)), "e2e", round(e.get("value",0)), "frac", r.get("frac"), "cpu", (d.get("cpu_baseline") or {}).get("value"), d.get("clocks"))
    except Exception as ex: print(f, "ERR", ex)
PY
