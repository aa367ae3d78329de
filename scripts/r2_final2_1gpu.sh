# Round-2 closing evidence on one B200 (after the recurrence work): GPU tests, the reference's own test files against
# the drop-in, the main bench line, the raw-window lines, launch list of the recurrence kernels.
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2final2; mkdir -p $O
nvidia-smi -L > $O/gpu.txt
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.txt; tail -3 $O/pytest_gpu.txt
if [ -d ab_tmp/ref ]; then (cd ab_tmp/ref && timeout 600 python -m pytest tests -q -p no:cacheprovider > ../../$O/pytest_reference_suite.txt 2>&1; echo "reference suite rc=$?" >> ../../$O/pytest_reference_suite.txt); tail -3 $O/pytest_reference_suite.txt; fi
timeout 600 python bench.py > $O/bench_bf16.json 2> $O/bench_bf16.err; echo "bench rc=$?"
timeout 900 python bench.py --workload raw_infer > $O/bench_raw_infer.json 2> $O/bench_raw_infer.err; echo "raw_infer rc=$?"
timeout 900 python bench.py --workload raw_train > $O/bench_raw_train.json 2> $O/bench_raw_train.err; echo "raw_train rc=$?"
python scripts/lstm_prof_target.py 64 > $O/plain_lstm.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/lstm_launches.csv python scripts/lstm_prof_target.py 64 > $O/ncu_lstm_launches.log 2>&1
echo "ncu lstm launches rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2final2/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        e=d.get("e2e",{}); r=d.get("roofline",{})
        print(f.split('/')[-1], "ms", round(d.get("ms_per_step",0),4), "value", round(d.get("value",0)), d.get("unit"), "e2e", round(e.get("value",0)), "frac", r.get("frac"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
    except Exception as ex: print(f, "ERR", ex)
PY
