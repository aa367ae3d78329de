cd $GRAFT_REPO_ROOT
O=gpurun_out/r2; mkdir -p $O
timeout -k 10 240 python -m pytest tests/test_gpu_encoders.py -m gpu -q -x -k "tensor_core_lstm" > $O/pytest_lstm.log 2>&1; echo "pytest rc=$?"
grep -v "^$" $O/pytest_lstm.log | tail -25 | cut -c1-250
timeout -k 10 600 python bench.py --workload raw_infer --no-cpu-baseline > $O/bench_raw_lstmseq.json 2> $O/bench_raw_lstmseq.err; echo "raw rc=$?"; tail -3 $O/bench_raw_lstmseq.err | cut -c1-300
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r2/bench_raw_lstmseq.json").read().strip().splitlines()[-1])
    print("ms", d["ms_per_step"], "value", d["value"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], "lib", d["library_recurrence"]["ms_per_step"], "diff", d["max_abs_logit_diff_vs_library"])
except Exception as e: print("ERR", e)
PY
