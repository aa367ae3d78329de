set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
nvidia-smi -L > gpurun_out/r2/gpu.txt
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest1.log
tail -5 gpurun_out/r2/pytest1.log
for v in default v2; do
  if [ $v = v2 ]; then export MSF_CHAIN=v2; else unset MSF_CHAIN; fi
  timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/r2/bench_$v.json 2> gpurun_out/r2/bench_$v.err; echo "bench $v rc=$?"
done
unset MSF_CHAIN
for c in 1 2 8; do
  MSF_CHAIN_CLUSTER=$c timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/r2/bench_c$c.json 2> gpurun_out/r2/bench_c$c.err; echo "bench c$c rc=$?"
done
MSF_B200_LIB=$PWD/multimodal-sensor-fusion-with-attention-rajeevatla_b200/libmsf_b200_timeline.so timeout 300 python scripts/step_timeline.py > gpurun_out/r2/timeline1.txt 2>&1; echo "timeline rc=$?"
for f in gpurun_out/r2/bench_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["ms_per_step"], d["value"], d["roofline"]["avg_launch_us"], d["roofline"]["frac"], [ (p["launch"][:12],p["us_per_launch"]) for p in d["roofline"]["per_launch"]])
except Exception as e: print("ERR", e)
PY
done
tail -30 gpurun_out/r2/timeline1.txt
