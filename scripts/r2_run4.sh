set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
export TL=$PWD/multimodal-sensor-fusion-with-attention-rajeevatla_b200/libmsf_b200_timeline.so
timeout 600 python -m pytest tests/test_gpu_fusion_bf16.py tests/test_gpu_head_fused.py tests/test_gpu_fullwidth.py -m gpu -x -q > gpurun_out/r2/pytest7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest4.log
tail -15 gpurun_out/r2/pytest7.log
MSF_B200_LIB=$TL timeout 300 python scripts/chain_stamps.py > gpurun_out/r2/stamps7.txt 2>&1; tail -16 gpurun_out/r2/stamps4.txt
for c in 4 1; do
  MSF_CHAIN_CLUSTER=$c timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/r2/bench7_c$c.json 2> gpurun_out/r2/bench7_c$c.err; echo "bench c$c rc=$?"
done
MSF_B200_LIB=$TL timeout 300 python scripts/step_timeline.py > gpurun_out/r2/timeline7.txt 2>&1; echo "timeline rc=$?"
for f in gpurun_out/r2/bench7_c*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["ms_per_step"], d["value"], d["roofline"]["avg_launch_us"], d["roofline"]["frac"], [ (p["launch"][:12],p["us_per_launch"]) for p in d["roofline"]["per_launch"]])
except Exception as e: print("ERR", e)
PY
done
grep -A9 "step 2" gpurun_out/r2/timeline7.txt | cut -c1-150
