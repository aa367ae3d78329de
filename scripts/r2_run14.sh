set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2; mkdir -p $O
nvidia-smi -L > $O/gpu.txt
timeout 300 python -m pytest tests/test_gpu_gemm.py -m gpu -q -x > $O/pytest14_gemm.log 2>&1; echo "gemm pytest rc=$?"
tail -15 $O/pytest14_gemm.log | cut -c1-300
timeout 900 python -m pytest tests -m gpu -q > $O/pytest14.log 2>&1; echo "pytest rc=$?" >> $O/pytest14.log
tail -12 $O/pytest14.log | cut -c1-220
timeout 600 python bench.py --no-cpu-baseline --no-strong > $O/bench14.json 2> $O/bench14.err; echo "bench rc=$?"; tail -3 $O/bench14.err | cut -c1-300
MSF_WG=v1 timeout 600 python bench.py --no-cpu-baseline --no-strong > $O/bench14_wgv1.json 2> $O/bench14_wgv1.err; echo "bench v1 rc=$?"
TL=$PWD/multimodal-sensor-fusion-with-attention-rajeevatla_b200/libmsf_b200_timeline.so
MSF_B200_LIB=$TL timeout 300 python scripts/step_timeline.py > $O/timeline14.txt 2>&1; echo "timeline rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2/bench14*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d.get("roofline",{})
        print(f, d["ms_per_step"], d["value"], r.get("frac"), r.get("avg_launch_us"), d.get("e2e"))
        print([ (p["launch"][:14],p["us_per_launch"]) for p in r.get("per_launch",[])])
    except Exception as e: print(f, "ERR", e)
PY
grep -B1 -A12 "step 2" $O/timeline14.txt | cut -c1-160
