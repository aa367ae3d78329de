set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/pytest17.log 2>&1; echo "pytest rc=$?" >> $O/pytest17.log
tail -8 $O/pytest17.log | cut -c1-220
timeout 600 python bench.py --no-cpu-baseline --no-strong > $O/bench17.json 2> $O/bench17.err; echo "bench rc=$?"; tail -3 $O/bench17.err | cut -c1-300
timeout 600 python bench.py --no-cpu-baseline --no-strong --packed-host > $O/bench17_packed.json 2> $O/bench17_packed.err; echo "bench rc=$?"
timeout 600 python bench.py --no-cpu-baseline --no-strong --host-dtype fp32 > $O/bench17_fp32.json 2> $O/bench17_fp32.err; echo "bench rc=$?"
timeout 600 python bench.py --no-cpu-baseline --no-strong --host-dtype fp32 --packed-host > $O/bench17_fp32_packed.json 2> $O/bench17_fp32_packed.err; echo "bench rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2/bench17*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        e=d["e2e"]
        print(f, d["ms_per_step"], "e2e", e["ms_per_step"], e["h2d_bytes_per_step"], e.get("host_feature_dtype"), e.get("host_batch"))
    except Exception as e: print(f, "ERR", e)
PY
