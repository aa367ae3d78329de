cd $GRAFT_REPO_ROOT
O=gpurun_out/r2; mkdir -p $O
N=${NGPU:-2}
if [ "$N" = "2" ]; then
timeout 600 python -m pytest tests/test_gpu_dp.py -m gpu -q > $O/pytest20_dp.log 2>&1; echo "dp pytest rc=$?"
tail -5 $O/pytest20_dp.log | cut -c1-250
fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N > $O/bench20_n$N.json 2> $O/bench20_n$N.err; echo "bench rc=$?"
tail -3 $O/bench20_n$N.err | cut -c1-300
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2/bench20*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        e=d["e2e"]
        print(f, d["n_gpus"], d["ms_per_step"], d["value"], "e2e", e["ms_per_step"], d.get("replicas_identical"), d["run"]["collective"][:60], d.get("strong_32768"))
    except Exception as e: print(f, "ERR", e)
PY
