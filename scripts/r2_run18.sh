set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2; mkdir -p $O
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_dp.py -m gpu -q > $O/pytest18_dp.log 2>&1; echo "dp pytest rc=$?"
tail -5 $O/pytest18_dp.log | cut -c1-250
N=${NGPU:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --no-strong > $O/bench18_n$N.json 2> $O/bench18_n$N.err; echo "bench rc=$?"
tail -3 $O/bench18_n$N.err | cut -c1-300
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2/bench18*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        e=d["e2e"]
        print(f, d["n_gpus"], d["ms_per_step"], d["value"], "e2e", e["ms_per_step"], d.get("replicas_identical"), d["run"]["collective"][:60])
    except Exception as e: print(f, "ERR", e)
PY
