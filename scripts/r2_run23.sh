cd $GRAFT_REPO_ROOT
O=gpurun_out/r2; mkdir -p $O
N=${NGPU:-2}
if [ "$N" = "2" ]; then
timeout 600 python -m pytest tests/test_gpu_dp.py -m gpu -q > $O/pytest23_dp.log 2>&1; echo "dp pytest rc=$?"
tail -5 $O/pytest23_dp.log | cut -c1-250
fi
for mcast in 1 0; do
MSF_DP_MULTICAST=$mcast DP_COMM=zshard timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 scripts/dp_phase_times.py > $O/dp_phase23_$mcast.txt 2>&1; grep "^rank" $O/dp_phase23_$mcast.txt | cut -c1-330
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --no-strong > $O/bench23_n$N.json 2> $O/bench23_n$N.err; echo "bench rc=$?"
tail -3 $O/bench23_n$N.err | cut -c1-300
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2/bench23*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        e=d["e2e"]
        print(f, d["n_gpus"], d["ms_per_step"], d["value"], "e2e", e["ms_per_step"], d.get("replicas_identical"), d["run"]["collective"][:60])
    except Exception as e: print(f, "ERR", e)
PY
