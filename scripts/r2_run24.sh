cd $GRAFT_REPO_ROOT
O=gpurun_out/r2; mkdir -p $O
N=${NGPU:-8}
for mcast in 1 0; do
MSF_DP_MULTICAST=$mcast DP_COMM=zshard timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 scripts/dp_phase_times.py > $O/dp_phase24_n${N}_$mcast.txt 2>&1; grep "^rank [01]/" $O/dp_phase24_n${N}_$mcast.txt | cut -c1-330
done
DP_COMM=p2p timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 scripts/dp_phase_times.py > $O/dp_phase24_n${N}_p2p.txt 2>&1; grep "^rank [01]/" $O/dp_phase24_n${N}_p2p.txt | cut -c1-330
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N > $O/bench24_n$N.json 2> $O/bench24_n$N.err; echo "bench rc=$?"
tail -3 $O/bench24_n$N.err | cut -c1-300
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2/bench24*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        e=d["e2e"]
        print(f, d["n_gpus"], d["ms_per_step"], d["value"], "e2e", e["ms_per_step"], d.get("replicas_identical"), d["run"]["collective"][:60], d.get("strong_32768"))
    except Exception as e: print(f, "ERR", e)
PY
