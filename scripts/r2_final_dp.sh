# Round-2 data-parallel evidence: bench.py at N GPUs (weak scaling 4096 windows per GPU + the strong-scaling record)
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2final; mkdir -p $O
N=${NGPU:-2}
if [ "$N" = "2" ]; then
timeout 600 python -m pytest tests/test_gpu_dp.py -m gpu -q > $O/pytest_gpu_dp.txt 2>&1; echo "dp pytest rc=$?" >> $O/pytest_gpu_dp.txt; tail -2 $O/pytest_gpu_dp.txt
fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N > $O/bench_bf16_n$N.json 2> $O/bench_bf16_n$N.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r2final/bench_bf16_n$N.json").read().strip().splitlines()[-1])
e=d["e2e"]
print(d["n_gpus"], "ms", d["ms_per_step"], "value", d["value"], "e2e ms", e["ms_per_step"], "identical", d.get("replicas_identical"), d["run"]["collective"][:50], d.get("strong_32768"))
PY
