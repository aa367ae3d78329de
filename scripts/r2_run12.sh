set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
timeout 300 python -m pytest tests/test_gpu_layernorm.py tests/test_gpu_dp.py -m gpu -q > gpurun_out/r2/pytest12.log 2>&1; tail -5 gpurun_out/r2/pytest12.log | cut -c1-200
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dp_phase_times.py > gpurun_out/r2/dp_phase_n2.txt 2>&1; grep "^rank" gpurun_out/r2/dp_phase_n2.txt
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 200 --warmup 10 > gpurun_out/r2/bench12_n2.json 2> gpurun_out/r2/bench12_n2.err; echo "bench n2 rc=$?"; tail -3 gpurun_out/r2/bench12_n2.err | cut -c1-300
cat gpurun_out/r2/bench12_n2.json | cut -c1-1500
