cd $GRAFT_REPO_ROOT
O=gpurun_out/r2; mkdir -p $O
N=${NGPU:-2}
CUDA_VISIBLE_DEVICES=0 timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest26.log 2>&1; echo "pytest rc=$?"
tail -3 $O/pytest26.log | cut -c1-250
timeout 600 python -m pytest tests/test_gpu_dp.py -m gpu -q > $O/pytest26_dp.log 2>&1; echo "dp pytest rc=$?"
tail -3 $O/pytest26_dp.log | cut -c1-250
DP_COMM=zshard timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 scripts/dp_phase_times.py > $O/dp_phase26.txt 2>&1; grep "^rank" $O/dp_phase26.txt | cut -c1-330
TL=$PWD/multimodal-sensor-fusion-with-attention-rajeevatla_b200/libmsf_b200_timeline.so
CUDA_VISIBLE_DEVICES=0 MSF_B200_LIB=$TL timeout 300 python scripts/step_timeline.py 4096 > $O/timeline26.txt 2>&1; echo "timeline rc=$?"
grep -A10 "step 2" $O/timeline26.txt | cut -c1-160
