cd $GRAFT_REPO_ROOT
O=gpurun_out/r2; mkdir -p $O
N=${NGPU:-2}
timeout 600 python -m pytest tests/test_gpu_dp.py tests/test_gpu_gemm.py -m gpu -q > $O/pytest25_dp.log 2>&1; echo "dp pytest rc=$?"
tail -3 $O/pytest25_dp.log | cut -c1-250
for mcast in 1 0; do
MSF_DP_MULTICAST=$mcast DP_COMM=zshard timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 scripts/dp_phase_times.py > $O/dp_phase25_$mcast.txt 2>&1; grep "^rank" $O/dp_phase25_$mcast.txt | cut -c1-330
done
TL=$PWD/multimodal-sensor-fusion-with-attention-rajeevatla_b200/libmsf_b200_timeline.so
CUDA_VISIBLE_DEVICES=0 MSF_B200_LIB=$TL timeout 300 python scripts/step_timeline.py 4096 > $O/timeline25.txt 2>&1; echo "timeline rc=$?"
grep -A10 "step 2" $O/timeline25.txt | cut -c1-160
