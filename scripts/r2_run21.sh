cd $GRAFT_REPO_ROOT
O=gpurun_out/r2; mkdir -p $O
N=${NGPU:-2}
for c in zshard p2p nccl; do
DP_COMM=$c timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 scripts/dp_phase_times.py > $O/dp_phase21_$c.txt 2>&1; grep "^rank" $O/dp_phase21_$c.txt | cut -c1-330
done
