cd $GRAFT_REPO_ROOT
O=gpurun_out/r2; mkdir -p $O
N=${NGPU:-2}
nvidia-smi topo -m | head -12
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 scripts/mc_probe.py > $O/mc_probe.txt 2>&1; grep "^rank" $O/mc_probe.txt | cut -c1-330; tail -5 $O/mc_probe.txt | cut -c1-200
