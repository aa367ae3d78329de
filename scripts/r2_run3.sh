set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
export TL=$PWD/multimodal-sensor-fusion-with-attention-rajeevatla_b200/libmsf_b200_timeline.so
for c in 1 4; do
MSF_CHAIN_CLUSTER=$c MSF_B200_LIB=$TL timeout 300 python scripts/chain_stamps.py > gpurun_out/r2/stamps_c$c.txt 2>&1; echo "stamps rc=$?"
cat gpurun_out/r2/stamps_c$c.txt | tail -12
done
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:chain3_kernel -s 4 -c 2 -o gpurun_out/r2/prof_chain3 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2/ncu3.log 2>&1
echo "ncu rc=$?"
tail -5 gpurun_out/r2/ncu3.log
