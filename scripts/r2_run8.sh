set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
P=$PWD/multimodal-sensor-fusion-with-attention-rajeevatla_b200
for v in timeline c3_5_2_2 c3_4_4_2 c3_3_3_3; do
  MSF_B200_LIB=$P/libmsf_b200_$v.so timeout 200 python scripts/chain_stamps.py > gpurun_out/r2/stamps8_$v.txt 2>&1; echo "== $v"; tail -15 gpurun_out/r2/stamps8_$v.txt
  MSF_B200_LIB=$P/libmsf_b200_$v.so timeout 200 python scripts/step_timeline.py 2>&1 | grep -A9 "step 2" | grep chain3
done
MSF_CHAIN=v2 MSF_B200_LIB=$P/libmsf_b200_timeline.so timeout 200 python scripts/step_timeline.py 2>&1 | grep -A9 "step 2" | cut -c1-150
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-strong > gpurun_out/r2/plain8.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:chain3_kernel -s 4 -c 2 -o gpurun_out/r2/prof_chain3b python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-strong > gpurun_out/r2/ncu8.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r2/plain8.log | cut -c1-600
