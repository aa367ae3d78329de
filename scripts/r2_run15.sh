set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2; mkdir -p $O
TL=$PWD/multimodal-sensor-fusion-with-attention-rajeevatla_b200/libmsf_b200_timeline.so
for B in 4096 2048 8192; do
MSF_TL_CTAS=1 MSF_B200_LIB=$TL timeout 300 python scripts/step_timeline.py $B > $O/timeline15_$B.txt 2>&1; echo "timeline rc=$?"
MSF_WG=v1 MSF_B200_LIB=$TL timeout 300 python scripts/step_timeline.py $B > $O/timeline15_v1_$B.txt 2>&1; echo "timeline rc=$?"
grep -A9 "step 2" $O/timeline15_$B.txt | cut -c1-160
grep -A9 "step 2" $O/timeline15_v1_$B.txt | cut -c1-160
done
python scripts/wg_prof_target.py > $O/plain15.log 2>&1 &&
ncu --set full --clock-control none --cache-control none --import-source on -k regex:wg2_kernel -s 2 -c 2 -o $O/wg2_prof -f python scripts/wg_prof_target.py > $O/ncu15.log 2>&1
echo "ncu rc=$?"
MSF_WG=v1 ncu --set full --clock-control none --cache-control none --import-source on -k regex:tc_gemm_kernel -s 2 -c 2 -o $O/wgv1_prof -f python scripts/wg_prof_target.py > $O/ncu15b.log 2>&1
echo "ncu rc=$?"
