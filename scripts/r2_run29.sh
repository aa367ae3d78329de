cd $GRAFT_REPO_ROOT
O=gpurun_out/r2; mkdir -p $O
timeout 300 python scripts/head_prof_target.py > $O/head_stamps29.txt 2>&1; echo "rc=$?"
cat $O/head_stamps29.txt | cut -c1-400
