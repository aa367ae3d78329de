cd $GRAFT_REPO_ROOT
O=gpurun_out/r2; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/pytest33.log 2>&1; echo "pytest rc=$?"
grep -v "^$" $O/pytest33.log | tail -30 | cut -c1-250
