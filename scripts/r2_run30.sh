cd $GRAFT_REPO_ROOT
O=gpurun_out/r2; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_encoders.py tests/test_gpu_attention.py -m gpu -q > $O/pytest30.log 2>&1; echo "pytest rc=$?"
grep -v "^$" $O/pytest30.log | tail -40 | cut -c1-250
