cd $GRAFT_REPO_ROOT
O=gpurun_out/r2; mkdir -p $O
timeout -k 10 240 python -m pytest tests/test_gpu_encoders.py -m gpu -q -x -k "tensor_core_lstm" > $O/pytest_lstm.log 2>&1; echo "pytest rc=$?"
grep -v "^$" $O/pytest_lstm.log | tail -12 | cut -c1-250
for d in 16 24 27; do MSF_LSTM_DBG=$d timeout -k 10 120 python scripts/lstm_prof.py 4096 256 2>&1 | grep -v Warning | tail -2 | cut -c1-300; done
timeout -k 10 120 python scripts/lstm_prof.py 4096 1024 2>&1 | tail -1
