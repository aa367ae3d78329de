cd $GRAFT_REPO_ROOT
for d in 16 17 18 20 24 31; do MSF_LSTM_DBG=$d timeout -k 10 120 python scripts/lstm_prof.py 4096 256 2>&1 | grep -v Warning | tail -2 | cut -c1-300; done
MSF_LSTM_STEPS=1 timeout -k 10 120 python scripts/lstm_prof.py 4096 256 2>&1 | tail -1
timeout -k 10 120 python scripts/lstm_prof.py 1024 256 2>&1 | tail -1
