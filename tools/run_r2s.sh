cd $GRAFT_REPO_ROOT
timeout 300 python tools/steps_hist.py 32768 8 3 > gpurun_out/r2s_hist.log 2>&1
