cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r3a_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r3a_pytest.log
timeout 300 python tools/step_overhead.py 32768 1 > gpurun_out/r3a_overhead1.log 2>&1
timeout 300 python tools/step_overhead.py 32768 8 > gpurun_out/r3a_overhead8.log 2>&1
