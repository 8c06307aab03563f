cd $GRAFT_REPO_ROOT
timeout 300 python tools/accel_ab.py 2>&1 | grep "accel=1" > gpurun_out/r3q_ab.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r3q_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r3q_pytest.log
