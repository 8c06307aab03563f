cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2x_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2x_pytest.log
timeout 600 python tools/accel_ab.py > gpurun_out/r2x_accel_ab.log 2>&1
timeout 600 python tools/shard_spread.py 32768 8 > gpurun_out/r2x_spread.log 2>&1
