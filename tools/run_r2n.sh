cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_accel.py -x -q -s > gpurun_out/r2n_accel_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2n_accel_tests.log
timeout 600 python tools/accel_ab.py > gpurun_out/r2n_accel_ab.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2n_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2n_pytest.log
