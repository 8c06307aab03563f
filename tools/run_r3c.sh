cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_accel.py tests/test_gpu_parity.py -x -q -k "accel or schedule or sharding or resolve" > gpurun_out/r3c_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r3c_tests.log
timeout 600 python tools/accel_ab.py > gpurun_out/r3c_accel_ab.log 2>&1
