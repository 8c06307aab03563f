cd $GRAFT_REPO_ROOT
timeout 600 python tools/shard_spread.py 32768 8 > gpurun_out/r2q_spread.log 2>&1
timeout 300 python tools/step_overhead.py 32768 8 > gpurun_out/r2q_overhead8.log 2>&1
timeout 300 python tools/step_overhead.py 32768 1 > gpurun_out/r2q_overhead1.log 2>&1
timeout 600 python -m pytest tests/test_gpu_accel.py tests/test_gpu_parity.py -x -q -k "accel or schedule or sharding or resolve" > gpurun_out/r2q_tests.log 2>&1
