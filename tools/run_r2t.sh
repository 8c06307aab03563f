cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_accel.py tests/test_gpu_parity.py -x -q -k "accel or schedule or sharding or resolve" > gpurun_out/r2t_tests.log 2>&1
timeout 600 python tools/accel_ab.py > gpurun_out/r2t_accel_ab.log 2>&1
timeout 600 python tools/shard_spread.py 32768 8 > gpurun_out/r2t_spread.log 2>&1
timeout 300 python tools/steps_hist.py 32768 8 3 > gpurun_out/r2t_hist.log 2>&1
