cd $GRAFT_REPO_ROOT
for q in -1 4 2; do
  echo "== quantum $q" >> gpurun_out/r2o_spread.log
  ARROWHEAD_B200_QUANTUM=$q timeout 600 python tools/shard_spread.py 32768 8 >> gpurun_out/r2o_spread.log 2>&1
done
timeout 300 python -m pytest tests/test_gpu_accel.py tests/test_gpu_parity.py -x -q -k "accel or schedule or sharding or resolve" > gpurun_out/r2o_tests.log 2>&1
