cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_accel.py -x -q -k fuzz > gpurun_out/r2y_fuzz.log 2>&1; echo "rc=$?" >> gpurun_out/r2y_fuzz.log
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "tdc" > gpurun_out/r2y_tdc_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2y_tdc_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2y_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/r2y_smoke.log
timeout 600 python bench.py --workload tdc --n 8192 > gpurun_out/r2y_bench_tdc.json 2> gpurun_out/r2y_bench_tdc.err
