cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r3d_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r3d_pytest.log
timeout 900 python bench.py > gpurun_out/r3d_bench_n1.json 2> gpurun_out/r3d_bench_n1.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3d_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/r3d_smoke.log
