cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2m_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2m_pytest.log
timeout 600 python bench.py > gpurun_out/r2m_bench_n1.json 2> gpurun_out/r2m_bench_n1.err
