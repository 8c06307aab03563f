set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k_pytest.log
python bench.py > gpurun_out/r2k_bench_n1.json 2> gpurun_out/r2k_bench_n1.err
python bench.py --workload symdpr1 --n 16384 --no-cpu > gpurun_out/r2k_bench_symdpr1.json 2> gpurun_out/r2k_bench_symdpr1.err
python bench.py --workload halfarrow --n 16384 --no-cpu > gpurun_out/r2k_bench_halfarrow.json 2> gpurun_out/r2k_bench_halfarrow.err
python tools/prof_kinds.py symdpr1 16384 1 > gpurun_out/r2k_prof_dpr1.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_ -o gpurun_out/r2k_symdpr1_n16384 -f python tools/prof_kinds.py symdpr1 16384 1 > gpurun_out/r2k_ncu_dpr1.log 2>&1
python tools/prof_kinds.py halfarrow 16384 1 > gpurun_out/r2k_prof_half.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_ -o gpurun_out/r2k_halfarrow_n16384 -f python tools/prof_kinds.py halfarrow 16384 1 > gpurun_out/r2k_ncu_half.log 2>&1
ls -la gpurun_out/*.ncu-rep
