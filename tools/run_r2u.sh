cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2u_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2u_pytest.log
timeout 900 python bench.py > gpurun_out/r2u_bench_n1.json 2> gpurun_out/r2u_bench_n1.err
timeout 600 python bench.py --workload symdpr1 --n 16384 --no-cpu > gpurun_out/r2u_bench_symdpr1.json 2> gpurun_out/r2u_bench_symdpr1.err
timeout 600 python bench.py --workload halfarrow --n 16384 --no-cpu > gpurun_out/r2u_bench_halfarrow.json 2> gpurun_out/r2u_bench_halfarrow.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2u_bench_reference.json 2> gpurun_out/r2u_bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2u_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-target > gpurun_out/r2u_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ -o gpurun_out/r2u_symarrow_n32768 -f python tools/prof_kinds.py symarrow 32768 1 > gpurun_out/r2u_ncu_full.log 2>&1
ls -la gpurun_out/*.ncu-rep
