cd $GRAFT_REPO_ROOT
timeout 900 python bench.py > gpurun_out/r3n_bench_n1.json 2> gpurun_out/r3n_bench_n1.err
timeout 600 python bench.py --workload symdpr1 --n 16384 --no-cpu > gpurun_out/r3n_bench_symdpr1.json 2> gpurun_out/r3n_bench_symdpr1.err
timeout 600 python bench.py --workload halfarrow --n 16384 --no-cpu > gpurun_out/r3n_bench_halfarrow.json 2> gpurun_out/r3n_bench_halfarrow.err
timeout 600 python bench.py --workload tdc --n 8192 > gpurun_out/r3n_bench_tdc.json 2> gpurun_out/r3n_bench_tdc.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r3n_bench_reference.json 2> gpurun_out/r3n_bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r3n_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-target > gpurun_out/r3n_ncu_launch.log 2>&1
for spec in "symarrow 32768" "symdpr1 16384" "halfarrow 16384"; do
  set -- $spec
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ -o /tmp/r3n_$1 -f python tools/prof_kinds.py $1 $2 1 > gpurun_out/r3n_ncu_$1.log 2>&1
  python tools/ncu_summary.py report /tmp/r3n_$1.ncu-rep gpurun_out/r3n_$1_full_n$2.md >> gpurun_out/r3n_ncu_$1.log 2>&1
  ncu -i /tmp/r3n_$1.ncu-rep --page raw --csv > gpurun_out/r3n_$1_raw.csv 2>/dev/null
done
timeout 600 python tools/parity_report.py > gpurun_out/r3n_parity.json 2> gpurun_out/r3n_parity.err
timeout 300 python tools/steps_hist.py 32768 8 3 > gpurun_out/r3n_hist.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3n_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/r3n_smoke.log
