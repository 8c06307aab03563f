cd $GRAFT_REPO_ROOT
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2r_bench_n8.json 2> gpurun_out/r2r_bench_n8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 10 --warmup 3 --n 65536 --no-e2e > gpurun_out/r2r_bench_n65536_g8.json 2> gpurun_out/r2r_bench_n65536_g8.err
python bench.py --steps 20 --warmup 3 --no-cpu --no-target > gpurun_out/r2r_bench_n1.json 2> gpurun_out/r2r_bench_n1.err
