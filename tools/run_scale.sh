cd $GRAFT_REPO_ROOT
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r3o_bench_n8.json 2> gpurun_out/r3o_bench_n8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 10 --warmup 3 --size 65536 --no-e2e > gpurun_out/r3o_bench_n65536_g8.json 2> gpurun_out/r3o_bench_n65536_g8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 20 --warmup 3 --no-e2e > gpurun_out/r3o_bench_n4.json 2> gpurun_out/r3o_bench_n4.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e > gpurun_out/r3o_bench_n2.json 2> gpurun_out/r3o_bench_n2.err
