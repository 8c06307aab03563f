cd $GRAFT_REPO_ROOT
run() { # name, env...
  name=$1; shift
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 40 --warmup 3 --no-e2e --no-allgather > gpurun_out/r2w_$name.json 2> gpurun_out/r2w_$name.err
}
run spin_s10 ARROWHEAD_B200_SPINWAIT=1
run block_s10 ARROWHEAD_B200_SPINWAIT=0
run spin_s50 ARROWHEAD_B200_SPINWAIT=1 BENCH_SAMPLER_INTERVAL=0.05
run block_s50 ARROWHEAD_B200_SPINWAIT=0 BENCH_SAMPLER_INTERVAL=0.05
run spin_s10_b ARROWHEAD_B200_SPINWAIT=1
