#!/bin/bash
# 8 GPUs: how much of the N=8 step is NCCL taking SMs from the persistent conv kernels?  One bench line per variant.
mkdir -p gpurun_out
run() {  # name, env..., -- extra bench args
  name=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 \
    bench.py --gpus 8 --steps 20 --warmup 5 --no-infer --no-cpu-baseline "$@" > gpurun_out/r02_n8_$name.json 2> gpurun_out/r02_n8_$name.err
  echo "$name rc=$?"
}
run chan4   NCCL_MAX_NCHANNELS=4 --
run thr128  NCCL_NTHREADS=128 NCCL_MIN_NCHANNELS=16 --
run bkt32   X=1 -- --bucket-mb 32
run outside MMPL_GRAPH_NCCL=0 --
# the full default line last (with the sliding-window record: plane exchange at 8 ranks)
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_bench_n8b.json 2> gpurun_out/r02_bench_n8b.err
echo "default rc=$?"
