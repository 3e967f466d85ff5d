#!/bin/bash
# 2 GPUs: multi-rank tests (in-graph NCCL, sharded sliding window) under a hard timeout, then the N=2 bench line
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_train.py -m gpu -x -q -k "nccl or sharded" > gpurun_out/r02_n2_tests10.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02_n2_tests10.log
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench_n2b.json 2> gpurun_out/r02_bench_n2b.err
echo "bench rc=$?" >> gpurun_out/r02_n2_tests10.log
echo done
