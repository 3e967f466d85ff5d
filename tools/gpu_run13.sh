#!/bin/bash
# 2 GPUs: sharded sliding window with the plane exchange (test + bench record), fused head at N=2
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_train.py -m gpu -x -q -k "nccl or sharded" > gpurun_out/r02_n2_tests13.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02_n2_tests13.log
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench_n2c.json 2> gpurun_out/r02_bench_n2c.err
echo "bench rc=$?" >> gpurun_out/r02_n2_tests13.log
echo done
