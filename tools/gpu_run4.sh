#!/bin/bash
# 2-GPU validation: NCCL DP tests (incl. in-graph exchange), sharded sliding window, bench at N=2 (train + infer record)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q -k "nccl or sharded" 2>&1 | tail -40 > gpurun_out/r02_n2_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err
timeout 600 python -m pytest tests/test_gpu_more.py tests/test_gpu_kernels.py -m gpu -q -k "feam3 or stem or classifier" 2>&1 | tail -30 > gpurun_out/r02_feam3.log
echo done
