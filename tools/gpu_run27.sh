#!/bin/bash
# N GPUs (first argument): multi-rank tests when N == 2, then the bench line launched the way the driver does
N=${1:-8}
mkdir -p gpurun_out
if [ "$N" = "2" ]; then
  timeout 420 python -m pytest tests/test_gpu_train.py -m gpu -x -q -k "nccl or sharded" > gpurun_out/r02_n2_tests_final.log 2>&1
  echo "tests rc=$?" >> gpurun_out/r02_n2_tests_final.log
fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
echo "bench rc=$?"
tail -2 gpurun_out/r02_bench_n$N.err
