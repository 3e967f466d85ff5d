#!/bin/bash
mkdir -p gpurun_out
MMPL_SW_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 \
  bench.py --gpus 8 --workload cfg4 --steps 2 --warmup 2 > gpurun_out/r02_sw8_trace.json 2> gpurun_out/r02_sw8_trace.err
echo "rc=$?"
grep "sw rank" gpurun_out/r02_sw8_trace.err | tail -16
