#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "not nccl and not sharded" > gpurun_out/r02_t23.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02_t23.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --conv-table > gpurun_out/r02_bench_k.json 2> gpurun_out/r02_bench_k.err
echo "bench rc=$?" >> gpurun_out/r02_t23.log
