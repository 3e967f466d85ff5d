#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "not nccl and not sharded" > gpurun_out/r02_t25.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02_t25.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --conv-table > gpurun_out/r02_bench_l.json 2> gpurun_out/r02_bench_l.err
echo "bench rc=$?" >> gpurun_out/r02_t25.log
MMPL_TC_PLANE_MAJOR=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-infer --conv-table > gpurun_out/r02_bench_l0.json 2> gpurun_out/r02_bench_l0.err
