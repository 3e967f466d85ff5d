#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "not nccl and not sharded" > gpurun_out/r02_t18.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02_t18.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-infer --conv-table > gpurun_out/r02_bench_i.json 2> gpurun_out/r02_bench_i.err
MMPL_COMPACT_DS=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-infer > gpurun_out/r02_bench_i0.json 2> gpurun_out/r02_bench_i0.err
echo done >> gpurun_out/r02_t18.log
