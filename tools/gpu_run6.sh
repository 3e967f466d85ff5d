#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fullsize.py -m gpu -q -k "stem or loss" 2>&1 | tail -15 > gpurun_out/r02_t6.log
timeout 300 python tools/bench_kernels.py stem > gpurun_out/r02_bk6.txt 2>&1
timeout 300 python tools/bench_kernels.py partial >> gpurun_out/r02_bk6.txt 2>&1
MMPL_LOSS_STAGED=0 timeout 300 python tools/bench_kernels.py partial >> gpurun_out/r02_bk6.txt 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-infer --conv-table > gpurun_out/r02_bench_e.json 2> gpurun_out/r02_bench_e.err
echo done
