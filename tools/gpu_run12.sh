#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "fused_classifier_loss or forward_partial_loss or partial_loss" > gpurun_out/r02_head12.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02_head12.log
timeout 300 python tools/bench_kernels.py head > gpurun_out/r02_bk12.txt 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --conv-table > gpurun_out/r02_bench_g.json 2> gpurun_out/r02_bench_g.err
echo "bench rc=$?" >> gpurun_out/r02_head12.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-infer --two-step-head > gpurun_out/r02_bench_g2.json 2> gpurun_out/r02_bench_g2.err
echo done
