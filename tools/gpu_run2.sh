#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=10 2>&1 | tail -60 > gpurun_out/r02_gputest2.log
timeout 300 python tools/bench_kernels.py loss > gpurun_out/r02_bk_loss.txt 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 --conv-table > gpurun_out/r02_bench_b.json 2> gpurun_out/r02_bench_b.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_b_ref.json 2> gpurun_out/r02_bench_b_ref.err
echo done
