#!/bin/bash
# first GPU pass of round 2: box facts, GPU test suite (new strict parity tests verbose), short bench
mkdir -p gpurun_out
{ nproc; free -g | head -2; nvidia-smi --query-gpu=name,memory.total --format=csv; } > gpurun_out/r02_box.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x --durations=15 -rA 2>&1 | tail -150 > gpurun_out/r02_gputest1.log
timeout 600 python -m pytest tests/test_gpu_parity_strict.py -m gpu -q -s 2>&1 | tail -80 > gpurun_out/r02_strict.log
timeout 600 python bench.py --steps 10 --warmup 3 --conv-table > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err
echo done
