#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "not nccl and not sharded" > gpurun_out/r02_t20.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02_t20.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --conv-table > gpurun_out/r02_bench_j.json 2> gpurun_out/r02_bench_j.err
MMPL_PSPLIT_DIRECT=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-infer > gpurun_out/r02_bench_j0.json 2> gpurun_out/r02_bench_j0.err
timeout 300 python tools/prof_1x1.py 32 64 1 20 > gpurun_out/r02_1x1_plain.log 2>&1 || { echo plain failed; tail -5 gpurun_out/r02_1x1_plain.log; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_tc_kernel|wgrad_tc_kernel" -c 3 -o gpurun_out/r02_1x1 python tools/prof_1x1.py 32 64 1 1 > gpurun_out/r02_1x1_ncu.log 2>&1
echo "ncu rc=$?" >> gpurun_out/r02_t20.log
