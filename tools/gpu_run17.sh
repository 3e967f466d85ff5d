#!/bin/bash
# 1 GPU: the whole GPU suite, the default bench line (with CPU / library baselines), the reference arm, cfg5 and cfg1 records
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputest17.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02_gputest17.log
timeout 900 python bench.py --conv-table > gpurun_out/r02_bench_h.json 2> gpurun_out/r02_bench_h.err
echo "bench rc=$?" >> gpurun_out/r02_gputest17.log
timeout 600 python bench.py --workload cfg5 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_cfg5.json 2> gpurun_out/r02_bench_cfg5.err
timeout 600 python bench.py --workload cfg1 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_cfg1.json 2> gpurun_out/r02_bench_cfg1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err
echo done >> gpurun_out/r02_gputest17.log
