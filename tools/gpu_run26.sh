#!/bin/bash
# final single-GPU record: whole GPU suite, default bench line (all baselines), reference arm, cfg5 / cfg1 records
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputest_final.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02_gputest_final.log
timeout 900 python bench.py --conv-table > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
echo "bench rc=$?" >> gpurun_out/r02_gputest_final.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-infer > gpurun_out/r02_bench_n1_20.json 2> gpurun_out/r02_bench_n1_20.err
timeout 600 python bench.py --workload cfg5 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_cfg5.json 2> gpurun_out/r02_bench_cfg5.err
timeout 600 python bench.py --workload cfg1 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_cfg1.json 2> gpurun_out/r02_bench_cfg1.err
timeout 600 python bench.py --workload cfg4 --steps 3 --warmup 2 > gpurun_out/r02_bench_cfg4.json 2> gpurun_out/r02_bench_cfg4.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02_gputest_final.log
echo done >> gpurun_out/r02_gputest_final.log
