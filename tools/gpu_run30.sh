#!/bin/bash
# final sanity on one GPU: smoke, whole GPU suite, default bench line
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?" > gpurun_out/r02_final.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputest_final.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02_final.log
timeout 900 python bench.py --conv-table > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
echo "bench rc=$?" >> gpurun_out/r02_final.log
