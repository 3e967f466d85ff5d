#!/bin/bash
# end-of-round single-GPU record: full GPU test suite, smoke, default bench line, then the ncu launch lists of one eager
# train step and of one sliding-window volume (ncu per-launch durations are cold-cache and serialised: compare shares)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02f_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r02f_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02f_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02f_smoke.log
timeout 600 python bench.py > gpurun_out/r02f_bench_n1.json 2> gpurun_out/r02f_bench_n1.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02f_bench_ref.json 2> gpurun_out/r02f_bench_ref.err; echo "ref rc=$?"
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-infer --eager"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r02f_launches.csv $CMD > gpurun_out/r02f_launches_ncu.log 2>&1
echo "ncu train rc=$?"
CMD="python bench.py --workload cfg4 --steps 1 --warmup 1"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02f_launches_sw.csv $CMD > gpurun_out/r02f_launches_sw_ncu.log 2>&1
echo "ncu sw rc=$?"
