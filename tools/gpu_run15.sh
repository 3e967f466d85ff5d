#!/bin/bash
# launch list of one eager train step (ncu per-launch durations are cold-cache and serialised: compare shares)
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-infer --eager"
timeout 600 $CMD > gpurun_out/r02_eager_plain.json 2> gpurun_out/r02_eager_plain.err || { echo "plain run failed"; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/r02_launches_ncu.log 2>&1
echo "ncu rc=$?"
