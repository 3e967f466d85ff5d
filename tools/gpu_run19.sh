#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/prof_1x1.py 32 64 1 20 > gpurun_out/r02_1x1_plain.log 2>&1 || { echo plain failed; tail -5 gpurun_out/r02_1x1_plain.log; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_tc_kernel|wgrad_tc_kernel" -c 3 -o gpurun_out/r02_1x1 python tools/prof_1x1.py 32 64 1 1 > gpurun_out/r02_1x1_ncu.log 2>&1
echo "ncu rc=$?"
