#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/prof_block.py 64 1 2 > gpurun_out/r02_block64_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/r02_block64_plain.log; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_tc_kernel|wgrad_tc_kernel" -c 4 -o gpurun_out/r02_block64 python tools/prof_block.py 64 1 1 > gpurun_out/r02_block64_ncu.log 2>&1
echo "ncu rc=$?"
