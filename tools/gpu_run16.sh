#!/bin/bash
# ncu --set full of the dominant tcgen05 conv kernel (full-resolution 32->32 block of cfg2: fprop, dgrad+GN-backward, wgrad)
mkdir -p gpurun_out
timeout 300 python tools/prof_block.py 32 0 2 > gpurun_out/r02_block_plain.log 2>&1 || { echo "plain run failed"; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_tc_kernel|wgrad_tc_kernel|gn_relu" -c 8 -o gpurun_out/r02_block python tools/prof_block.py 32 0 1 > gpurun_out/r02_block_ncu.log 2>&1
echo "ncu rc=$?"
