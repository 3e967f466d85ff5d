#!/bin/bash
# A/B: conv_tc with the GroupNorm-backward epilogue code compiled out (instruction-cache hypothesis)
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-infer --conv-table > gpurun_out/r02_ic0.json 2> gpurun_out/r02_ic0.err
MMPL_LIB=$PWD/multimodal-pl_b200/libmmpl_b200_nogn.so timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-infer --conv-table > gpurun_out/r02_ic1.json 2> gpurun_out/r02_ic1.err
echo done
