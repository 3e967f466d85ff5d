#!/bin/bash
# A/B: one vs two epilogue warpgroups in conv_tc (per-key conv table)
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-infer --conv-table > gpurun_out/r02_eg1.json 2> gpurun_out/r02_eg1.err
MMPL_LIB=$PWD/multimodal-pl_b200/libmmpl_b200_eg2.so timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-infer --conv-table > gpurun_out/r02_eg2.json 2> gpurun_out/r02_eg2.err
MMPL_TC_EPI_GROUPS=2 MMPL_LIB=$PWD/multimodal-pl_b200/libmmpl_b200_eg2.so timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-infer --conv-table > gpurun_out/r02_eg2f.json 2> gpurun_out/r02_eg2f.err
echo done
