#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_aux_nets.py -m gpu -q 2>&1 | grep -E "^E  |passed|failed|Error|error" | cut -c1-600 | tail -40 > gpurun_out/r02_aux9.log
timeout 300 python tools/bench_kernels.py stem > gpurun_out/r02_bk9.txt 2>&1
echo done
