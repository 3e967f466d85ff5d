#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "stem" 2>&1 | tail -40 > gpurun_out/r02_stem.log
timeout 900 python -m pytest tests/test_gpu_parity_strict.py -m gpu -q -k "fp32_all or base64" 2>&1 | grep -E "^E  |passed|failed|Error" | cut -c1-1500 > gpurun_out/r02_strict3.log
timeout 600 python -m pytest tests/test_gpu_sliding.py tests/test_gpu_unet.py tests/test_gpu_more.py -m gpu -q 2>&1 | tail -30 > gpurun_out/r02_sliding3.log
timeout 300 python tools/bench_kernels.py stem > gpurun_out/r02_bk_stem.txt 2>&1
timeout 300 python tools/bench_kernels.py partial > gpurun_out/r02_bk_loss2.txt 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --conv-table > gpurun_out/r02_bench_c.json 2> gpurun_out/r02_bench_c.err
echo done
