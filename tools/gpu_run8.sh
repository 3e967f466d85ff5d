#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_parity_strict.py::test_cfg2_full_extent_forward_loss_backward_vs_oracle 2>&1 | tail -30 > gpurun_out/r02_gputest8.log
timeout 300 python tools/bench_kernels.py stem > gpurun_out/r02_bk8.txt 2>&1
timeout 300 python tools/bench_kernels.py partial >> gpurun_out/r02_bk8.txt 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --conv-table > gpurun_out/r02_bench_f.json 2> gpurun_out/r02_bench_f.err
echo done
