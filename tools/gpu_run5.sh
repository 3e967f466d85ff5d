#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_parity_strict.py::test_cfg2_full_extent_forward_loss_backward_vs_oracle 2>&1 | tail -40 > gpurun_out/r02_gputest5.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-infer --conv-table > gpurun_out/r02_bench_d.json 2> gpurun_out/r02_bench_d.err
python tools/prof_stem.py 3 > gpurun_out/plain_stem.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:stem_tc -s 3 -c 3 -o gpurun_out/r02_stem python tools/prof_stem.py 3 > gpurun_out/ncu_stem.log 2>&1
echo done
