#!/bin/bash
mkdir -p gpurun_out
python tools/prof_loss.py 3 > gpurun_out/plain_loss.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"partial_loss|cls_" -s 8 -c 4 -o gpurun_out/r02_loss_staged python tools/prof_loss.py 3 > gpurun_out/ncu_loss1.log 2>&1
MMPL_LOSS_STAGED=0 ncu --set full --clock-control none --import-source on -k regex:"partial_loss" -s 4 -c 2 -o gpurun_out/r02_loss_reg python tools/prof_loss.py 3 > gpurun_out/ncu_loss2.log 2>&1
echo done
