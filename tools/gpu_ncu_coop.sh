#!/bin/bash
# GPU box: one ncu --set full capture of the cooperative-tier kernel (C5 by default, one wave of 148 instances)
set -u
cfg=${1:-c5}
mkdir -p gpurun_out
python tools/coop_one.py $cfg 148 > gpurun_out/coop_one_$cfg.txt 2>&1; cat gpurun_out/coop_one_$cfg.txt
ncu --set full --clock-control none --import-source on -k regex:bo_solve_kernel -c 1 -f -o gpurun_out/coop_$cfg \
    python tools/coop_one.py $cfg 148 > gpurun_out/ncu_coop_$cfg.log 2>&1
tail -3 gpurun_out/ncu_coop_$cfg.log
ls -la gpurun_out/*.ncu-rep
