#!/bin/bash
# GPU box: ncu --set full captures of the two kernels of the headline bench (K1 streaming evaluation, C2 solve kernel)
set -u
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:bo_eval_kernel -s 3 -c 1 -f -o gpurun_out/fk_jac \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/ncu_fk.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bo_solve_kernel -s 2 -c 1 -f -o gpurun_out/solve \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/ncu_solve.log 2>&1
ls -la gpurun_out/*.ncu-rep
