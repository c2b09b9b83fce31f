#!/bin/bash
# Run on the GPU box (under gpurun): tests, bench, and the ncu captures the profiles/ summaries come from.
set -u
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
# launch list of the bench command (cold-cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
# full captures of the two kernels
ncu --set full --clock-control none --import-source on -k regex:bo_eval_kernel -s 3 -c 1 -f -o gpurun_out/fk_jac \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_fk.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bo_solve_kernel -s 2 -c 1 -f -o gpurun_out/solve \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_solve.log 2>&1
ls -la gpurun_out
