#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu.py -x -q -m gpu -k "joint_space or mpc or c3 or c4 or c5 or axis" > gpurun_out/pytest_jsp.log 2>&1; tail -3 gpurun_out/pytest_jsp.log
