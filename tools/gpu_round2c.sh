#!/bin/bash
# GPU box: all GPU tests, then the driver-style bench line (all configs)
set -u
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|rc=|converged \(1e-8\)|oracle KKT|polish|basin" gpurun_out/pytest_gpu.log | cut -c1-250
python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench.json"))
print("c2", d["value"], d["ms_per_step"], d["converged_fraction"], d["e2e"]["value"])
for k, c in d["configs"].items():
    print(k, {x: c.get(x) for x in ("value", "ms_per_step", "converged_fraction", "mean_iterations")}, c["e2e"]["value"])
PY
