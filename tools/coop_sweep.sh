#!/bin/bash
# sweep of the cooperative tier's plan knobs with the profiling build.  usage: tools/coop_sweep.sh compile|run
mode=${1:-run}
export BO_SOLVE_MAX_WARPS=8
for cfg in "c5 1" "c5 2" "c5 4" "c5 8" "c4 2" "c4 4" "c4 8" "c3 2"; do
  set -- $cfg
  export BO_SOLVE_WARPS=$2
  if [ "$mode" = "compile" ]; then
    python tools/coop_profile.py $1 0 compile | tail -1 | grep -o "'solve_g': [0-9]*\|'solve_steps': [0-9]*" | paste - -
  else
    echo "=== $1 solve_warps=$2"
    timeout 120 python tools/coop_profile.py $1 | grep -E "solves|total|per iteration [0-9]|status"
  fi
done
