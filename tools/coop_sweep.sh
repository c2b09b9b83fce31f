#!/bin/bash
# sweep of the factor program's lanes-per-target (G) and warps (W) with the profiling build.
# usage: tools/coop_sweep.sh compile|run
mode=${1:-run}
for cfg in "c4 2 8" "c4 8 8" "c4 4 4" "c4 2 4" "c5 2 8" "c5 8 8" "c5 4 4" "c5 2 4"; do
  set -- $cfg
  export BO_FAC_G=$2 BO_FAC_WARPS=$3
  if [ "$mode" = "compile" ]; then
    python tools/coop_profile.py $1 0 compile | tail -1 | grep -o "'ldl_g': [0-9]*\|'ldl_warps': [0-9]*\|'factor_steps': [0-9]*" | paste - - -
  else
    echo "=== $1 fac_g=$2 fac_warps=$3"
    timeout 120 python tools/coop_profile.py $1 | grep -E "factor  |total|per iteration [0-9]"
  fi
done
