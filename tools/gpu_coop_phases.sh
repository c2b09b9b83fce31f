#!/bin/bash
# GPU box: phase breakdown of the cooperative tier (profiling build); extra environment per variant in $VARIANTS ("A=1 B=2;C=3")
set -u
mkdir -p gpurun_out
IFS=';' read -ra VARS <<< "${VARIANTS:-default}"
for v in "${VARS[@]}"; do
  for c in ${CONFIGS:-c3 c5 c4}; do
    tag=$(echo "$v" | tr ' =' '__')
    if [ "$v" = "default" ]; then python tools/coop_profile.py $c > gpurun_out/coop_phases_${c}_$tag.txt 2>&1
    else env $v python tools/coop_profile.py $c > gpurun_out/coop_phases_${c}_$tag.txt 2>&1; fi
    echo "== $c [$v]"; grep -E "factor  |solves  |total  |per iteration [0-9]|'ldl_g'" gpurun_out/coop_phases_${c}_$tag.txt | sed "s/.*'ldl_g'/ ldl_g/" | cut -c1-200
  done
done
