#!/bin/bash
# GPU box: phase breakdown of the cooperative tier at several CTA sizes
set -u
mkdir -p gpurun_out
for c in c5 c4; do for t in 128 512; do
  python tools/coop_profile.py $c $t > gpurun_out/coop_phases_${c}_tpb$t.txt 2>&1
  echo "== $c tpb $t"; grep -E "kkt tape|factor  |solves  |assembly|total  |per iteration [0-9]" gpurun_out/coop_phases_${c}_tpb$t.txt | cut -c1-160
done; done
