#!/bin/bash
# One `ncu --set full` capture per hot kernel of the 256^3 benchmark step (run under gpurun, 1 GPU).
# usage: tools/ncu_full.sh <tag>      -> gpurun_out/prof_<tag>_<kernel>.ncu-rep
TAG=${1:-r01}
CMD="python bench.py --grid 256 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
for spec in "k_fz:120" "k_fy:200" "k_fx:100" "k_iz:100" "k_update_mm10:4" "k_cg_update:60" "k_pk1_tangent:4"; do
  k=${spec%%:*}; s=${spec##*:}
  ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c 1 -f -o gpurun_out/prof_${TAG}_$k $CMD > gpurun_out/ncu_${TAG}_$k.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
