#!/bin/bash
# One `ncu --set full` capture per hot kernel of the 256^3 benchmark step (run under gpurun, 1 GPU).
# usage: tools/ncu_full.sh <tag> [kernel:skip ...]   -> gpurun_out/prof_<tag>_<kernel>.ncu-rep
# (gpurun brings back at most 64 MiB: capture four kernels per call)
TAG=${1:-r01}; shift
SPECS=${@:-"k_fz:120 k_fyf:100 k_fx:100 k_fyi:100"}
CMD="python bench.py --grid 256 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
for spec in $SPECS; do
  k=${spec%%:*}; s=${spec##*:}
  ncu --set full --clock-control none -k regex:$k -s $s -c 1 -f -o gpurun_out/prof_${TAG}_$k $CMD > gpurun_out/ncu_${TAG}_$k.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
