#!/bin/bash
# flow-exchange resident tilings with 12 / 14 warps (70 / 71, one CTA per SM) beside 68 / 69 / 63 / 65
# (variants 60 / 61 / 62 / 68 are compiled only with  make EXTRA=-DPVC_ALL_VARIANTS  since the end of the round; 73 / 74 were experiments and are gone)
for cfg in "Shoebox 640 1000 1" "FloorPlanScene 768 1000 1" "FloorPlanScene 896 1000 1" "FloorPlanScene 1024 1000 1" "FloorPlanScene 512 1000 2" "FloorPlanScene 256 1000 8" "FloorPlanScene 300 1000 4" "FloorPlanScene 512 1000 4"; do
  for v in 0 69 68 70 71 63; do
    python tools/gpu_time_one.py $cfg $v 5 2>&1 | tail -1 | cut -c1-120
  done
done
