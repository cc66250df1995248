#!/bin/bash
# two-CTA-per-SM resident tilings with the barrier-free row exchange (68 = 10 warps, 69 = 8 warps) against the CTA-barrier ones (61, 60)
# (variants 60 / 61 / 62 / 68 are compiled only with  make EXTRA=-DPVC_ALL_VARIANTS  since the end of the round; 73 / 74 were experiments and are gone)
# and whatever the automatic selection picks (0), over the grids where the small tilings are candidates
for cfg in "Shoebox 256 1000 1" "Shoebox 384 1000 1" "Shoebox 512 2000 1" "Shoebox 640 1000 1" "FloorPlanScene 768 1000 1" "FloorPlanScene 256 1000 8" "FloorPlanScene 512 1000 2" "FloorPlanScene 191 1187 1" "FloorPlanScene 300 1000 4"; do
  for v in 0 60 69 68; do
    python tools/gpu_time_one.py $cfg $v 5 2>&1 | tail -1 | cut -c1-150
  done
done
