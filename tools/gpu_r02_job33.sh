#!/bin/bash
# what the automatic selection picks after the small flow tilings (69 / 72 / 70 / 71) joined the candidates, over plugin-sized grids and batches
for cfg in "FloorPlanScene 0 0 1" "FloorPlanScene 191 1187 1" "Shoebox 256 1000 1" "Shoebox 384 1000 1" "Shoebox 512 2000 1" "Shoebox 576 1000 1" "Shoebox 640 1000 1" "FloorPlanScene 768 1000 1" \
           "FloorPlanScene 896 1000 1" "FloorPlanScene 1024 1000 1" "FloorPlanScene 256 1000 8" "FloorPlanScene 300 1000 4" "FloorPlanScene 512 1000 2" "FloorPlanScene 512 1000 4" "FloorPlanScene 768 1000 4" "BigRoom 1024 4000 4"; do
  python tools/gpu_time_one.py $cfg 0 5 2>&1 | tail -1 | cut -c1-130
done
