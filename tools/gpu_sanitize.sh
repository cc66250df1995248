#!/bin/bash
# compute-sanitizer passes over small solves: the default kernels (variant 47 and the small-tile 50 that "auto" picks here), the
# state-out-through-TMA experiment (49), the kernel without the publisher warp (40), a one-launch-per-4-steps kernel (22) and
# the generational kernel without warp specialisation (30); then a multi-tile, two-source case on the default.
for v in 47 50 49 40 22 30; do
  for tool in memcheck racecheck; do
    echo "== variant $v $tool (FloorPlanScene 70x70, T=435)"
    timeout 600 compute-sanitizer --tool $tool --print-limit 5 python tools/gpu_small_case.py $v FloorPlanScene 2>&1 | grep -E "ERROR SUMMARY|Error|RACECHECK SUMMARY|hazard|Invalid|^[0-9]+ " | head -8
  done
done
echo "== variant 47 memcheck (FloorPlanScene 300x300, 2 sources, T=120: 7 x 3 tiles, dependency counters, source groups)"
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python tools/gpu_small_case.py 47 FloorPlanScene 300 2 120 2>&1 | grep -E "ERROR SUMMARY|Error|Invalid|^[0-9]+ " | head -8
echo "== initcheck default"
timeout 600 compute-sanitizer --tool initcheck --print-limit 5 python tools/gpu_small_case.py 0 FloorPlanScene 2>&1 | grep -E "ERROR SUMMARY|Uninitialized" | head -5
