#!/bin/bash
# compute-sanitizer passes over a small solve (default kernel, TMA persistent kernel, baseline kernel)
for v in 0 22 30; do
  for tool in memcheck racecheck; do
    echo "== variant $v $tool"
    compute-sanitizer --tool $tool --print-limit 5 python tools/gpu_small_case.py $v FloorPlanScene 2>&1 | grep -E "ERROR SUMMARY|Error|RACECHECK SUMMARY|hazard|Invalid|^[0-9]+ " | head -8
  done
done
echo "== initcheck default"
compute-sanitizer --tool initcheck --print-limit 5 python tools/gpu_small_case.py 0 FloorPlanScene 2>&1 | grep -E "ERROR SUMMARY|Uninitialized" | head -5
