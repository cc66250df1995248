#!/bin/bash
# round 2, end: compute-sanitizer memcheck / initcheck over the resident tilings added last (69 = 8 warps, two CTAs per SM; 72 / 70 / 71 =
# 10 / 12 / 14 warps, one CTA per SM; all with the barrier-free row exchange) and the 48-register analyzer instantiation (dense grids).
mkdir -p gpurun_out
O=gpurun_out/r02_sanitizer_c.txt
: > $O
run() { echo "== $1: $2" >> $O; shift 2; timeout 900 compute-sanitizer "$@" 2>&1 | grep -E "ERROR SUMMARY|Error|Invalid|Uninitialized|^[0-9]+ " | head -8 >> $O; }
for v in 69 72 70 71; do
  run "memcheck" "variant $v, FloorPlanScene 300x300, 2 sources, T=120" --tool memcheck --print-limit 5 python tools/gpu_small_case.py $v FloorPlanScene 300 2 120
done
run "initcheck" "variant 70, FloorPlanScene 300x300, 2 sources, T=120" --tool initcheck --print-limit 5 python tools/gpu_small_case.py 70 FloorPlanScene 300 2 120
run "memcheck" "auto, FloorPlanScene 1024x1024, T=120 (48-register analyzer instantiation, 18-warp tiles)" --tool memcheck --print-limit 5 python tools/gpu_small_case.py 0 FloorPlanScene 1024 1 120
cat $O
