#!/bin/bash
# round 2: compute-sanitizer memcheck / initcheck over small solves on the resident tilings (4-, 8-, 16-, 18-warp tiles: CTA-barrier
# and barrier-free row exchange, 256-bit mailbox) and the generational kernel.  racecheck is not run on the resident kernel: its
# row exchange and mailbox are tagged data races BY DESIGN (the reader polls the row itself; see pvc_internal.h, namespace flow).
mkdir -p gpurun_out
O=gpurun_out/r02_sanitizer.txt
: > $O
for v in 67 60 63 65 47; do
  echo "== variant $v memcheck (FloorPlanScene 70x70, T=435)" >> $O
  timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python tools/gpu_small_case.py $v FloorPlanScene 2>&1 | grep -E "ERROR SUMMARY|Error|Invalid|^[0-9]+ " | head -6 >> $O
done
for v in 60 65 47; do
  echo "== variant $v memcheck (FloorPlanScene 300x300, 2 sources, T=120)" >> $O
  timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python tools/gpu_small_case.py $v FloorPlanScene 300 2 120 2>&1 | grep -E "ERROR SUMMARY|Error|Invalid|^[0-9]+ " | head -6 >> $O
done
for v in 0 65; do
  echo "== variant $v initcheck (FloorPlanScene 70x70)" >> $O
  timeout 900 compute-sanitizer --tool initcheck --print-limit 5 python tools/gpu_small_case.py $v FloorPlanScene 2>&1 | grep -E "ERROR SUMMARY|Uninitialized|^[0-9]+ " | head -5 >> $O
done
cat $O
