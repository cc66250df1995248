#!/bin/bash
# round 2, job 4: edge-case diagnosis, mailbox hand-over with fixed register layout, mbarrier edge sync variants, per-CTA busy times
mkdir -p gpurun_out
O=gpurun_out/r02_job4.txt
: > $O
echo "== edge-case diagnosis" >> $O
timeout 600 python tools/gpu_diag_edge.py >> $O 2>&1
echo "== timing" >> $O
for cfg in "BigRoom 1024 1000 1 64" "BigRoom 1024 1000 1 65" "BigRoom 1024 1000 1 70" "BigRoom 1024 1000 1 71" "BigRoom 1024 1000 1 61" "BigRoom 1024 1000 1 67" \
           "BigRoom 1024 1000 4 65" "BigRoom 1024 1000 4 71" "BigRoom 1024 1000 4 47" "FloorPlanScene 1024 1000 4 65" "FloorPlanScene 1024 1000 4 71" \
           "Shoebox 512 2000 1 60" "Shoebox 512 2000 1 66" "Shoebox 512 2000 1 61" "Shoebox 512 2000 1 67" "Shoebox 512 2000 1 63" "Shoebox 512 2000 1 69" \
           "FloorPlanScene 0 0 1 60" "FloorPlanScene 0 0 1 66" "FloorPlanScene 0 0 1 64" "FloorPlanScene 0 0 1 70" "FloorPlanScene 0 0 1 69"; do
  timeout 120 python tools/gpu_time_one.py $cfg 2>&1 | tail -1 >> $O
done
echo "== traces (tuning build)" >> $O
export PVC_LIB_PATH=$PWD/planeverb_b200/lib_tune/libplaneverb_b200.so
for cfg in "BigRoom 1024 1000 1 65" "BigRoom 1024 1000 1 71" "none 1024 1000 1 65" "Shoebox 512 2000 1 60" "Shoebox 512 2000 1 66" "FloorPlanScene 0 0 1 60"; do
  PVC_RES_TRACE=1 timeout 120 python tools/gpu_time_one.py $cfg 2>&1 | tail -3 >> $O
done
for dbg in 1 2; do
  for cfg in "BigRoom 1024 1000 1 65" "BigRoom 1024 1000 1 71" "none 1024 1000 1 65"; do
    PVC_RES_DEBUG=$dbg timeout 120 python tools/gpu_time_one.py $cfg >> $O 2>&1
  done
done
unset PVC_LIB_PATH
cut -c1-400 $O | tail -150
