#!/bin/bash
# round 2, job 11: linear-form general path in the generational kernel, L2 persistence experiment, full GPU suite
mkdir -p gpurun_out
O=gpurun_out/r02_job11.txt
: > $O
echo "== full GPU suite" >> $O
timeout 1500 python -m pytest tests -m gpu -x -q --durations=6 2>&1 | tail -25 >> $O
echo "== generational kernel with the linear-form general path (was: BigRoom x4 7.62 ms, FloorPlan x4 10.0 ms, HugeRoom 768 x4 5.28)" >> $O
for cfg in "BigRoom 1024 1000 4 47" "FloorPlanScene 1024 1000 4 47" "HugeRoom 768 1000 4 47" "HugeRoom 2048 400 2 47" "FloorPlanScene 2048 400 2 47"; do
  timeout 200 python tools/gpu_time_one.py $cfg 2>&1 | tail -1 >> $O
done
echo "== L2 access-policy window on the ping-pong state (tuning build)" >> $O
export PVC_LIB_PATH=$PWD/planeverb_b200/lib_tune/libplaneverb_b200.so
for mb in 0 40 64 90; do
  for cfg in "BigRoom 1024 1000 4 47" "HugeRoom 2048 400 2 47"; do
    PVC_L2_PERSIST=$mb timeout 200 python tools/gpu_time_one.py $cfg 2>&1 | tail -2 | cut -c1-250 >> $O
  done
done
for g in 3 4; do
  PVC_L2_PERSIST=90 PVC_GROUP_SRC=$g timeout 200 python tools/gpu_time_one.py BigRoom 1024 1000 4 47 2>&1 | tail -1 >> $O
done
unset PVC_LIB_PATH
cut -c1-300 $O
