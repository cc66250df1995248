#!/bin/bash
# round 2, job 3: mailbox (LL) hand-over of the resident kernel: atomicity stress, parity, traces, timing sweep, full GPU suite
mkdir -p gpurun_out
O=gpurun_out/r02_job3.txt
: > $O
echo "== 16-byte atomicity stress" >> $O
timeout 60 tools/micro/vec16_atomicity 3 >> $O 2>&1
echo "== parity + timing" >> $O
timeout 900 python tools/gpu_res_check.py parity time >> $O 2>&1
echo "== traces (tuning build)" >> $O
export PVC_LIB_PATH=$PWD/planeverb_b200/lib_tune/libplaneverb_b200.so
for cfg in "BigRoom 1024 1000 1 64" "BigRoom 1024 1000 1 65" "BigRoom 1024 1000 1 61" "Shoebox 512 2000 1 60" "Shoebox 512 2000 1 62" "FloorPlanScene 0 0 1 60"; do
  PVC_RES_TRACE=1 timeout 120 python tools/gpu_time_one.py $cfg 2>&1 | tail -2 >> $O
done
for dbg in 1 2; do
  for cfg in "BigRoom 1024 1000 1 64" "Shoebox 512 2000 1 60"; do
    PVC_RES_DEBUG=$dbg timeout 120 python tools/gpu_time_one.py $cfg >> $O 2>&1
  done
done
unset PVC_LIB_PATH
echo "== full GPU suite" >> $O
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 >> $O 2>&1
tail -150 $O
