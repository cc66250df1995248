#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r02_job6.txt
: > $O
echo "== hybrid barrier variants 67/68 vs 65/64, 66" >> $O
for cfg in "BigRoom 1024 1000 1 65" "BigRoom 1024 1000 1 67" "BigRoom 1024 1000 1 64" "BigRoom 1024 1000 1 68" "BigRoom 1024 1000 4 67" "FloorPlanScene 1024 1000 4 67"; do
  timeout 120 python tools/gpu_time_one.py $cfg 2>&1 | tail -1 >> $O
done
export PVC_LIB_PATH=$PWD/planeverb_b200/lib_tune/libplaneverb_b200.so
echo "== traces: hand-over latency; with and without nanosleep (dbg=4)" >> $O
for dbg in 0 4; do
for cfg in "BigRoom 1024 1000 1 65" "BigRoom 1024 1000 1 67" "Shoebox 512 2000 1 60" "FloorPlanScene 0 0 1 60"; do
  PVC_RES_DEBUG=$dbg PVC_RES_TRACE=1 timeout 120 python tools/gpu_time_one.py $cfg 2>&1 | tail -4 | cut -c1-300 >> $O
done
done
unset PVC_LIB_PATH
python tools/gpu_res_check.py parity 2>&1 | tail -1 >> $O
cut -c1-400 $O | tail -80
