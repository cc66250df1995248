#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r02_job8.txt
: > $O
python tools/gpu_res_check.py parity 2>&1 | grep -v ": OK " | tail -3 >> $O
echo "== timing (4 mailbox words in flight per row)" >> $O
for cfg in "BigRoom 1024 1000 1 65" "BigRoom 1024 1000 1 64" "BigRoom 1024 1000 1 66" "BigRoom 1024 1000 1 61" "BigRoom 1024 1000 1 70" \
           "BigRoom 1024 1000 4 65" "BigRoom 1024 1000 4 66" "BigRoom 1024 1000 4 47" "FloorPlanScene 1024 1000 4 65" \
           "Shoebox 512 2000 1 60" "Shoebox 512 2000 1 69" "Shoebox 512 2000 1 61" "Shoebox 512 2000 1 70" "Shoebox 512 2000 1 63" "Shoebox 512 2000 4 60" "Shoebox 512 2000 4 63" "Shoebox 512 2000 4 47" \
           "FloorPlanScene 0 0 1 60" "FloorPlanScene 0 0 1 69" "FloorPlanScene 0 0 1 63" "FloorPlanScene 0 0 1 64" "HugeRoom 256 1000 8 63" "HugeRoom 256 1000 8 60" "HugeRoom 768 1000 4 63" "HugeRoom 768 1000 4 47"; do
  timeout 120 python tools/gpu_time_one.py $cfg 2>&1 | tail -1 >> $O
done
echo "== traces" >> $O
export PVC_LIB_PATH=$PWD/planeverb_b200/lib_tune/libplaneverb_b200.so
for cfg in "BigRoom 1024 1000 1 65" "Shoebox 512 2000 1 60" "FloorPlanScene 0 0 1 60"; do
  PVC_RES_TRACE=1 timeout 120 python tools/gpu_time_one.py $cfg 2>&1 | tail -4 | cut -c1-260 >> $O
done
unset PVC_LIB_PATH
cut -c1-400 $O | tail -80
