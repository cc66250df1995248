#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r02_job9.txt
: > $O
export PVC_LIB_PATH=$PWD/planeverb_b200/lib_tune/libplaneverb_b200.so
for cfg in "BigRoom 1024 1000 1 65" "BigRoom 1024 1000 1 67" "Shoebox 512 2000 1 60" "none 512 2000 1 60"; do
  PVC_RES_TRACE=1 timeout 120 python tools/gpu_time_one.py $cfg 2>&1 | tail -5 | cut -c1-330 >> $O
done
unset PVC_LIB_PATH
cat $O
