#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r02_job12.txt
: > $O
python tools/gpu_res_check.py parity 2>&1 | grep -v ": OK " | tail -3 >> $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3 >> $O
echo "== early record (resident + generational)" >> $O
for cfg in "BigRoom 1024 1000 1 65" "BigRoom 1024 1000 1 61" "BigRoom 1024 1000 4 65" "BigRoom 1024 1000 4 47" "FloorPlanScene 1024 1000 4 65" "FloorPlanScene 1024 1000 4 47" \
           "Shoebox 512 2000 1 60" "Shoebox 512 2000 4 60" "Shoebox 512 2000 4 47" "FloorPlanScene 0 0 1 60" "HugeRoom 256 1000 8 60" "HugeRoom 768 1000 4 47" "HugeRoom 2048 400 2 47"; do
  timeout 200 python tools/gpu_time_one.py $cfg 2>&1 | tail -1 >> $O
done
export PVC_LIB_PATH=$PWD/planeverb_b200/lib_tune/libplaneverb_b200.so
for cfg in "BigRoom 1024 1000 1 65" "Shoebox 512 2000 1 60"; do
  PVC_RES_TRACE=1 timeout 120 python tools/gpu_time_one.py $cfg 2>&1 | tail -5 | cut -c1-330 >> $O
done
unset PVC_LIB_PATH
cat $O
