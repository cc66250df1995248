#!/bin/bash
# round 2, job 5: live padding row fix, single-word polling, 5-row-warp variant
mkdir -p gpurun_out
O=gpurun_out/r02_job5.txt
: > $O
echo "== edge-case diagnosis" >> $O
timeout 600 python tools/gpu_diag_edge.py >> $O 2>&1
echo "== parity" >> $O
timeout 900 python tools/gpu_res_check.py parity 2>&1 | grep -v ": OK " >> $O
echo "== timing" >> $O
for cfg in "BigRoom 1024 1000 1 64" "BigRoom 1024 1000 1 65" "BigRoom 1024 1000 1 66" "BigRoom 1024 1000 1 61" \
           "BigRoom 1024 1000 4 65" "BigRoom 1024 1000 4 66" "BigRoom 1024 1000 4 47" "FloorPlanScene 1024 1000 4 65" "FloorPlanScene 1024 1000 4 66" \
           "Shoebox 512 2000 1 60" "Shoebox 512 2000 1 61" "Shoebox 512 2000 1 62" "Shoebox 512 2000 1 63" "Shoebox 512 2000 4 60"  "Shoebox 512 2000 4 63" "Shoebox 512 2000 4 47" \
           "FloorPlanScene 0 0 1 60" "FloorPlanScene 0 0 1 63" "FloorPlanScene 0 0 1 64" "FloorPlanScene 0 0 1 66"; do
  timeout 120 python tools/gpu_time_one.py $cfg 2>&1 | tail -1 >> $O
done
echo "== traces (tuning build)" >> $O
export PVC_LIB_PATH=$PWD/planeverb_b200/lib_tune/libplaneverb_b200.so
for cfg in "BigRoom 1024 1000 1 65" "BigRoom 1024 1000 1 66" "Shoebox 512 2000 1 60" "FloorPlanScene 0 0 1 60"; do
  PVC_RES_TRACE=1 timeout 120 python tools/gpu_time_one.py $cfg 2>&1 | tail -3 | cut -c1-330 >> $O
done
for cfg in "BigRoom 1024 1000 1 65" "BigRoom 1024 1000 1 66"; do
  PVC_RES_DEBUG=2 timeout 120 python tools/gpu_time_one.py $cfg >> $O 2>&1
done
unset PVC_LIB_PATH
echo "== ncu: resident kernel 1024^2 (variant 65 and 66)" >> $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:residentKernel -s 2 -c 1 -o gpurun_out/r02_prof_res65 -f \
    python tools/gpu_time_one.py BigRoom 1024 400 1 65 3 >> /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:residentKernel -s 2 -c 1 -o gpurun_out/r02_prof_res66 -f \
    python tools/gpu_time_one.py BigRoom 1024 400 1 66 3 >> /dev/null 2>&1
cut -c1-400 $O | tail -150
