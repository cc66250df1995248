#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r02_job22.txt
: > $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 >> $O
for cfg in "BigRoom 1024 4000 4 0" "FloorPlanScene 1024 1000 4 0" "HugeRoom 768 1000 4 0" "Shoebox 512 2000 4 0" "Shoebox 512 2000 1 0" "HugeRoom 256 1000 8 0" "FloorPlanScene 0 0 1 0" "HugeRoom 2048 400 2 0" "BigRoom 1024 1000 1 0" "BigRoom 1024 1000 1 64" "HugeRoom 768 1000 1 0" "HugeRoom 768 1000 1 60" "HugeRoom 768 1000 1 62" "BigRoom 900 1000 4 0" "BigRoom 900 1000 4 65" "BigRoom 900 1000 4 47"; do
  timeout 200 python tools/gpu_time_one.py $cfg 2>&1 | tail -1 >> $O
done
export PVC_LIB_PATH=$PWD/planeverb_b200/lib_tune/libplaneverb_b200.so
PVC_RES_TRACE=1 timeout 120 python tools/gpu_time_one.py BigRoom 1024 1000 1 65 2>&1 | tail -6 | cut -c1-330 >> $O
unset PVC_LIB_PATH
python bench.py --steps 3 --warmup 3 --no-extras > gpurun_out/r02_bench_flow.json 2> gpurun_out/r02_bench_flow.err
tail -c 2500 gpurun_out/r02_bench_flow.json >> $O
cat $O
