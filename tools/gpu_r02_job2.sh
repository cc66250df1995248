#!/bin/bash
# round 2, job 2: new tests, per-pass timeline of the resident kernel (tuning build), ncu capture, first full bench
mkdir -p gpurun_out
O=gpurun_out/r02_job2.txt
: > $O
echo "== new tests" >> $O
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_fullsize.py -x -q --durations=5 >> $O 2>&1
echo "== traces (tuning build)" >> $O
export PVC_LIB_PATH=$PWD/planeverb_b200/lib_tune/libplaneverb_b200.so
for cfg in "BigRoom 1024 1000 1 64" "BigRoom 1024 1000 1 61" "BigRoom 1024 1000 1 62" "Shoebox 512 2000 1 60" "Shoebox 512 2000 1 61" "FloorPlanScene 0 0 1 60" "FloorPlanScene 0 0 1 64"; do
  PVC_RES_TRACE=1 timeout 120 python tools/gpu_time_one.py $cfg >> $O 2>&1
done
echo "== debug switches: 1 = no history stores, 2 = no neighbour wait / halo reload (results invalid)" >> $O
for dbg in 1 2 3; do
  for cfg in "BigRoom 1024 1000 1 64" "Shoebox 512 2000 1 60" "FloorPlanScene 0 0 1 60" "FloorPlanScene 0 0 1 64"; do
    PVC_RES_DEBUG=$dbg timeout 120 python tools/gpu_time_one.py $cfg >> $O 2>&1
  done
done
unset PVC_LIB_PATH
echo "== ncu capture of the resident kernel (1024^2, 1 source, T=400)" >> $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:residentKernel -s 2 -c 1 -o gpurun_out/r02_prof_res64 -f \
    python tools/gpu_time_one.py BigRoom 1024 400 1 64 3 >> $O 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:residentKernel -s 2 -c 1 -o gpurun_out/r02_prof_res60 -f \
    python tools/gpu_time_one.py Shoebox 512 2000 1 60 3 >> $O 2>&1
echo "== bench (default)" >> $O
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_first.json 2> gpurun_out/r02_bench_first.err
tail -c 3000 gpurun_out/r02_bench_first.json >> $O
tail -5 gpurun_out/r02_bench_first.err >> $O
tail -120 $O
