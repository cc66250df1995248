#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --cache-control none --import-source on -k regex:fusedStep -s 350 -c 1 -o gpurun_out/prof_fused_v22_warm -f \
    python bench.py --steps 1 --warmup 3 --T 400 --no-cpu-baseline > /dev/null 2> gpurun_out/ncu_fused_v22.err
