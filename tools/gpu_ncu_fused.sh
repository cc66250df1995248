#!/bin/bash
mkdir -p gpurun_out
for v in 7 0; do
ncu --set full --clock-control none --cache-control none --import-source on -k regex:fusedStep -s 60 -c 1 -o gpurun_out/prof_fused_v${v}_warm -f \
    python bench.py --steps 1 --warmup 3 --T 400 --no-cpu-baseline --variant $v > /dev/null 2> gpurun_out/ncu_fused_v$v.err
done
