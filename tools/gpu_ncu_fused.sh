#!/bin/bash
mkdir -p gpurun_out
for v in 0 1; do
ncu --set full --clock-control none --import-source on -k regex:fusedStepKernel -s 60 -c 1 -o gpurun_out/prof_fused_v$v -f \
    python bench.py --steps 1 --warmup 3 --T 400 --no-cpu-baseline --variant $v > /dev/null 2> gpurun_out/ncu_fused_v$v.err
done
ncu --set full --clock-control none --import-source on -k regex:encodeResponseKernel -s 1 -c 1 -o gpurun_out/prof_encode2 -f \
    python bench.py --steps 1 --warmup 3 --T 400 --no-cpu-baseline > /dev/null 2> gpurun_out/ncu_encode2.err
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2.json
cat gpurun_out/bench_r2.json
