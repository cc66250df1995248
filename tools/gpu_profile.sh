#!/bin/bash
# Round-1 GPU measurement pass: bench line, ncu launch list of the same command, full ncu capture of the
# two dominant kernels. Outputs under gpurun_out/.
set -x
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
tail -c 3000 gpurun_out/bench_default.json
for v in 1 2 4; do python bench.py --steps 2 --warmup 3 --variant $v --no-cpu-baseline > gpurun_out/bench_variant$v.json 2>> gpurun_out/bench_default.err; done
python bench.py --steps 2 --warmup 3 --step-kernel 1 --no-cpu-baseline > gpurun_out/bench_baseline_kernel.json 2>> gpurun_out/bench_default.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 3020 -c 1100 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.json 2> gpurun_out/ncu_launches.err
ncu --set full --clock-control none --import-source on -k regex:fusedStepKernel -s 40 -c 2 -o gpurun_out/prof_fused -f \
    python bench.py --steps 1 --warmup 3 --T 400 --no-cpu-baseline > /dev/null 2> gpurun_out/ncu_fused.err
ncu --set full --clock-control none --import-source on -k regex:encodeResponseKernel -s 1 -c 1 -o gpurun_out/prof_encode -f \
    python bench.py --steps 1 --warmup 3 --T 400 --no-cpu-baseline > /dev/null 2> gpurun_out/ncu_encode.err
ncu --set full --clock-control none --import-source on -k regex:listenerDirectionKernel -s 1 -c 1 -o gpurun_out/prof_walk -f \
    python bench.py --steps 1 --warmup 3 --T 400 --no-cpu-baseline > /dev/null 2> gpurun_out/ncu_walk.err
ls -la gpurun_out
