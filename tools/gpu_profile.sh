#!/bin/bash
# GPU measurement pass: bench line, ncu launch list of the same command, full ncu capture of the dominant
# kernels, per-CTA timeline. Outputs under gpurun_out/ (tools/make_profile_summary.py copies summaries into profiles/).
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
tail -c 2500 gpurun_out/bench_default.json
# launch list of the same command (every kernel of the process; the summary keeps the last two solves: per solve 4 step-kernel
# launches of <= 256 generations x 4 steps + the analyzer's encode / walk-link / walk-jump / walk-resolve kernels)
PVC_NO_GRAPHS=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.json 2> gpurun_out/ncu_launches.err
PVC_NO_GRAPHS=1 ncu --set full --clock-control none --import-source on -k regex:stepKernel -s 4 -c 1 -o gpurun_out/prof_fused -f \
    python bench.py --steps 1 --warmup 3 --T 400 --no-cpu-baseline > /dev/null 2> gpurun_out/ncu_fused.err
PVC_NO_GRAPHS=1 ncu --set full --clock-control none --import-source on -k regex:encodeResponseKernel -s 3 -c 1 -o gpurun_out/prof_encode -f \
    python bench.py --steps 1 --warmup 3 --T 400 --no-cpu-baseline > /dev/null 2> gpurun_out/ncu_encode.err
PVC_NO_GRAPHS=1 ncu --set full --clock-control none --import-source on -k regex:walkJumpKernel -s 16 -c 1 -o gpurun_out/prof_walk -f \
    python bench.py --steps 1 --warmup 3 --T 400 --no-cpu-baseline > /dev/null 2> gpurun_out/ncu_walk.err
ls -la gpurun_out | tail -12
