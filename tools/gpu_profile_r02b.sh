#!/bin/bash
# end-of-round refresh of the profile evidence that the last session's analyzer changes touched: the launch list of the bench command and
# the full capture of encodeResponseKernel (now the 48-register instantiation).  The captures of the two step kernels
# (gpurun_out/r02_prof_res / r02_prof_ws2, tools/gpu_profile_r02.sh) are still current: those kernels did not change.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-verify --no-extras > gpurun_out/r02_bench_under_ncu.json 2> gpurun_out/r02_ncu_launches.err
ncu --set full --clock-control none --import-source on -k regex:encodeResponseKernel -s 3 -c 1 -o gpurun_out/r02_prof_encode -f \
    python bench.py --steps 1 --warmup 3 --T 400 --no-cpu-baseline --no-verify --no-extras > /dev/null 2> gpurun_out/r02_ncu_encode.err
ncu --set full --clock-control none --import-source on -k regex:encodeResponseKernel -s 1 -c 1 -o gpurun_out/r02_prof_encode_huge -f \
    python tools/gpu_time_one.py HugeRoom 2048 4000 2 0 2 > /dev/null 2> gpurun_out/r02_ncu_encode_huge.err
ls -la gpurun_out | grep -E "r02_prof_encode|r02_launches"
