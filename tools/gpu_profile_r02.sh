#!/bin/bash
# round 2 measurement pass on one B200: the bench line, the ncu launch list of the same command, full ncu captures of the two step
# kernels and of the analyzer kernel.  Outputs under gpurun_out/ (tools/make_profile_summary_r02.py turns them into profiles/r02_*).
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err
tail -c 1500 gpurun_out/r02_bench_default.json; tail -3 gpurun_out/r02_bench_default.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err
# launch list of the same command (every kernel of the process)
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-verify --no-extras > gpurun_out/r02_bench_under_ncu.json 2> gpurun_out/r02_ncu_launches.err
# full captures: the resident step kernel of the bench workload (--T 400: one launch = one source, 100 passes; the 13th launch is
# the first of the timed solve), the generational kernel on config 4's grid (2048^2, two sources), the analyzer
ncu --set full --clock-control none --import-source on -k regex:residentKernel -s 12 -c 1 -o gpurun_out/r02_prof_res -f \
    python bench.py --steps 1 --warmup 3 --T 400 --no-cpu-baseline --no-verify --no-extras > /dev/null 2> gpurun_out/r02_ncu_res.err
ncu --set full --clock-control none --import-source on -k regex:stepKernel -s 2 -c 1 -o gpurun_out/r02_prof_ws2 -f \
    python tools/gpu_time_one.py HugeRoom 2048 400 2 0 3 > /dev/null 2> gpurun_out/r02_ncu_ws2.err
ncu --set full --clock-control none --import-source on -k regex:encodeResponseKernel -s 3 -c 1 -o gpurun_out/r02_prof_encode -f \
    python bench.py --steps 1 --warmup 3 --T 400 --no-cpu-baseline --no-verify --no-extras > /dev/null 2> gpurun_out/r02_ncu_encode.err
# streamed solver: config 4's eight 2048^2 sources as one batch (history 800 samples, 5 chunks): timing and the launch list of one solve
python tools/gpu_time_one.py HugeRoom 2048 4000 8 0 3 800 > gpurun_out/r02_streamed_time.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_streamed_launches.csv \
    python tools/gpu_time_one.py HugeRoom 2048 4000 8 0 2 800 > /dev/null 2> gpurun_out/r02_ncu_streamed.err
ls -la gpurun_out | grep r02_ | tail -14
