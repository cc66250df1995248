#!/bin/bash
python tools/gpu_small_case.py 7 SmallRoom; python tools/gpu_small_case.py 7 none; python tools/gpu_small_case.py 5 none; python tools/gpu_small_case.py 0 none
ncu --set full --clock-control none --cache-control none --import-source on -k regex:fusedStep -s 250 -c 1 -o gpurun_out/prof_small_v7 -f python tools/gpu_small_case.py 7 SmallRoom > /dev/null 2>&1
ncu --set full --clock-control none --cache-control none --import-source on -k regex:fusedStep -s 250 -c 1 -o gpurun_out/prof_small_v7_empty -f python tools/gpu_small_case.py 7 none > /dev/null 2>&1
