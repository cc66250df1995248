#!/bin/bash
# analyzer timing after the 5-FMA Horner logf (was 6 DP ops): all-onset scenes + the latency pair; then the parity suites that cover RT60
set -x
python tools/gpu_time_one.py HugeRoom 2048 4000 2 0 3
python tools/gpu_time_one.py FloorPlanScene 1024 4000 4 0 3
python tools/gpu_time_one.py Shoebox 512 2000 1 0 6
python tools/gpu_time_one.py FloorPlanScene 0 0 1 0 6
python tools/gpu_time_one.py BigRoom 1024 4000 4 0 3
python -m pytest tests/test_gpu_parity.py tests/test_gpu_streamed.py -x -q 2>&1 | tail -4
