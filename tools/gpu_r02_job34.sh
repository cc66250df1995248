#!/bin/bash
# L2-resident bands in the generational kernel's item order (grids whose state exceeds the L2): sweep of band height and chunk length on
# config 4's grid (HugeRoom 2048^2, two sources, T = 1000) in a -DPVC_TUNING build; PVC_BAND_ROWS=0 = the plain order
T=planeverb_b200/lib_tune/libplaneverb_b200.so
for cfg in "0 16" "15 8" "15 4" "22 8" "22 4" "11 4" "11 8" "15 16" "8 4"; do
  set -- $cfg
  echo "band=$1 gens=$2: $(PVC_LIB_PATH=$PWD/$T PVC_BAND_ROWS=$1 PVC_GROUP_GENS=$2 python tools/gpu_time_one.py HugeRoom 2048 1000 2 47 3 | cut -c1-90)"
done
echo "release library (default policy):"
python tools/gpu_time_one.py HugeRoom 2048 1000 2 0 3 | cut -c1-120
python tools/gpu_time_one.py HugeRoom 2048 1000 1 0 3 | cut -c1-120
python tools/gpu_time_one.py FloorPlanScene 1536 1000 2 0 3 | cut -c1-120
python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_streamed.py -x -q 2>&1 | tail -3
