#!/bin/bash
# round 2, late: compute-sanitizer over the kernels added after tools/gpu_sanitize_r02.sh ran -- the streamed solver's chunk kernels
# (forwardChunkKernel, backwardChunkKernel, initCarryKernel, chunked launches of the generational kernel), walkSmallKernel (links in
# shared memory) and the analyzer's batched remainders.  racecheck is not run on walkSmallKernel: its in-place pointer jumping reads a
# link while another thread shortens it BY DESIGN (either value lies on the same walk), like walkJumpKernel does in global memory.
mkdir -p gpurun_out
O=gpurun_out/r02_sanitizer_b.txt
: > $O
run() { echo "== $1: $2" >> $O; shift 2; timeout 900 compute-sanitizer "$@" 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Error|Invalid|Uninitialized|hazard|^[0-9]+ " | head -8 >> $O; }
run "memcheck" "streamed, FloorPlanScene 70x70, T=435, history 104 (K=5), variant auto" --tool memcheck --print-limit 5 python tools/gpu_small_case.py 0 FloorPlanScene 0 1 0 104
run "memcheck" "streamed, FloorPlanScene 300x300, 2 sources, T=203, history 64 (K=4)" --tool memcheck --print-limit 5 python tools/gpu_small_case.py 0 FloorPlanScene 300 2 203 64
run "initcheck" "streamed, FloorPlanScene 70x70, T=435, history 104" --tool initcheck --print-limit 5 python tools/gpu_small_case.py 0 FloorPlanScene 0 1 0 104
run "memcheck" "full history, FloorPlanScene 127x127 (walkSmallKernel with 64 KB of links, batched remainders)" --tool memcheck --print-limit 5 python tools/gpu_small_case.py 0 FloorPlanScene 127 1 0
run "initcheck" "full history, FloorPlanScene 70x70 (walkSmallKernel)" --tool initcheck --print-limit 5 python tools/gpu_small_case.py 0 FloorPlanScene
cat $O
