#!/bin/bash
# round 2 multi-GPU pass on ONE box with G GPUs (gpurun --gpus G): the multi-device C-ABI tests, then the bench line at every
# N <= G (with the config 4 / config 5 extras), launched the way the driver launches it.  Outputs under gpurun_out/.
G=${1:-8}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r02_pytest_multi_g$G.txt
timeout 600 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
tail -c 900 gpurun_out/r02_bench_n1.json
P=29500
for N in 2 4 8; do
  [ $N -le $G ] || continue
  P=$((P+1))
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P \
      bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
  echo "N=$N rc=$?"; tail -c 1200 gpurun_out/r02_bench_n$N.json; tail -3 gpurun_out/r02_bench_n$N.err
done
