#!/bin/bash
# N-GPU bench exactly as the driver launches it (torchrun, one rank per GPU). Usage: bash tools/gpu_multi.sh N [bench args]
N=${1:-2}; shift
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 30 --warmup 5 "$@" > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err
echo "bench N=$N rc=$?"; tail -1 gpurun_out/bench_n$N.log | cut -c1-6000; tail -5 gpurun_out/bench_n$N.err | cut -c1-500
