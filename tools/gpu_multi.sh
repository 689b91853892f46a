#!/bin/bash
# N-GPU checks (run with gpurun --gpus N): bench.py under torchrun exactly as the driver launches it, the reference arm,
# and the DDP-style training step benchmark.
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err; echo "bench N=$N rc=$?"; tail -1 gpurun_out/bench_n$N.log | cut -c1-700; tail -3 gpurun_out/bench_n$N.err
timeout 600 $TR --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_ref_n$N.log 2>&1; echo "ref N=$N rc=$?"; tail -1 gpurun_out/bench_ref_n$N.log | cut -c1-300
timeout 900 $TR --master-port 29513 tools/bench_train.py --batch 128 --steps 5 --warmup 2 > gpurun_out/bench_train_n$N.log 2>&1; echo "train N=$N rc=$?"; tail -2 gpurun_out/bench_train_n$N.log | cut -c1-700
