#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err; echo "bench N=$N rc=$?"; tail -1 gpurun_out/bench_n$N.log | cut -c1-700; tail -2 gpurun_out/bench_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_ref_n$N.log 2>&1; echo "ref rc=$?"; tail -1 gpurun_out/bench_ref_n$N.log | cut -c1-400
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/bench_train.py --batch 128 --steps 10 --warmup 3 > gpurun_out/train_n$N.log 2>&1; echo "train rc=$?"; tail -1 gpurun_out/train_n$N.log | cut -c1-500
