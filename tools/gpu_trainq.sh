#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_gpu.py tests/test_trainstep_gpu.py tests/test_loss_gpu.py -m gpu -q --no-header -rf -x > gpurun_out/pytest_train.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_train.log | cut -c1-300
timeout 900 python tools/bench_train.py --batch 64 --steps 5 --warmup 2 --profile > gpurun_out/bench_train.log 2> gpurun_out/bench_train.err; echo "bench_train rc=$?"; tail -1 gpurun_out/bench_train.log | cut -c1-400; grep -v "^-\|^$" gpurun_out/bench_train.err | sed -n 4,34p | cut -c1-100,150-215
