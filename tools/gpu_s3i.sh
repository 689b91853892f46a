#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_gpu.py tests/test_trainstep_gpu.py -m gpu -q --no-header -rf -x > gpurun_out/pytest_train.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_train.log | cut -c1-300
for b in 128 16; do
timeout 500 python tools/bench_train.py --batch $b --steps 10 --warmup 3 --profile > gpurun_out/train_b$b.log 2> gpurun_out/train_b$b.err; echo "train b$b rc=$?"; tail -1 gpurun_out/train_b$b.log | cut -c1-260
grep "aten::clone\|aten::copy_\|FillFunctor\|sgd_ema\|Memcpy DtoD" gpurun_out/train_b$b.err | cut -c1-60,150-215
tail -3 gpurun_out/train_b$b.err | cut -c1-200
done
