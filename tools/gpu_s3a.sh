#!/bin/bash
# session 3, call A: GPU tests on HEAD + conv timelines
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --no-header -rf -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log | cut -c1-300
for a in "128 128 1 1 40" "256 256 1 1 20" "64 64 1 1 80" "128 128 3 1 40" "256 256 1 1 40" "512 256 1 1 20" "64 64 1 1 160"; do
  timeout 120 python tools/conv_timeline.py $a 2>&1 | tail -11
done > gpurun_out/conv_timeline.log 2>&1
cat gpurun_out/conv_timeline.log
