#!/bin/bash
# Quick GPU iteration: conv/model/NMS parity + bench without the CPU legs.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_model_gpu.py tests/test_nms_gpu.py tests/test_pointwise_gpu.py -m gpu -q --no-header -rf -x > gpurun_out/pytest_quick.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/pytest_quick.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_quick.log 2> gpurun_out/bench_quick.err; echo "bench rc=$?"; tail -1 gpurun_out/bench_quick.log | cut -c1-1600; tail -3 gpurun_out/bench_quick.err
