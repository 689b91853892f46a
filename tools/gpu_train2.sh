#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_train_gpu.py tests/test_trainstep_gpu.py tests/test_conv_gpu.py -m gpu -q --no-header -rf > gpurun_out/pytest_train2.log 2>&1; echo "train2 pytest rc=$?"
tail -40 gpurun_out/pytest_train2.log | cut -c1-600
