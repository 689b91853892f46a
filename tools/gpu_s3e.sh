#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_nms_gpu.py tests/test_loss_gpu.py -m gpu -q --no-header -rf > gpurun_out/pytest_nmsloss.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_nmsloss.log | cut -c1-300
