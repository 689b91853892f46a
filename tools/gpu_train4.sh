#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_trainstep_gpu.py -m gpu -q --no-header -rf -s > gpurun_out/pytest_train4.log 2>&1; echo "train4 pytest rc=$?"
grep -E "rel-L2|cosine|loss |passed|failed|Error" gpurun_out/pytest_train4.log | head -40
timeout 300 python tools/diag_train.py mini_v6 > gpurun_out/diag.log 2>&1; head -70 gpurun_out/diag.log | cut -c1-150
