#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_trainstep_gpu.py -m gpu -q --no-header -rf -s > gpurun_out/pytest_train3.log 2>&1; echo "train3 pytest rc=$?"
grep -E "rel-L2|cosine|loss |passed|failed" gpurun_out/pytest_train3.log | head -40
timeout 900 python tools/bench_train.py --batch 64 --steps 5 --warmup 2 > gpurun_out/bench_train.log 2>&1; echo "bench_train rc=$?"; tail -3 gpurun_out/bench_train.log | cut -c1-900
