#!/bin/bash
# Re-establish the measured state on a fresh box: GPU tests (with durations), smoke, both bench arms, train bench.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --no-header -rf --durations=15 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -25 gpurun_out/pytest_gpu.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -1 gpurun_out/bench.log; tail -3 gpurun_out/bench.err
timeout 300 python tools/bench_train.py --batch 128 --steps 10 --warmup 3 > gpurun_out/bench_train.log 2>&1; echo "train rc=$?"; tail -1 gpurun_out/bench_train.log
