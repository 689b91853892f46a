#!/bin/bash
# final validation of the round: full GPU suite, smoke, default bench, reference arm
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --no-header -rf > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 800 python bench.py > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err; echo "bench rc=$?"; tail -1 gpurun_out/bench_default.log | cut -c1-2200
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-300
