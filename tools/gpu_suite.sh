#!/bin/bash
# Full GPU suite + smoke + default bench (with the CPU baseline leg) + reference arm + train bench with a kernel table.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --no-header -rf > gpurun_out/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -5 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
AY2_NMS_TRACE=1 timeout 300 python tools/nms_probe.py > gpurun_out/nms_probe.log 2>&1; echo "probe rc=$?"; grep -i "nms trace" gpurun_out/nms_probe.log | tail -1 | cut -c1-420; grep "nms_sort_scan" gpurun_out/nms_probe.log | head -1 | cut -c1-250
timeout 900 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -1 gpurun_out/bench.log | cut -c1-2500; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "ref rc=$?"; tail -1 gpurun_out/bench_ref.log | cut -c1-600
timeout 900 python tools/bench_train.py --batch 64 --steps 5 --warmup 2 --profile > gpurun_out/bench_train.log 2> gpurun_out/bench_train.err; echo "bench_train rc=$?"; tail -1 gpurun_out/bench_train.log | cut -c1-700; grep -v "^-\|^$" gpurun_out/bench_train.err | head -45 | cut -c1-200
