#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_nms_gpu.py tests/test_model_gpu.py tests/test_conv_gpu.py -m gpu -q --no-header -rf -x > gpurun_out/pytest_quick.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/pytest_quick.log
AY2_NMS_TRACE=1 timeout 300 python tools/nms_probe.py > gpurun_out/nms_probe.log 2>&1; echo "probe rc=$?"; grep -v "^-\|^$" gpurun_out/nms_probe.log | grep -i "nms\|dets\|conv_tc\|Self CUDA time" | cut -c1-260 | head -40
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_quick.log 2> gpurun_out/bench_quick.err; echo "bench rc=$?"; tail -1 gpurun_out/bench_quick.log | cut -c1-1800; tail -3 gpurun_out/bench_quick.err
