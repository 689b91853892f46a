#!/bin/bash
mkdir -p gpurun_out
AY2_NMS_TRACE=1 timeout 300 python tools/nms_probe.py > gpurun_out/nms_probe.log 2>&1; echo "probe rc=$?"; grep -i "nms trace" gpurun_out/nms_probe.log | tail -3 | cut -c1-400; grep "nms_sort_scan" gpurun_out/nms_probe.log | head -2 | cut -c1-250
