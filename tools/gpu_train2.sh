#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_trainstep_gpu.py -m gpu -q --no-header -rf -s > gpurun_out/pytest_train2.log 2>&1; echo "train2 pytest rc=$?"
tail -40 gpurun_out/pytest_train2.log | cut -c1-600
grep -E "rel-L2|cosine|loss " gpurun_out/pytest_train2.log | head -30
# profile of the benchmarked (fused) path: launch list + full capture of all conv launches of one step
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_step.py 2 > gpurun_out/prof_launch.log 2>&1; echo "launch list rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 104 -c 52 -o gpurun_out/prof_conv python tools/profile_step.py 2 > gpurun_out/prof_full.log 2>&1; echo "full rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nms_ -s 0 -c 4 -o gpurun_out/prof_nms python tools/profile_step.py 2 > gpurun_out/prof_nms.log 2>&1; echo "nms rc=$?"
