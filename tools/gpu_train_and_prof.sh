#!/bin/bash
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_train_gpu.py -m gpu -q --no-header -rf > gpurun_out/pytest_train.log 2>&1; echo "train pytest rc=$?"
tail -25 gpurun_out/pytest_train.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_step.py 2 > gpurun_out/prof_launch.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 104 -c 10 -o gpurun_out/prof_conv python tools/profile_step.py 2 > gpurun_out/prof_full.log 2>&1; echo "full rc=$?"
tail -2 gpurun_out/prof_full.log
