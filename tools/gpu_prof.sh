#!/bin/bash
mkdir -p gpurun_out
# launch list of 2 eager steps (skip the calibration forward: 61 launches at bs 4 + setup)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_step.py 2 > gpurun_out/prof_launch.log 2>&1; echo "launch list rc=$?"
tail -2 gpurun_out/prof_launch.log
# full capture: first 6 conv launches of the second bs64 step (calibration step = 55 conv launches; step = 55)
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 110 -c 8 -o gpurun_out/prof_conv python tools/profile_step.py 2 > gpurun_out/prof_full.log 2>&1; echo "full rc=$?"
tail -2 gpurun_out/prof_full.log
ls -la gpurun_out/
