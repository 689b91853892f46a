#!/bin/bash
# Profiles of the benchmarked (fused) step. Keeps gpurun_out small (< 64 MiB): metric CSVs + one small full capture.
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_step.py 2 > gpurun_out/prof_launch.log 2>&1; echo "launch list rc=$?"; tail -1 gpurun_out/prof_launch.log
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:conv_tc -s 104 -c 52 --csv --log-file gpurun_out/conv52.csv python tools/profile_step.py 2 > gpurun_out/prof_conv52.log 2>&1; echo "conv52 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nms_ -s 0 -c 2 -o gpurun_out/prof_nms python tools/profile_step.py 1 > gpurun_out/prof_nms.log 2>&1; echo "nms rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 104 -c 3 -o gpurun_out/prof_conv3 python tools/profile_step.py 2 > gpurun_out/prof_conv3.log 2>&1; echo "conv3 rc=$?"
du -sh gpurun_out
