#!/bin/bash
# Profiles of ONE benchmarked step (eager replay of the graph's launch sequence, cudaProfilerStart/Stop around the last step).
# Keeps gpurun_out small (< 64 MiB): metric CSVs + small full captures.
mkdir -p gpurun_out
N="--profile-from-start off --clock-control none"
timeout 600 ncu $N --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches.csv python tools/profile_step.py 3 > gpurun_out/prof_launch.log 2>&1; echo "launch list rc=$?"; tail -1 gpurun_out/prof_launch.log
timeout 900 ncu $N --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed -k regex:conv_ --csv --log-file gpurun_out/conv52.csv python tools/profile_step.py 3 > gpurun_out/prof_conv52.log 2>&1; echo "conv metrics rc=$?"
timeout 600 ncu $N --set full --import-source on -k regex:nms_ -o gpurun_out/prof_nms -f python tools/profile_step.py 3 > gpurun_out/prof_nms.log 2>&1; echo "nms rc=$?"
# full captures: launch 13 (256->256 1x1 40x40, N tile 256), 15 (128->128 3x3 40x40), 21 (256->512 3x3 s2), the chain kernel
timeout 600 ncu $N --set full --import-source on -k regex:conv_tc -s 12 -c 4 -o gpurun_out/prof_conv3 -f python tools/profile_step.py 3 > gpurun_out/prof_conv3.log 2>&1; echo "conv3 rc=$?"
timeout 600 ncu $N --set full --import-source on -k regex:conv_chain -c 1 -o gpurun_out/prof_chain -f python tools/profile_step.py 3 > gpurun_out/prof_chain.log 2>&1; echo "chain rc=$?"
du -sh gpurun_out
