#!/bin/bash
# The ONE GPU-box runner (replaces the per-experiment scripts of round 1). Usage on the box (through gpurun):
#   bash tools/gpu.sh <step> [<step> ...]      logs go to gpurun_out/<step>.log; every step is bounded by `timeout`
# Steps:
#   tests        pytest -m gpu (whole suite; AY2_PYTEST_ARGS narrows it, e.g. "-k nms")
#   smoke        __graft_entry__.smoke()
#   bench        bench.py at the default settings            bench_ref   bench.py --impl reference (short)
#   parity       tools/parity_report.py (measured errors -> gpurun_out/r02_parity.json)
#   launches     ncu launch list of one benchmarked step     -> gpurun_out/launches.csv
#   convmetrics  per-launch DRAM / L2 / tensor metrics of the conv kernels -> gpurun_out/conv_metrics.csv
#   prof:<regex> one `ncu --set full` capture of the first 3 launches matching <regex> -> gpurun_out/prof_<regex>.ncu-rep
#   py:<script>  python <script> (any tool under tools/)
mkdir -p gpurun_out
N="--profile-from-start off --clock-control none"
for step in "$@"; do
  case "$step" in
    tests)   timeout 1500 python -m pytest tests -m gpu -q --no-header -rf ${AY2_PYTEST_ARGS--x} > gpurun_out/tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/tests.log | cut -c1-300 ;;
    smoke)   timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log ;;
    bench)   timeout 900 python bench.py ${AY2_BENCH_ARGS} > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -1 gpurun_out/bench.log | cut -c1-3000; tail -3 gpurun_out/bench.err ;;
    bench_ref) timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "bench_ref rc=$?"; tail -1 gpurun_out/bench_ref.log | cut -c1-400 ;;
    parity)  timeout 900 python tools/parity_report.py ${AY2_PARITY_ARGS} > gpurun_out/parity.log 2>&1; echo "parity rc=$?"; tail -30 gpurun_out/parity.log ;;
    launches) timeout 600 ncu $N --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches.csv python tools/profile_step.py 3 > gpurun_out/launches.log 2>&1; echo "launches rc=$?" ;;
    convmetrics) timeout 900 ncu $N --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed -k regex:conv_ --csv --log-file gpurun_out/conv_metrics.csv python tools/profile_step.py 3 > gpurun_out/conv_metrics.log 2>&1; echo "convmetrics rc=$?" ;;
    prof:*)  rx="${step#prof:}"; timeout 900 ncu $N --set full --import-source on -k "regex:$rx" -c ${AY2_PROF_COUNT:-3} -o "gpurun_out/prof_$rx" -f python ${AY2_PROF_SCRIPT:-tools/profile_step.py 3} > "gpurun_out/prof_$rx.log" 2>&1; echo "prof $rx rc=$?" ;;
    py:*)    s="${step#py:}"; b=$(basename "${s%% *}" .py); timeout 1200 python $s > "gpurun_out/$b.log" 2>&1; echo "$b rc=$?"; tail -25 "gpurun_out/$b.log" | cut -c1-400 ;;
    *) echo "unknown step $step" ;;
  esac
done
du -sh gpurun_out | tail -1
