#!/bin/bash
# A/B a conv-kernel change: conv + model parity tests, then the bench with and without the env knob given as $1 (e.g. AY2_CONV_PDL=0).
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_model_gpu.py -m gpu -q --no-header -rf -x > gpurun_out/pytest_quick.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_quick.log
for rep in 1 2; do
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_a.log 2> gpurun_out/bench_a.err; echo "A rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_a.log').read().strip().splitlines()[-1]); print('A', d['value'], d['ms_per_step'], d['roofline']['conv_ms_per_step'], d['e2e']['value'])
PY
env $1 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_b.log 2> gpurun_out/bench_b.err; echo "B($1) rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_b.log').read().strip().splitlines()[-1]); print('B', d['value'], d['ms_per_step'], d['roofline']['conv_ms_per_step'], d['e2e']['value'])
PY
done
tail -3 gpurun_out/bench_a.err
