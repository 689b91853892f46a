#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_chain_gpu.py tests/test_model_gpu.py -m gpu -q --no-header -rf -x > gpurun_out/pytest_chain.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_chain.log | cut -c1-250
for a in "32 160" "64 80" "128 40"; do timeout 120 python tools/chain_timeline.py $a 2>&1 | grep "kernel ms"; done
for mc in 32 128; do
AY2_FUSE_BOTTLENECK_MAX_C=$mc timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$mc.log 2> gpurun_out/bench.err; python - <<PY
import json
l=json.loads(open('gpurun_out/bench_$mc.log').read().strip().splitlines()[-1])
print('maxc $mc value',round(l['value']),'ms/step',round(l['ms_per_step'],3),'e2e',round(l['e2e']['value']),'roof',round(l['roofline']['frac'],3),'conv_ms',round(l['roofline']['conv_ms_per_step'],3), l['roofline']['kernel'][:60])
PY
done
timeout 600 python tools/bench_tucker.py 0.5 2>&1 | tail -3
