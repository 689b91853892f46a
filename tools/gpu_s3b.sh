#!/bin/bash
# session 3, call B: conv tests + timelines + bench after the epilogue restructure
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_model_gpu.py tests/test_nms_gpu.py -m gpu -q --no-header -rf -x > gpurun_out/pytest_conv.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_conv.log | cut -c1-300
for a in "128 128 1 1 40" "256 256 1 1 20" "256 256 1 1 40" "64 64 1 1 80"; do
  timeout 120 python tools/conv_timeline.py $a 2>&1 | tail -16
done > gpurun_out/conv_timeline4.log 2>&1
grep -A9 "^conv" gpurun_out/conv_timeline4.log | grep "^conv\|first acc\|first store\|roles"
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
l=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print('value',round(l['value']),'ms/step',round(l['ms_per_step'],3),'e2e',round(l['e2e']['value']),'roof',round(l['roofline']['frac'],3),'conv_ms',round(l['roofline']['conv_ms_per_step'],3))
PY
