#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pointwise_gpu.py tests/test_model_gpu.py -m gpu -q --no-header -rf -x > gpurun_out/pytest_part.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_part.log | cut -c1-300
AY2_NMS_TRACE=1 AY2_NMS_TRACE_ONLY=1 timeout 300 python tools/nms_probe.py 2>&1 | grep -i "trace" | tail -2 | cut -c1-600
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
l=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print('value',round(l['value']),'ms/step',round(l['ms_per_step'],3),'e2e',round(l['e2e']['value']),'roof',round(l['roofline']['frac'],3),'conv_ms',round(l['roofline']['conv_ms_per_step'],3),'nms_ms',round(l['roofline']['nms_ms_per_step'],4))
PY
