#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --no-header -rf -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log | cut -c1-300
for h in 1 0; do
AY2_CONV_HALO=$h timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_halo$h.log 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<PY
import json
l=json.loads(open('gpurun_out/bench_halo$h.log').read().strip().splitlines()[-1])
print('halo $h value',round(l['value']),'ms/step',round(l['ms_per_step'],3),'e2e',round(l['e2e']['value']),'roof',round(l['roofline']['frac'],3),'conv_ms',round(l['roofline']['conv_ms_per_step'],3))
PY
cp gpurun_out/conv_layers.json gpurun_out/conv_layers_halo$h.json
done
