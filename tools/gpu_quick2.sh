#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_chain_gpu.py tests/test_model_gpu.py tests/test_train_gpu.py -m gpu -q --no-header -rf -x > gpurun_out/pytest_q.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_q.log | cut -c1-250
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; python - <<'PY'
import json
l=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print('value',round(l['value']),'ms/step',round(l['ms_per_step'],3),'e2e',round(l['e2e']['value']),'roof',round(l['roofline']['frac'],3),'conv_ms',round(l['roofline']['conv_ms_per_step'],3))
d=json.load(open('gpurun_out/conv_layers.json'))
for r in d:
    if 'chain' in r: print(r['chain'], r['hw'], round(r['ms'],3))
print('chain ms', sum(r['ms'] for r in d if 'chain' in r), 'total', sum(r['ms'] for r in d))
PY
tail -3 gpurun_out/bench.err
