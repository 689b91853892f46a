#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_chain_gpu.py tests/test_model_gpu.py -m gpu -q --no-header -rf -x > gpurun_out/pytest_chain.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_chain.log | cut -c1-250
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -1 gpurun_out/bench.log | cut -c1-400; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/conv_layers.json'))
for r in d:
    if 'chain' in r: print(r['chain'], r['hw'], round(r['ms'],3))
print('chain ms', sum(r['ms'] for r in d if 'chain' in r), 'total', sum(r['ms'] for r in d))
PY
