#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_chain_gpu.py tests/test_model_gpu.py -m gpu -q --no-header -rf -x > gpurun_out/pytest_chain.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_chain.log | cut -c1-300
timeout 120 python tools/chain_timeline.py 32 160 2>&1 | grep "kernel ms\|tile 3\|tile 4"
for sp in 1 0; do
AY2_CHAIN_SPLIT=$sp timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_split$sp.log 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<PY
import json
l=json.loads(open('gpurun_out/bench_split$sp.log').read().strip().splitlines()[-1])
L=json.load(open('gpurun_out/conv_layers.json'))
print('split $sp value',round(l['value']),'ms/step',round(l['ms_per_step'],3),'chain ms',[round(x['ms'],4) for x in L if 'chain' in x])
PY
done
