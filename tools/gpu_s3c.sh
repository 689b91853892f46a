#!/bin/bash
mkdir -p gpurun_out
for e in 0 2; do
  echo "== AY2_CONV_EXPERIMENT=$e"
  AY2_CONV_EXPERIMENT=$e timeout 300 python -m pytest tests/test_conv_gpu.py -m gpu -q --no-header -rf 2>&1 | tail -12 | cut -c1-260
done
echo "== forced halo (AY2_CONV_HALO=2)"
AY2_CONV_HALO=2 timeout 300 python -m pytest tests/test_conv_gpu.py -m gpu -q --no-header -rf 2>&1 | tail -8 | cut -c1-260
for a in "128 128 3 1 40" "64 64 3 1 80"; do
  for h in 0 1; do AY2_CONV_HALO=$h timeout 120 python tools/conv_timeline.py $a 2>&1 | tail -16 | grep "^conv\|first acc\|roles"; done
done
