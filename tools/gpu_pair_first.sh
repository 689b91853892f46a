#!/bin/bash
# Gate for a risky kernel change: the CTA-pair parity cases alone under a short timeout; the full suite + bench then run
# with the pair form on (gate passed) or off (AY2_CONV_PAIR=0), so the call yields data either way.
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_conv_gpu.py -k cta_pair -q --no-header -rf -x > gpurun_out/pair_gate.log 2>&1
rc=$?
echo "pair gate rc=$rc"; tail -15 gpurun_out/pair_gate.log | cut -c1-400
if [ $rc -ne 0 ]; then export AY2_CONV_PAIR=0; echo "PAIR DISABLED for the rest of this call"; fi
AY2_PYTEST_ARGS="--timeout=600" bash tools/gpu.sh py:tools/diag_head.py tests bench
