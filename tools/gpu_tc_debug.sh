#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py -m gpu -q --maxfail=60 --tb=line > gpurun_out/tc_conv.log 2>&1; echo "conv exit $?"
grep -E "^/|^E |passed|failed|Error|timeout|ni_b200" gpurun_out/tc_conv.log | cut -c1-260 | head -60
