#!/bin/bash
# opt-in deferred join of the side-stream filter gradients (NI_WGRAD_DEFER=1): test-suite with it on, then A/B bench lines
cd /root/repo
mkdir -p gpurun_out
( NI_WGRAD_DEFER=1 timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 ) > gpurun_out/defer_tests.log; tail -2 gpurun_out/defer_tests.log
for b in 32 256; do
  for dv in 1 0; do
    ( NI_WGRAD_DEFER=$dv timeout 600 python bench.py --batch $b --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 ) > gpurun_out/defer_b${b}_d${dv}.json
    python - <<PY
import json
d=json.load(open('gpurun_out/defer_b${b}_d${dv}.json'))
print('batch ${b} defer ${dv}: %.3f ms/step  e2e %.3f  loss %.4f  checksum %.9f' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['loss'], d['param_checksum']['l2']))
PY
  done
done
