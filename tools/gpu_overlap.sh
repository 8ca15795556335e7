#!/bin/bash
# filter gradient on a side stream next to the input gradient of the same layer: test-suite with the overlap on, then A/B bench lines
cd /root/repo
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 ) > gpurun_out/overlap_tests.log; tail -3 gpurun_out/overlap_tests.log
for b in 256 32; do
  for ov in 1 0 1 0; do
    ( NI_WGRAD_OVERLAP=$ov timeout 600 python bench.py --batch $b --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 ) > gpurun_out/overlap_b${b}_ov${ov}.json
    python - <<PY
import json
d=json.load(open('gpurun_out/overlap_b${b}_ov${ov}.json'))
print('batch ${b} overlap ${ov}: %.3f ms/step  e2e %.3f  loss %.4f  checksum %.9f' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['loss'], d['param_checksum']['l2']))
PY
  done
done
