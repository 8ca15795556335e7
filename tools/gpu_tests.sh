#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
for f in "$@"; do
  timeout 900 python -m pytest tests/$f.py -m gpu -q --maxfail=40 --tb=line > gpurun_out/$f.log 2>&1
  echo "$f exit $?" >> gpurun_out/summary.txt
  grep -E "^/|^E |passed|failed|Error" gpurun_out/$f.log | cut -c1-300 | tail -25
done
cat gpurun_out/summary.txt
