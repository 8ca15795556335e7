#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 ) > gpurun_out/t11_tests.log; tail -2 gpurun_out/t11_tests.log
( timeout 900 python bench.py --layer-report gpurun_out/layers_r1e.json 2>&1 | tail -1 ) > gpurun_out/bench_r1e.json; head -c 400 gpurun_out/bench_r1e.json; echo
