#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 ) > gpurun_out/t7_tests.log; tail -2 gpurun_out/t7_tests.log
( timeout 600 python bench.py --steps 10 --warmup 3 --layer-report gpurun_out/layers_r1c.json 2>&1 | tail -2 ) > gpurun_out/bench_r1c.log
head -c 500 gpurun_out/bench_r1c.log
( timeout 300 python bench.py --steps 10 --warmup 3 --batch 32 --no-cpu-baseline 2>&1 | tail -1 ) > gpurun_out/bench32_r1c.log
head -c 300 gpurun_out/bench32_r1c.log
