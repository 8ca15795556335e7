#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_feed.py tests/test_conv_gpu.py -x -q -m gpu 2>&1 | tail -5 ) > gpurun_out/t8_tests.log; tail -2 gpurun_out/t8_tests.log
( timeout 200 python tools/profile_conv.py 0 3 2>&1 ) > gpurun_out/conv_2bprod.log; grep -A1 "fprop\|dgrad" gpurun_out/conv_2bprod.log | grep ms
( timeout 600 python bench.py --steps 10 --warmup 3 --layer-report gpurun_out/layers_r1d.json 2>&1 | tail -2 ) > gpurun_out/bench_r1d.log
head -c 900 gpurun_out/bench_r1d.log
