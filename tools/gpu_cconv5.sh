#!/bin/bash
# constrained 5x5 kernels with constant-bank weights / restructured filter gradient: tests, timing, headline bench
cd /root/repo
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 ) > gpurun_out/cconv5_tests.log; tail -3 gpurun_out/cconv5_tests.log
timeout 300 python tools/profile_manip.py 20 > gpurun_out/cconv5_time.json 2> gpurun_out/cconv5_time.err; cat gpurun_out/cconv5_time.json; tail -2 gpurun_out/cconv5_time.err
( timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 ) > gpurun_out/cconv5_bench_c4.json; head -c 330 gpurun_out/cconv5_bench_c4.json; echo
