#!/bin/bash
# end-of-round check: full GPU test-suite, smoke(), default bench (N = 1), reference arm
cd /root/repo
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 ) > gpurun_out/final_tests.log; tail -2 gpurun_out/final_tests.log
( timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 ) > gpurun_out/final_smoke.log; cat gpurun_out/final_smoke.log
( timeout 900 python bench.py 2>&1 | tail -1 ) > gpurun_out/final_bench.json; head -c 700 gpurun_out/final_bench.json; echo
( timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 ) > gpurun_out/final_reference.json; head -c 400 gpurun_out/final_reference.json; echo
