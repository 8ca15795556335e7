#!/bin/bash
# everything the driver runs at round end, plus profiles: all GPU tests, smoke, bench (+cpu baseline), djpeg timing
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 1500 python -m pytest tests -m gpu -q --tb=line > gpurun_out/gpu_tests.log 2>&1; echo "pytest -m gpu exit $?"; tail -3 gpurun_out/gpu_tests.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 300 python tools/profile_djpeg.py 1280 20 > gpurun_out/djpeg_time.json 2>&1; cat gpurun_out/djpeg_time.json
timeout 1200 python bench.py --steps 5 --warmup 3 --layer-report gpurun_out/layers.json > gpurun_out/bench.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench.log | cut -c1-400
