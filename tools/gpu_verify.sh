#!/bin/bash
# verification of the current tree on one B200: full GPU test-suite, smoke(), sanitizer passes over the dJPEG kernels (generation-4 forward),
# bench lines of config 1 (dJPEG round trip) and config 4 (headline)
cd /root/repo
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 ) > gpurun_out/verify_tests.log; tail -2 gpurun_out/verify_tests.log
( timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 ) > gpurun_out/verify_smoke.log; cat gpurun_out/verify_smoke.log
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck; do
  timeout 400 $CS --tool $tool --error-exitcode 99 --print-limit 20 python -m pytest tests/test_djpeg_gpu.py -x -q -m gpu -p no:cacheprovider > /tmp/san.log 2>&1
  echo "djpeg $tool exit $?"
  ( grep -vE "^$|Host Frame" /tmp/san.log | head -40; echo ...; grep -vE "^$|Host Frame" /tmp/san.log | tail -8 ) > gpurun_out/r2_sanitizer_djpeg4_${tool}.log
  grep -E 'ERROR SUMMARY|passed|failed' gpurun_out/r2_sanitizer_djpeg4_${tool}.log | tail -2
done
( timeout 600 python bench.py --config c1 2>&1 | tail -1 ) > gpurun_out/verify_bench_c1.json; head -c 900 gpurun_out/verify_bench_c1.json; echo
( timeout 900 python bench.py 2>&1 | tail -1 ) > gpurun_out/verify_bench_c4.json; head -c 600 gpurun_out/verify_bench_c4.json; echo
