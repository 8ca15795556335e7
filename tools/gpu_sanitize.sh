#!/bin/bash
# compute-sanitizer passes over the kernels with hand-rolled pipelines: memcheck + racecheck on the dJPEG tests and on
# the tcgen05 conv tests (16-warp mbarrier / TMEM pipelines). Summaries -> gpurun_out/r2_sanitizer_*.log (copied to profiles/).
#   ./gpu.sh 1500 'bash tools/gpu_sanitize.sh'
cd /root/repo
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
run() {  # name tool timeout pytest-args...
  name=$1; tool=$2; t=$3; shift 3
  timeout $t $CS --tool $tool --error-exitcode 99 --print-limit 20 python -m pytest "$@" -x -q -m gpu -p no:cacheprovider > /tmp/san.log 2>&1
  ( grep -vE "^$|Host Frame" /tmp/san.log | head -60; echo ...; grep -vE "^$|Host Frame" /tmp/san.log | tail -8 ) > gpurun_out/r2_sanitizer_${name}_${tool}.log
  echo "$name $tool exit ${PIPESTATUS[0]}: $(grep -E 'ERROR SUMMARY|passed|failed|error' gpurun_out/r2_sanitizer_${name}_${tool}.log | tr '\n' ' ' | cut -c1-200)"
}
run djpeg memcheck 300 tests/test_djpeg_gpu.py
run djpeg racecheck 300 tests/test_djpeg_gpu.py
run manip memcheck 300 tests/test_manip_gpu.py
run conv memcheck 420 tests/test_conv_gpu.py -k "not many_tiles"
run conv racecheck 420 tests/test_conv_gpu.py -k "(tc or subpixel) and not many_tiles"
run dcn memcheck 420 tests/test_dcn_gpu.py tests/test_layers_gpu.py
