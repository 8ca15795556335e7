#!/bin/bash
# wgrad generation 3 (four producer warps): correctness, per-layer timing (vs generation 2), role spans, step bench
cd /root/repo
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_models_gpu.py -x -q -m gpu 2>&1 | tail -8 ) > gpurun_out/t5_tests.log; tail -2 gpurun_out/t5_tests.log
( NI_TC_DEBUG=1 timeout 200 python tools/profile_conv.py 0 1 2 3 2>&1 ) > gpurun_out/conv_wg3.log
( NI_TC_WGRAD_V2=1 timeout 200 python tools/profile_conv.py 0 1 2 3 2>&1 ) > gpurun_out/conv_wg2.log
( NI_B200_LIB=$PWD/neural_imaging_b200/libni_b200_prof.so NI_TC_DEBUG=1 timeout 200 python tools/tc_prof.py 0 2 2>&1 | grep -A12 "^wgrad" ) > gpurun_out/tcprof_wg3.log
( timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --layer-report gpurun_out/layers_wg3.json 2>&1 | tail -3 ) > gpurun_out/bench_wg3.log
head -c 400 gpurun_out/bench_wg3.log
