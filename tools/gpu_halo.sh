#!/bin/bash
mkdir -p gpurun_out
./tools/gpu_tests.sh test_conv_gpu test_models_gpu
timeout 300 python tools/profile_conv.py > gpurun_out/tc_conv_halo.log 2>&1; grep -E "tflops" -B1 gpurun_out/tc_conv_halo.log | grep -v "^--" | paste - - | sed 's/ \+/ /g' | cut -c1-150
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --layer-report gpurun_out/layers_halo.json > gpurun_out/bench_halo.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench_halo.log | cut -c1-300
