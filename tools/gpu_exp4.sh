#!/bin/bash
# generation 3b gemm: correctness (conv tests incl. cold multi-tile, 3 fresh processes), per-layer timing, step bench
cd /root/repo
mkdir -p gpurun_out
for i in 1 2 3; do ( timeout 300 python -m pytest tests/test_conv_gpu.py -x -q -m gpu -k "persistent or tc_path" 2>&1 | tail -4 ) > gpurun_out/t4_cold$i.log; tail -1 gpurun_out/t4_cold$i.log; done
( timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_models_gpu.py -x -q -m gpu 2>&1 | tail -8 ) > gpurun_out/t4_tests.log; tail -2 gpurun_out/t4_tests.log
( NI_TC_DEBUG=1 timeout 200 python tools/profile_conv.py 0 1 2 3 2>&1 ) > gpurun_out/conv_gen3b.log
( timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --layer-report gpurun_out/layers_gen3b.json 2>&1 | tail -3 ) > gpurun_out/bench_gen3b.log
head -c 600 gpurun_out/bench_gen3b.log
