#!/bin/bash
# generation-4 dJPEG forward (persistent CTAs, TMA ring): parity tests on the shipping library, then the variant sweep on the development library
cd /root/repo
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_djpeg_gpu.py tests/test_tf_graph_golden_gpu.py -x -q -m gpu 2>&1 | tail -15 ) > gpurun_out/djpeg4_tests.log; tail -3 gpurun_out/djpeg4_tests.log
NI_B200_LIB=$PWD/neural_imaging_b200/libni_b200_dev.so timeout 300 python tools/profile_djpeg.py 1280 40 > gpurun_out/djpeg4_variants.json 2> gpurun_out/djpeg4_variants.err; cat gpurun_out/djpeg4_variants.json; tail -3 gpurun_out/djpeg4_variants.err
timeout 300 python tools/profile_djpeg.py 1280 40 > gpurun_out/djpeg4_default.json 2>> gpurun_out/djpeg4_variants.err; cat gpurun_out/djpeg4_default.json
( timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 ) > gpurun_out/djpeg4_alltests.log; tail -2 gpurun_out/djpeg4_alltests.log
