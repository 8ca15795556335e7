#!/bin/bash
# role profiles (profile build) + converter-split A/B timing + correctness of the split variant
cd /root/repo
mkdir -p gpurun_out
P=$PWD/neural_imaging_b200/libni_b200_prof.so
( NI_B200_LIB=$P NI_TC_DEBUG=1 timeout 200 python tools/tc_prof.py 2>&1 ) > gpurun_out/tcprof_split0.log
( NI_B200_LIB=$P NI_TC_CONV_SPLIT=1 timeout 200 python tools/tc_prof.py 2>&1 ) > gpurun_out/tcprof_split1.log
( timeout 200 python tools/profile_conv.py 0 1 2 3 2>&1 ) > gpurun_out/conv_split0.log
( NI_TC_CONV_SPLIT=1 timeout 200 python tools/profile_conv.py 0 1 2 3 2>&1 ) > gpurun_out/conv_split1.log
( NI_TC_CONV_SPLIT=1 timeout 300 python -m pytest tests/test_conv_gpu.py -x -q -m gpu 2>&1 | tail -5 ) > gpurun_out/conv_split1_tests.log
tail -3 gpurun_out/conv_split1_tests.log
