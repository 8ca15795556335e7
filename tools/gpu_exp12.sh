#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
( timeout 200 python tools/profile_conv.py 2 6 7 2>&1 ) > gpurun_out/conv_alt0.log
( NI_TC_ALT64=1 timeout 200 python tools/profile_conv.py 2 6 7 2>&1 ) > gpurun_out/conv_alt1.log
( NI_TC_ALT64=1 timeout 300 python -m pytest tests/test_conv_gpu.py -x -q -m gpu 2>&1 | tail -3 ) > gpurun_out/t12_alt.log; tail -1 gpurun_out/t12_alt.log
