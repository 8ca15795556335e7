#!/bin/bash
# hardware probes + resident-weights / tap-rotation A/B + correctness
cd /root/repo
mkdir -p gpurun_out
( timeout 200 python tools/hw_probes.py 2>&1 ) > gpurun_out/hw_probes.log
( NI_TC_DEBUG=1 timeout 200 python tools/profile_conv.py 0 1 2 3 2>&1 ) > gpurun_out/conv_resident.log
( NI_TC_B_RESIDENT=0 NI_TC_ROT=7 timeout 200 python tools/profile_conv.py 0 1 2 3 2>&1 ) > gpurun_out/conv_rot7.log
( timeout 300 python -m pytest tests/test_conv_gpu.py tests/test_models_gpu.py -x -q -m gpu 2>&1 | tail -15 ) > gpurun_out/tests_resident.log
( NI_TC_CONV_SPLIT=1 timeout 300 python -m pytest tests/test_conv_gpu.py -x -q -m gpu 2>&1 | grep -v "^  " | tail -40 ) > gpurun_out/tests_split1.log
tail -3 gpurun_out/tests_resident.log; tail -3 gpurun_out/tests_split1.log
