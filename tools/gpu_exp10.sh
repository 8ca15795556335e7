#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_metrics_gpu.py tests/test_dcn_gpu.py -x -q -m gpu 2>&1 | tail -3 ) > gpurun_out/t10_tests.log; tail -1 gpurun_out/t10_tests.log
( NI_TC_DEBUG=1 timeout 200 python tools/profile_conv.py 0 3 2>&1 ) > gpurun_out/conv_slots2.log
( NI_TC_DEBUG=1 NI_TC_SLOTS4=100000 timeout 200 python tools/profile_conv.py 0 3 2>&1 ) > gpurun_out/conv_slots4.log
grep '"ms"' gpurun_out/conv_slots2.log | tr '\n' ' '; echo; grep '"ms"' gpurun_out/conv_slots4.log | tr '\n' ' '; echo
( NI_TC_SLOTS4=100000 timeout 200 python -m pytest tests/test_conv_gpu.py -x -q -m gpu 2>&1 | tail -3 ) > gpurun_out/t10_slots4.log; tail -1 gpurun_out/t10_slots4.log
