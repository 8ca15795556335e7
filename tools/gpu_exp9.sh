#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
for d in 7 0; do ( NI_TC_DEFER=$d timeout 200 python tools/profile_conv.py 0 1 2 3 2>&1 ) > gpurun_out/conv_defer$d.log; done
echo ok
