#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
for np in 4 8; do for df in 0 1; do
( NI_TC_WG_NP=$np NI_TC_WG_DEFER=$df timeout 200 python tools/profile_conv.py 0 1 2 3 2>&1 ) > gpurun_out/conv_wg3_np${np}_d${df}.log
done; done
echo done
