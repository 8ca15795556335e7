#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
(
for d in 0 1 2 4 3; do
echo "== resident on, dbg $d"; NI_TC_DBG=$d timeout 120 python tools/tc_repro3.py 2>&1 | grep -v "^    "
done
) > gpurun_out/repro3.log
cat gpurun_out/repro3.log
