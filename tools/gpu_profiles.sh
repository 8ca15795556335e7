#!/bin/bash
# ncu evidence for profiles/: launch list of one step (kernel-by-kernel launches), --set full captures of dJPEG fwd/bwd and of the top conv kernels
cd /root/repo
mkdir -p gpurun_out
timeout 300 python tools/profile_djpeg.py 1280 20 > gpurun_out/djpeg_time.json 2>&1; cat gpurun_out/djpeg_time.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches_step.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_bench.log 2>&1; echo "ncu launches exit $?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:djpeg -s 3 -c 2 -o gpurun_out/prof_djpeg -f python tools/profile_djpeg.py 1280 1 > gpurun_out/ncu_djpeg.log 2>&1; echo "ncu djpeg exit $?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_tc3_gemm -s 2 -c 1 -o gpurun_out/prof_conv_fprop -f python tools/profile_conv.py 2 > gpurun_out/ncu_conv.log 2>&1; echo "ncu conv fprop exit $?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_tc3_wgrad -s 2 -c 1 -o gpurun_out/prof_conv_wgrad -f python tools/profile_conv.py 2 >> gpurun_out/ncu_conv.log 2>&1; echo "ncu conv wgrad exit $?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_tc3_gemm -s 2 -c 1 -o gpurun_out/prof_conv_fprop128 -f python tools/profile_conv.py 3 >> gpurun_out/ncu_conv.log 2>&1; echo "ncu conv fprop128 exit $?"
ls -la gpurun_out | tail -12
