#!/bin/bash
# Planned FIRST GPU call of round 2: the per-rank batch of the 8-GPU run (B = 32) is where the step is furthest from
# its ideal (12.2 vs 10.5 ms), and the full-resolution U-Net layers / the FAN 3->32 layer are where the B = 256 step loses its time
# (profiles/r1_layer_table_final.txt). One call: tests, B = 32 bench with a layer report, ncu launch list at B = 32, full captures of the two
# tcgen05 kernels on the shapes that sit at 25-50 % of the 3xTF32 ceiling, and the l3ic timing.
#   ./gpu.sh 1500 'bash tools/gpu_round2_first.sh'
cd /root/repo
mkdir -p gpurun_out
( timeout 300 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 ) > gpurun_out/r2_tests.log; tail -1 gpurun_out/r2_tests.log
( timeout 300 python bench.py --batch 32 --steps 10 --warmup 3 --no-cpu-baseline --layer-report gpurun_out/r2_layers_b32.json 2>&1 | tail -1 ) > gpurun_out/r2_bench_b32.json; cut -c1-300 gpurun_out/r2_bench_b32.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/r2_launches_b32.csv python bench.py --batch 32 --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/r2_ncu_launches.log 2>&1; echo "ncu launches exit $?"
# profile_conv.py shapes: 1 = 128x128 32->32 k3 (N = 32 tiles), 6 = 64x64 64->64 k3, 4 = FAN 3->32 k5 (direct FP32), 5 = 32->12 output conv
timeout 240 ncu --set full --clock-control none --import-source on -k regex:conv_tc3_gemm -s 2 -c 1 -o gpurun_out/r2_prof_gemm_n32 -f python tools/profile_conv.py 1 > gpurun_out/r2_ncu_conv.log 2>&1; echo "ncu gemm N=32 exit $?"
timeout 240 ncu --set full --clock-control none --import-source on -k regex:conv_tc3_wgrad -s 2 -c 1 -o gpurun_out/r2_prof_wgrad_n32 -f python tools/profile_conv.py 1 >> gpurun_out/r2_ncu_conv.log 2>&1; echo "ncu wgrad N=32 exit $?"
timeout 240 ncu --set full --clock-control none --import-source on -k regex:conv_fewin -s 2 -c 1 -o gpurun_out/r2_prof_fan_first -f python tools/profile_conv.py 4 >> gpurun_out/r2_ncu_conv.log 2>&1; echo "ncu FAN 3->32 exit $?"
timeout 120 python tools/profile_l3ic.py 10 > gpurun_out/r2_l3ic_time.json 2>&1; cut -c1-300 gpurun_out/r2_l3ic_time.json
ls -la gpurun_out | tail -12
