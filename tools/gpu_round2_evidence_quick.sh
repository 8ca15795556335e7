#!/bin/bash
# Subset of tools/gpu_round2_evidence.sh for changes that touch only the direct FP32 convolutions: GPU suite, smoke, the bench lines that
# contain those layers (c4, c4 at B = 32, c5), the launch list and the direct-kernel captures. Everything else in gpurun_out/r2 is kept.
cd /root/repo
O=gpurun_out/r2
mkdir -p $O
( timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -4 ) > $O/tests_gpu.log; tail -2 $O/tests_gpu.log
cp gpurun_out/parity_report.json $O/parity_report.json 2>/dev/null
( timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 ) > $O/smoke.log; cat $O/smoke.log
( timeout 900 python bench.py --steps 20 --warmup 5 --layer-report $O/layers_b256.json 2>&1 | tail -1 ) > $O/bench_c4.json; head -c 300 $O/bench_c4.json; echo
( timeout 300 python bench.py --batch 32 --steps 20 --warmup 5 --no-cpu-baseline --layer-report $O/layers_b32.json 2>&1 | tail -1 ) > $O/bench_c4_b32.json; head -c 260 $O/bench_c4_b32.json; echo
( timeout 900 python bench.py --config c5 --steps 10 --warmup 3 --sweep 2>&1 | tail -1 ) > $O/bench_c5.json; echo "c5: $(head -c 260 $O/bench_c5.json)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file $O/launches_step.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > $O/ncu_bench.log 2>&1; echo "ncu launches exit $?"
timeout 300 ncu --set full --clock-control none -k regex:"fewin|manyin3|direct_wgrad" -s 6 -c 3 -o $O/prof_direct -f python tools/profile_conv.py 4 > $O/ncu_direct.log 2>&1; echo "ncu direct exit $?"
timeout 120 python tools/profile_conv.py 4 5 > $O/direct_time.json 2>&1
ncu -i $O/prof_direct.ncu-rep --page raw --csv > $O/prof_direct.raw.csv 2>/dev/null
du -sm gpurun_out | tail -1
