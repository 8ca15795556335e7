#!/bin/bash
# last call of the round: full GPU test-suite, smoke(), headline bench + B = 32, ncu launch list on the final tree
cd /root/repo
O=gpurun_out/last; mkdir -p $O
( timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 ) > $O/tests_gpu.log; tail -2 $O/tests_gpu.log
( timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 ) > $O/smoke.log; cat $O/smoke.log
( timeout 900 python bench.py 2>&1 | tail -1 ) > $O/bench_c4.json; echo "c4: $(head -c 330 $O/bench_c4.json)"
( timeout 600 python bench.py --batch 32 --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 ) > $O/bench_c4_b32.json; echo "c4 b32: $(head -c 330 $O/bench_c4_b32.json)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file $O/launches_step.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > $O/ncu_bench.log 2>&1; echo "ncu launches exit $?"
