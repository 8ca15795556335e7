#!/bin/bash
# final-tree evidence (round 2, last session): bench lines of the five configurations + B = 32 + reference arm, ncu launch list of one step
cd /root/repo
O=gpurun_out/final; mkdir -p $O
( timeout 900 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 ) > $O/bench_c4.json; echo "c4: $(head -c 330 $O/bench_c4.json)"
( timeout 600 python bench.py --batch 32 --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 ) > $O/bench_c4_b32.json; echo "c4 b32: $(head -c 330 $O/bench_c4_b32.json)"
for c in c1 c2 c3 c5; do
  ( timeout 600 python bench.py --config $c --steps 10 --warmup 3 2>&1 | tail -1 ) > $O/bench_$c.json
  echo "$c: $(head -c 330 $O/bench_$c.json)"
done
( timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 ) > $O/bench_reference_arm.json; echo "ref: $(head -c 300 $O/bench_reference_arm.json)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file $O/launches_step.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > $O/ncu_bench.log 2>&1; echo "ncu launches exit $?"
ls -la $O | tail -12
