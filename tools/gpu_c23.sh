#!/bin/bash
cd /root/repo
O=gpurun_out/final; mkdir -p $O
for c in c2 c3; do
  ( timeout 600 python bench.py --config $c --steps 10 --warmup 3 2>&1 | tail -1 ) > $O/bench_$c.json
  echo "$c: $(head -c 330 $O/bench_$c.json)"
done
