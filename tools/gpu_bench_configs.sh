#!/bin/bash
# bench.py on every BASELINE.json configuration (c4 = headline = default), 1 GPU. Lines -> gpurun_out/r2_bench_<config>.json
#   ./gpu.sh 1800 'bash tools/gpu_bench_configs.sh'
cd /root/repo
mkdir -p gpurun_out
for c in c1 c2 c3 c5; do
  extra=""; [ $c = c5 ] && extra="--sweep"
  ( timeout 600 python bench.py --config $c --steps 10 --warmup 3 $extra 2>&1 | tail -1 ) > gpurun_out/r2_bench_$c.json
  echo "$c: $(head -c 420 gpurun_out/r2_bench_$c.json)"
done
( timeout 900 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 ) > gpurun_out/r2_bench_c4.json; echo "c4: $(head -c 420 gpurun_out/r2_bench_c4.json)"
