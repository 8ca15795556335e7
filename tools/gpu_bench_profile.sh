#!/bin/bash
# full-size bench + ncu launch list of one step + ncu full capture of the dJPEG kernels
mkdir -p gpurun_out
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_full.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench_full.log | cut -c1-600
timeout 300 python tools/profile_djpeg.py 1280 20 > gpurun_out/djpeg_time.json 2>&1; cat gpurun_out/djpeg_time.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 576 -c 192 --csv --log-file gpurun_out/launches_step.csv python bench.py --steps 1 --warmup 3 --batch 64 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:djpeg -s 2 -c 2 -o gpurun_out/prof_djpeg -f python tools/profile_djpeg.py 1280 1 > gpurun_out/ncu_djpeg.log 2>&1; echo "ncu djpeg exit $?"
ls -la gpurun_out | head -30
