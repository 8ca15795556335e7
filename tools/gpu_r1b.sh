#!/bin/bash
# round-1 evidence bundle: GPU tests, smoke, dJPEG timing, full bench + layer report, ncu launch list, ncu --set full captures
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --tb=line -x > gpurun_out/gpu_tests.log 2>&1; echo "pytest -m gpu exit $?"; tail -3 gpurun_out/gpu_tests.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 300 python tools/profile_djpeg.py 1280 20 > gpurun_out/djpeg_time.json 2>&1; cat gpurun_out/djpeg_time.json
timeout 900 python bench.py --steps 5 --warmup 3 --layer-report gpurun_out/layers.json > gpurun_out/bench.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench.log | cut -c1-1500
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3400 --csv --log-file gpurun_out/launches_step.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu launches exit $?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_tc -c 9 -o gpurun_out/prof_conv -f python tools/profile_conv.py 2 0 > gpurun_out/ncu_conv.log 2>&1; echo "ncu conv exit $?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:djpeg -s 2 -c 2 -o gpurun_out/prof_djpeg -f python tools/profile_djpeg.py 1280 1 > gpurun_out/ncu_djpeg.log 2>&1; echo "ncu djpeg exit $?"
ls -la gpurun_out | head -30
