#!/bin/bash
# Round-2 evidence for profiles/: full GPU test-suite (+ parity report), smoke, bench lines of every BASELINE configuration, the reference
# arm, the ncu launch list of one step, `ncu --set full` captures of the dominant kernels and of the new HBM-side kernels.
#   ./gpu.sh 2400 'bash tools/gpu_round2_evidence.sh'
cd /root/repo
O=gpurun_out/r2
mkdir -p $O
( timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -4 ) > $O/tests_gpu.log; tail -2 $O/tests_gpu.log
cp gpurun_out/parity_report.json $O/parity_report.json 2>/dev/null
( timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 ) > $O/smoke.log; cat $O/smoke.log
( timeout 900 python bench.py --steps 20 --warmup 5 --layer-report $O/layers_b256.json 2>&1 | tail -1 ) > $O/bench_c4.json; head -c 300 $O/bench_c4.json; echo
( timeout 300 python bench.py --batch 32 --steps 20 --warmup 5 --no-cpu-baseline --layer-report $O/layers_b32.json 2>&1 | tail -1 ) > $O/bench_c4_b32.json; head -c 260 $O/bench_c4_b32.json; echo
for c in c1 c2 c3; do ( timeout 600 python bench.py --config $c --steps 20 --warmup 5 2>&1 | tail -1 ) > $O/bench_$c.json; echo "$c: $(head -c 260 $O/bench_$c.json)"; done
( timeout 900 python bench.py --config c5 --steps 10 --warmup 3 --sweep 2>&1 | tail -1 ) > $O/bench_c5.json; echo "c5: $(head -c 260 $O/bench_c5.json)"
( timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 ) > $O/bench_reference_arm.json; head -c 300 $O/bench_reference_arm.json; echo
timeout 120 python tools/profile_djpeg.py 1280 20 > $O/djpeg_time.json 2>&1; cat $O/djpeg_time.json
timeout 120 python tools/profile_manip.py 10 > $O/manip_time.json 2>&1; cat $O/manip_time.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file $O/launches_step.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > $O/ncu_bench.log 2>&1; echo "ncu launches exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:djpeg -s 3 -c 2 -o $O/prof_djpeg -f python tools/profile_djpeg.py 1280 1 > $O/ncu_djpeg.log 2>&1; echo "ncu djpeg exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc3_gemm -s 2 -c 1 -o $O/prof_conv_fprop -f python tools/profile_conv.py 2 > $O/ncu_conv.log 2>&1; echo "ncu conv fprop exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc3_wgrad -s 2 -c 1 -o $O/prof_conv_wgrad -f python tools/profile_conv.py 2 >> $O/ncu_conv.log 2>&1; echo "ncu conv wgrad exit $?"
timeout 300 ncu --set full --clock-control none -k regex:"manip_stack|cconv5" -c 5 -o $O/prof_manip -f python tools/profile_manip.py 0 > $O/ncu_manip.log 2>&1; echo "ncu manip exit $?"
timeout 300 ncu --set full --clock-control none -k regex:"fewin|manyin3|direct_wgrad" -s 6 -c 3 -o $O/prof_direct -f python tools/profile_conv.py 4 > $O/ncu_direct.log 2>&1; echo "ncu direct exit $?"
timeout 300 ncu --set full --clock-control none -k regex:"latent_" -s 4 -c 2 -o $O/prof_latent -f python bench.py --config c3 --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_latent.log 2>&1; echo "ncu latent exit $?"
timeout 120 python tools/profile_conv.py 4 5 > $O/direct_time.json 2>&1
# gpurun copies back at most 64 MiB: summarise the captures on the box and drop the raw reports that do not fit
for f in prof_djpeg prof_conv_fprop prof_conv_wgrad prof_manip prof_direct prof_latent; do
  ncu -i $O/$f.ncu-rep --page raw --csv > $O/$f.raw.csv 2>/dev/null
done
du -sm gpurun_out | tail -1
while [ $(du -sm gpurun_out | cut -f1) -gt 58 ]; do big=$(ls -S $O/*.ncu-rep 2>/dev/null | head -1); [ -z "$big" ] && break; echo "dropping $big (raw CSV kept)"; rm -f $big; done
ls -la $O | tail -30
