#!/bin/bash
# Copy / summarise what tools/gpu_round2_evidence.sh brought back (gpurun_out/r2/) into profiles/r2_* (run in the dev container).
cd /root/repo
S=gpurun_out/r2; P=profiles
for c in c1 c2 c3 c4 c5; do cp $S/bench_$c.json $P/r2_bench_$c.json; done
cp $S/bench_c4_b32.json $P/r2_bench_c4_b32.json; cp $S/bench_reference_arm.json $P/r2_bench_reference_arm.json
cp $S/layers_b256.json $P/r2_layers_b256.json; cp $S/layers_b32.json $P/r2_layers_b32.json
cp $S/launches_step.csv $P/r2_launches_step.csv; python tools/launch_shares.py $S/launches_step.csv > $P/r2_launch_shares_step.txt 2>&1
cp $S/parity_report.json $P/parity_report.json; cp $S/djpeg_time.json $P/r2_djpeg_time.json; cp $S/manip_time.json $P/r2_manip_time.json
cp $S/direct_time.json $P/r2_direct_time.json 2>/dev/null
cat $S/tests_gpu.log $S/smoke.log > $P/r2_tests_gpu.log
python tools/ncu_summary.py $S/prof_djpeg.ncu-rep > $P/r2_djpeg_ncu.txt 2>&1
python tools/ncu_summary.py $S/prof_conv_fprop.ncu-rep > $P/r2_conv_tc_ncu.txt 2>&1; python tools/ncu_summary.py $S/prof_conv_wgrad.ncu-rep >> $P/r2_conv_tc_ncu.txt 2>&1
python tools/ncu_summary.py $S/prof_manip.ncu-rep > $P/r2_manip_cconv5_ncu.txt 2>&1
[ -f $S/prof_direct.ncu-rep ] && python tools/ncu_summary.py $S/prof_direct.ncu-rep > $P/r2_direct_ncu.txt 2>&1
[ -f $S/prof_latent.ncu-rep ] && python tools/ncu_summary.py $S/prof_latent.ncu-rep > $P/r2_latent_ncu.txt 2>&1
python tools/ncu_traffic.py $S/prof_djpeg.ncu-rep $S/prof_conv_fprop.ncu-rep $P/r2_ncu_traffic.json > /dev/null
python - <<'PY'
import json
for f in ['bench_c4','bench_c4_b32','bench_c1','bench_c2','bench_c3','bench_c5','bench_reference_arm']:
    d=json.loads(open('gpurun_out/r2/%s.json'%f).read().strip().splitlines()[-1])
    print('%-22s %12.2f %-10s %9.3f ms/step  e2e %s  roofline.frac %s' % (f, d.get('value') or 0, d.get('unit'), d.get('ms_per_step') or 0, (d.get('e2e') or {}).get('value'), (d.get('roofline') or {}).get('frac')))
PY
