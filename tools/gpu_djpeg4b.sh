#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
NI_B200_LIB=$PWD/neural_imaging_b200/libni_b200_dev.so timeout 300 python tools/profile_djpeg.py 1280 40 > gpurun_out/djpeg4b_variants.json 2> gpurun_out/djpeg4b_variants.err; cat gpurun_out/djpeg4b_variants.json; tail -3 gpurun_out/djpeg4b_variants.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:djpeg_fwd -s 2 -c 1 -o gpurun_out/prof_djpeg_fwd4 -f python tools/profile_djpeg.py 1280 1 > gpurun_out/ncu_djpeg4.log 2>&1; echo "ncu exit $?"; ls -la gpurun_out/*.ncu-rep
