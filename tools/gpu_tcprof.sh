#!/bin/bash
# usage (from the dev container): tools/gpu_tcprof.sh  -- builds the DEVELOPMENT + profiling variant (libni_b200_dev.so: -DNI_DEV -DNI_TC_PROFILE,
# probes and the in-kernel role profiler), runs tools/tc_prof.py with it on a B200, removes the variant again
cd /root/repo
NI_BUILD_TAG=dev NI_NVCC_EXTRA=-DNI_TC_PROFILE python neural_imaging_b200/build.py --force > /dev/null
/usr/local/graft/bin/gpurun --timeout 600 -- "NI_B200_LIB=neural_imaging_b200/libni_b200_dev.so timeout 300 python tools/tc_prof.py 2>&1 | tail -70"
rm -rf neural_imaging_b200/libni_b200_dev.so neural_imaging_b200/build_dev
