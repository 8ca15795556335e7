#!/bin/bash
# usage (from the dev container): tools/gpu_tcprof.sh  -- builds the profiling variant, runs tools/tc_prof.py on a B200, rebuilds the normal library
cd /root/repo
NI_NVCC_EXTRA=-DNI_TC_PROFILE python neural_imaging_b200/build.py --force > /dev/null
/usr/local/graft/bin/gpurun --timeout 600 -- "timeout 300 python tools/tc_prof.py 2>&1 | tail -70"
python neural_imaging_b200/build.py --force > /dev/null
