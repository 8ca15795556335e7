#!/bin/bash
# usage: ./gpu.sh <timeout-seconds> '<command>'  -- rebuilds libni_b200.so here (nvcc cross-compiles), then runs on a B200
set -e
cd /root/repo
python neural_imaging_b200/build.py > /dev/null
T=$1; shift
/usr/local/graft/bin/gpurun --timeout $T -- "$@"
