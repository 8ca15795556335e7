#!/bin/bash
./tools/gpu_tests.sh test_dcn_gpu
timeout 600 python tools/profile_dcn.py 64 64 5 2>&1 | tail -40
