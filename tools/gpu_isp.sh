#!/bin/bash
./tools/gpu_tests.sh test_isp_gpu test_dcn_gpu
