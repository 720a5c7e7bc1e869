#!/bin/bash
# round 2: the kinematic user functor (mppi_user_model.kind 1) on the GPU + the neighbouring suites it touches on the host side
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_user_model.py tests/test_gpu_closed_loop.py -m gpu -x -q > gpurun_out/kin_tests.log 2>&1
echo "rc=$?" >> gpurun_out/kin_tests.log
tail -15 gpurun_out/kin_tests.log
