#!/bin/bash
mkdir -p gpurun_out
timeout 140 python profiles/kinematic_timing.py > gpurun_out/kinematic_timing.txt 2>&1
echo "rc=$?" >> gpurun_out/kinematic_timing.txt
cat gpurun_out/kinematic_timing.txt
