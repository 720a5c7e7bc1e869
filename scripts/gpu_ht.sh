#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 100 python profiles/host_timing.py > $O/host_timing_r02.txt 2>&1; cat $O/host_timing_r02.txt
