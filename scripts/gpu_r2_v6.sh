#!/bin/bash
# round 2, visit 6 (2 GPUs): multi-GPU tests, bench lines at N=1 and N=2 (parity + config5), closed-loop test file
N=${1:-2}
TAG=${2:-r02m2}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_closed_loop.py -m gpu -q 2>&1 | tail -8 > $O/pytest_cl_${TAG}.log; tail -3 $O/pytest_cl_${TAG}.log
bash scripts/gpu_multi.sh $N $TAG
