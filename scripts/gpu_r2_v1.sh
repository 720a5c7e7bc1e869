#!/bin/bash
# round 2, visit 1: the whole GPU suite with the un-gated / new parity tests, generator microbenchmark, bench line, near-goal probe
TAG=${1:-r02a}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu_${TAG}.txt 2>&1
nproc >> $O/gpu_${TAG}.txt
timeout 1500 python -m pytest tests -m gpu -q --durations=15 2>&1 | tail -80 > $O/pytest_gpu_${TAG}.log
tail -5 $O/pytest_gpu_${TAG}.log
./profiles/microbench/gen > $O/gen_${TAG}.txt 2>&1; cat $O/gen_${TAG}.txt
timeout 600 python bench.py --steps 100 --warmup 10 > $O/bench_${TAG}.json 2> $O/bench_${TAG}.err
tail -c 2500 $O/bench_${TAG}.json; tail -3 $O/bench_${TAG}.err
timeout 300 python profiles/closed_loop_probe.py mixed > $O/closed_loop_${TAG}.txt 2>&1; cat $O/closed_loop_${TAG}.txt
