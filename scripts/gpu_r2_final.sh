#!/bin/bash
# Round-2 evidence visit (1 GPU), most important first: tests, contract bench lines (both arms), launch lists, full ncu captures
# of the dominant kernels, all-config sweep, timelines, near-goal probe, microbenchmarks, sanitizer.
# usage (here): gpurun --timeout 2400 -- 'bash scripts/gpu_r2_final.sh [tag]'
TAG=${1:-r02}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu_${TAG}.txt 2>&1
nproc >> $O/gpu_${TAG}.txt; grep -m1 "model name" /proc/cpuinfo >> $O/gpu_${TAG}.txt
timeout 900 python -m pytest tests -m gpu -q --durations=10 2>&1 | tail -40 > $O/pytest_gpu_${TAG}.log
tail -3 $O/pytest_gpu_${TAG}.log
timeout 600 python bench.py --steps 100 --warmup 10 > $O/bench_${TAG}.json 2> $O/bench_${TAG}.err
tail -c 600 $O/bench_${TAG}.json; tail -3 $O/bench_${TAG}.err
timeout 300 python bench.py --impl reference --steps 10 --warmup 3 > $O/bench_reference_${TAG}.json 2> $O/bench_reference_${TAG}.err
tail -c 400 $O/bench_reference_${TAG}.json
for P in mixed f32 f64; do
  timeout 200 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 40 --csv \
    --log-file $O/launches_${P}_${TAG}.csv python profiles/profile_step.py $P 65536 64 5 > $O/ncu_${P}_${TAG}.log 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rollout -s 4 -c 1 \
  -o $O/prof_rollout_mixed_${TAG} -f python profiles/profile_step.py mixed 65536 64 4 > $O/ncu_full_${TAG}.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:reduce_screen -s 4 -c 1 \
  -o $O/prof_reduce_mixed_${TAG} -f python profiles/profile_step.py mixed 65536 64 4 >> $O/ncu_full_${TAG}.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rollout -s 4 -c 1 \
  -o $O/prof_rollout_c5_${TAG} -f python profiles/profile_step.py mixed 2097152 128 4 >> $O/ncu_full_${TAG}.log 2>&1
tail -2 $O/ncu_full_${TAG}.log
timeout 400 python profiles/sweep_configs.py 30 > $O/sweep_${TAG}.jsonl 2>&1
timeout 100 python profiles/reduce_timeline.py mixed > $O/reduce_timeline_${TAG}.txt 2>&1
timeout 100 python profiles/reduce_timeline.py mixed 262144 128 >> $O/reduce_timeline_${TAG}.txt 2>&1
timeout 300 python profiles/closed_loop_probe.py mixed > $O/closed_loop_${TAG}.txt 2>&1
timeout 100 python profiles/host_timing.py > $O/host_timing_${TAG}.txt 2>&1
./profiles/microbench/pipes > $O/pipes_${TAG}.txt 2>&1
./profiles/microbench/gen > $O/gen_${TAG}.txt 2>&1
timeout 600 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -12 > $O/sanitizer_memcheck_${TAG}.log
tail -2 $O/sanitizer_memcheck_${TAG}.log
timeout 600 compute-sanitizer --tool racecheck python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -12 > $O/sanitizer_racecheck_${TAG}.log
tail -2 $O/sanitizer_racecheck_${TAG}.log
ls $O | wc -l
