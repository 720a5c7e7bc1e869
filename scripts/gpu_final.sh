#!/bin/bash
# Budget-bounded GPU visit, most important evidence first: tests, contract bench line (both arms), launch list,
# full ncu captures of the two kernels, all-config sweep, reduce timeline, sanitizer.
# usage (here): gpurun --timeout 1080 -- 'bash scripts/gpu_final.sh [tag]'
TAG=${1:-r01L}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu_${TAG}.txt 2>&1
nproc >> $O/gpu_${TAG}.txt; grep -m1 "model name" /proc/cpuinfo >> $O/gpu_${TAG}.txt
timeout 420 python -m pytest tests -m gpu -q 2>&1 | tail -30 > $O/pytest_gpu_${TAG}.log
tail -3 $O/pytest_gpu_${TAG}.log
timeout 300 python bench.py --steps 100 --warmup 10 > $O/bench_${TAG}.json 2> $O/bench_${TAG}.err
tail -c 1200 $O/bench_${TAG}.json; tail -3 $O/bench_${TAG}.err
timeout 200 python bench.py --impl reference --steps 10 --warmup 3 > $O/bench_reference_${TAG}.json 2> $O/bench_reference_${TAG}.err
timeout 200 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 40 --csv \
  --log-file $O/launches_mixed_${TAG}.csv python profiles/profile_step.py mixed 65536 64 5 > $O/ncu_mixed_${TAG}.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rollout -s 4 -c 1 \
  -o $O/prof_rollout_mixed_${TAG} -f python profiles/profile_step.py mixed 65536 64 4 > $O/ncu_full_${TAG}.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:reduce_screen -s 4 -c 1 \
  -o $O/prof_reduce_mixed_${TAG} -f python profiles/profile_step.py mixed 65536 64 4 >> $O/ncu_full_${TAG}.log 2>&1
tail -2 $O/ncu_full_${TAG}.log
timeout 300 python profiles/sweep_configs.py 30 > $O/sweep_${TAG}.jsonl 2>&1
timeout 100 python profiles/reduce_timeline.py mixed > $O/reduce_timeline_${TAG}.txt 2>&1
for P in f32 f64; do
  timeout 200 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 40 --csv \
    --log-file $O/launches_${P}_${TAG}.csv python profiles/profile_step.py $P 65536 64 5 > $O/ncu_${P}_${TAG}.log 2>&1
done
./profiles/microbench/pipes > $O/pipes_${TAG}.txt 2>&1
timeout 300 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -12 > $O/sanitizer_memcheck_${TAG}.log
tail -2 $O/sanitizer_memcheck_${TAG}.log
ls $O | wc -l
