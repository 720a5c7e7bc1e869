#!/bin/bash
# One full GPU visit: tests, bench (both arms), C++ facade demo, microbenchmark, ncu launch lists + full captures,
# reduce timeline, all-config sweep.  usage (here): gpurun --timeout 1800 -- 'bash scripts/gpu_round.sh [tag]'
TAG=${1:-r01}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu_${TAG}.txt 2>&1
nproc >> $O/gpu_${TAG}.txt; grep -m1 "model name" /proc/cpuinfo >> $O/gpu_${TAG}.txt
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 > $O/pytest_gpu_${TAG}.log
tail -3 $O/pytest_gpu_${TAG}.log
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > $O/bench_reference_${TAG}.json 2> $O/bench_reference_${TAG}.err
timeout 600 python bench.py --steps 100 --warmup 10 > $O/bench_${TAG}.json 2> $O/bench_${TAG}.err
tail -c 1500 $O/bench_${TAG}.json; tail -3 $O/bench_${TAG}.err
g++ -std=c++17 -Iinclude examples/ros_control_node.cpp -Lmotion_planning_b200/lib -lmppi_b200 \
  -Wl,-rpath,$PWD/motion_planning_b200/lib -o /tmp/ros_node 2>&1 | grep -v Wcomment | head -5
timeout 300 /tmp/ros_node 4096 64 3000 > $O/cpp_node_${TAG}.log 2>&1; tail -2 $O/cpp_node_${TAG}.log
./profiles/microbench/pipes > $O/pipes_${TAG}.txt 2>&1
for P in mixed f32 f64; do
  timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 40 --csv \
    --log-file $O/launches_${P}_${TAG}.csv python profiles/profile_step.py $P 65536 64 5 > $O/ncu_${P}_${TAG}.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rollout -s 4 -c 1 \
  -o $O/prof_rollout_mixed_${TAG} -f python profiles/profile_step.py mixed 65536 64 4 > $O/ncu_full_${TAG}.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:reduce_screen -s 4 -c 1 \
  -o $O/prof_reduce_mixed_${TAG} -f python profiles/profile_step.py mixed 65536 64 4 >> $O/ncu_full_${TAG}.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rollout -s 4 -c 1 \
  -o $O/prof_rollout_f32_${TAG} -f python profiles/profile_step.py f32 65536 64 4 >> $O/ncu_full_${TAG}.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rollout -s 4 -c 1 \
  -o $O/prof_rollout_c5_${TAG} -f python profiles/profile_step.py mixed 262144 128 4 >> $O/ncu_full_${TAG}.log 2>&1
python profiles/reduce_timeline.py mixed > $O/reduce_timeline_${TAG}.txt 2>&1
python profiles/sweep_configs.py 30 > $O/sweep_${TAG}.jsonl 2>&1
ls $O | wc -l
for TOOL in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $TOOL python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -12 > $O/sanitizer_${TOOL}_${TAG}.log
  tail -2 $O/sanitizer_${TOOL}_${TAG}.log
done
