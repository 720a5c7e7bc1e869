#!/bin/bash
# mid-size GPU visit: tests, kernel-variant comparison, bench, launch list, one full ncu capture of the rollout kernel
TAG=${1:-ws}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -40 > $O/pytest_gpu_${TAG}.log
tail -5 $O/pytest_gpu_${TAG}.log
python profiles/variants.py 65536 64 "" block64 block128 fast > $O/variants_${TAG}.txt 2>&1
python profiles/variants.py 262144 128 "" block64 block128 fast >> $O/variants_${TAG}.txt 2>&1
python profiles/variants.py 262144 64 "" block64 block128 fast >> $O/variants_${TAG}.txt 2>&1
cat $O/variants_${TAG}.txt
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu > $O/bench_${TAG}.json 2> $O/bench_${TAG}.err
tail -c 2500 $O/bench_${TAG}.json; tail -3 $O/bench_${TAG}.err
for P in mixed; do
  timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 40 --csv \
    --log-file $O/launches_${P}_${TAG}.csv python profiles/profile_step.py $P 65536 64 5 > $O/ncu_${P}_${TAG}.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rollout -s 4 -c 1 \
  -o $O/prof_rollout_mixed_${TAG} -f python profiles/profile_step.py mixed 65536 64 4 > $O/ncu_full_${TAG}.log 2>&1
tail -3 $O/ncu_full_${TAG}.log
python profiles/reduce_timeline.py mixed > $O/reduce_timeline_${TAG}.txt 2>&1; cat $O/reduce_timeline_${TAG}.txt
